"""Builds libvt_b200.so in-tree with nvcc for sm_100a (B200).  `python -m vla_touch_b200.build [--force]`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libvt_b200.so")
SOURCES = ["vt_host.cu"]
DEPS = ["vt_host.cu", "vt_wgrad.cuh", "vt_persist.cuh", "vt_attn_pp.cuh", "vt_resize.cuh", "vt_dataset.cuh", "vt_lstm_tc.cuh", "vt_gemm.cuh", "vt_attn.cuh", "vt_mlp.cuh", "vt_rowproj.cuh", "vt_elem.cuh", "vt_lstm.cuh", "vt_bwd.cuh", "vt_ptx.cuh",
        os.path.join("..", "..", "include", "vt_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, debug_knobs: bool = False) -> str:
    """debug_knobs: compile the developer knobs / timestamps of csrc/vt_gemm.cuh in (VT_GEMM_DEBUG env variable; tools/ only) --
    into a SEPARATE file, lib/libvt_b200_dbg.so (selected with VT_LIB=...), so that the release library is never replaced."""
    if debug_knobs:
        return _compile(OUT.replace("libvt_b200.so", "libvt_b200_dbg.so"), ["-DVT_DEBUG_KNOBS=1"], verbose, "build_dbg.log")
    if not force and not stale():
        return OUT
    return _compile(OUT, [], verbose, "build.log")


def _compile(out: str, extra, verbose: bool, logname: str) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "lib", logname), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    if verbose:
        print(log)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv or "--debug-knobs" in sys.argv, verbose="-v" in sys.argv, debug_knobs="--debug-knobs" in sys.argv))
