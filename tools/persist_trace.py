"""Developer tool (GPU box): timeline of the persistent sampler kernel (csrc/vt_persist.cuh) from the instrumented library.

    python -m vla_touch_b200.build --debug-knobs          # -> vla_touch_b200/lib/libvt_b200_dbg.so (here, nvcc cross-compiles)
    VT_LIB=vla_touch_b200/lib/libvt_b200_dbg.so python tools/persist_trace.py [batch] [out.npz]   # on the GPU box

Per (worker = CTA pair, local tile) the leader CTA records clock64() at: 1/2 the TMA thread before / after the dependency wait,
3 the MMA thread owns the accumulator, 4 first operand stage landed, 5 last MMA + commit issued, 6 the first epilogue warp enters
the tile, 7 accumulator complete (acc_full observed), 8 GroupNorm statistics done, 9 epilogue body done, 10 tile published.
Prints, per layer of one steady-state step: tiles, and the mean cycles of each phase; then the utilisation of the pairs."""
import ctypes as C
import os
import sys

os.environ["VT_GEMM_DEBUG"] = "512"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("VT_LIB", os.path.join(ROOT, "vla_touch_b200", "lib", "libvt_b200_dbg.so"))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from vla_touch_b200 import native as nv
from vla_touch_b200 import shapes as shp
from vla_touch_b200 import synthetic as syn
from vla_touch_b200.engine import BridgeEngine


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "persist_trace.npz")
    A, T = 7, 64
    v = syn.synth_state_dict(shp.unet_shapes(A), 1, "v.")
    s = syn.synth_state_dict(shp.unet_shapes(A), 2, "s.")
    eng = BridgeEngine(dino=None, enc_sd=None, v_sd=v, s_sd=s, action_dim=A, state_dim=A, force_dim=64, use_force=True, B=B, T=T,
                       device="cuda:0")
    eng.cond.copy_(syn.det_normal("t.cond", (B, 256), 1).cuda())
    eng.x.copy_(syn.det_uniform("t.x", (B, T, A), 1, -1.0, 1.0).cuda())
    eng.run_ranges(["film_c", "xprior"])
    L = nv.lib()
    L.vt_debug_persist_trace.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
    dims = (C.c_int32 * 3)()
    if L.vt_debug_persist_trace(None, None, dims) != 0:
        raise SystemExit("library built without the instrumentation (python -m vla_touch_b200.build --debug-knobs; VT_LIB=...)")
    W, NT, NS = dims[0], dims[1], dims[2]
    tr = np.zeros((W, NT, NS), np.int64)
    cal = np.zeros((W, 4), np.int64)
    for _ in range(2):
        eng.run_steps()
        torch.cuda.synchronize()
    L.vt_debug_persist_trace(tr.ctypes.data, cal.ctypes.data, dims)     # discard the warm-up
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run_steps()
    e1.record()
    torch.cuda.synchronize()
    L.vt_debug_persist_trace(tr.ctypes.data, cal.ctypes.data, dims)
    d = eng.plan.descs[eng.sampler_range[0]]
    np.savez_compressed(out, tr=tr, cal=cal, tags=np.array(d.tags), ms=e0.elapsed_time(e1))
    print(f"launch {e0.elapsed_time(e1):.3f} ms, {eng.n_steps} steps, batch {B}; trace -> {out}")
    report(tr, cal, list(d.tags))


def report(tr, cal, tags, step=2):
    W = int((cal[:, 0] != 0).sum())
    # common time base: ns = gt0 + (clk - clk0) * rate, rate from the two calibration points of each worker
    rate = (cal[:W, 3] - cal[:W, 1]) / np.maximum(cal[:W, 2] - cal[:W, 0], 1)          # ns per cycle
    t0 = cal[:W, 1].min()
    ghz = 1.0 / rate.mean()
    print(f"{W} workers, SM clock {ghz:.3f} GHz, kernel {(cal[:W, 3].max() - t0) / 1e6:.3f} ms on the globaltimer")

    def to_ns(w, clk):
        return cal[w, 1] - t0 + (clk - cal[w, 0]) * rate[w]

    rows = []
    for w in range(W):
        for k in range(tr.shape[1]):
            r = tr[w, k]
            if r[6] == 0:
                continue
            st, l, tile = int(r[0] >> 40), int((r[0] >> 20) & 0xFFFFF), int(r[0] & 0xFFFFF)
            rows.append((st, l, tile, w, r))
    steps = sorted({x[0] for x in rows})
    print("steps traced:", steps)
    sel = [x for x in rows if x[0] == step]
    print(f"\nstep {step}: per layer, mean cycles over the layer's tiles (leader CTA)")
    print(f"{'layer':44s} {'tiles':>5s} {'depwait':>8s} {'ops->1st':>8s} {'mainloop':>8s} {'acc wait':>8s} {'epi p1':>7s} {'epi p2':>7s} {'publish':>7s}"
          f" {'span us':>8s} {'start us':>8s}")
    step_t0 = min(to_ns(x[3], x[4][1]) for x in sel)
    for l in sorted({x[1] for x in sel}):
        xs = [x for x in sel if x[1] == l]
        f = lambda a, b: np.mean([x[4][b] - x[4][a] for x in xs if x[4][a] and x[4][b]]) if xs else 0
        start = min(to_ns(x[3], x[4][1]) for x in xs)
        end = max(to_ns(x[3], x[4][10]) for x in xs)
        # the main loop cannot start before both the accumulator is free (3) and the dependencies are met (2)
        first = np.mean([x[4][4] - max(x[4][3], x[4][2]) for x in xs])
        print(f"{l:2d} {tags[l][5:45]:41s} {len(xs):5d} {f(1, 2):8.0f} {first:8.0f} {f(4, 5):8.0f} {f(6, 7):8.0f} {f(7, 8):7.0f} {f(8, 9):7.0f} {f(9, 10):7.0f}"
              f" {(end - start) / 1e3:8.1f} {(start - step_t0) / 1e3:8.1f}")
    step_end = max(to_ns(x[3], x[4][10]) for x in sel)
    print(f"step {step} span {(step_end - step_t0) / 1e3:.1f} us")
    # utilisation of each pair inside the step: MMA-thread busy (4 -> 5), epilogue busy (7 -> 10)
    mma = np.zeros(W)
    epi = np.zeros(W)
    for st, l, tile, w, r in sel:
        mma[w] += (r[5] - r[4]) * rate[w]
        epi[w] += (r[10] - r[7]) * rate[w]
    span = step_end - step_t0
    print(f"pairs: main loop busy {100 * mma.mean() / span:.1f} % (min {100 * mma.min() / span:.1f}, max {100 * mma.max() / span:.1f}), "
          f"epilogue busy {100 * epi.mean() / span:.1f} %")
    # timeline of one pair
    w = 0
    print(f"\nworker {w}, step {step}: tile by tile (us from the step's start)")
    for st, l, tile, ww, r in sorted([x for x in sel if x[3] == w], key=lambda x: x[4][1]):
        u = lambda k: (to_ns(w, r[k]) - step_t0) / 1e3
        print(f"  L{l:2d} tile {tile:3d}: dep {u(1):7.1f}->{u(2):7.1f}  mma {u(3):7.1f} first {u(4):7.1f} issued {u(5):7.1f} | epi enter {u(6):7.1f} acc {u(7):7.1f} p1 {u(8):7.1f} done {u(9):7.1f} pub {u(10):7.1f}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1].endswith(".npz"):
        z = np.load(sys.argv[1])
        report(z["tr"], z["cal"], [str(t) for t in z["tags"]], int(sys.argv[2]) if len(sys.argv) > 2 else 2)
    else:
        main()
