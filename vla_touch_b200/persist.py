"""Host side of the persistent multi-layer launch (csrc/vt_persist.cuh, include/vt_b200.h vt_persist_desc): turns the op list of
one U-Net evaluation -- the very vt_gemm_desc records the multi-launch form replays one kernel each
(unet.build_unet_eval; DiffusionConditionalUnet1D.forward, conditional_unet_1D.py:194-247) -- plus, in the sampler, the
Euler-Maruyama update that closes each step (bridge_model.py:343-385) into ONE descriptor = one kernel launch for all steps.

The only thing the kernel needs beyond the layers themselves is who produces what: `analyse` derives it from the descriptors
(buffer + channel window of every layer's A operand, residual and output), and checks the two properties the kernel's counters
rely on: (1) every read-after-write is an explicit dependency, (2) every write-after-read / write-after-write on a re-used
buffer is ordered by a chain of those dependencies, within a step and across consecutive steps.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import native as nv


class Collector:
    """Stands in for a Plan while a builder lists the ops of one evaluation: only `add` and `len` are used by unet._conv."""

    def __init__(self) -> None:
        self.descs: List[C.Structure] = []
        self.tags: List[str] = []

    def add(self, desc, tag: str = "") -> int:
        self.descs.append(desc)
        self.tags.append(tag)
        return len(self.descs) - 1

    def __len__(self) -> int:
        return len(self.descs)


def _region(tensors: Sequence[torch.Tensor], address: int, elem: int):
    """(index of the registered tensor holding `address`, element offset inside it)"""
    for i, t in enumerate(tensors):
        base = t.data_ptr()
        if base <= address < base + t.numel() * t.element_size():
            return i, (address - base) // elem
    raise KeyError(f"address {address:#x} is not inside a registered plan tensor")


def _windows(tensors, d: nv.GemmDesc):
    """Channel windows a layer touches: A operand, residual, output as (buffer, first channel, end channel, row key)."""
    es_in = 2 if d.in_dtype == nv.VT_BF16 else 4
    es_out = 2 if d.out_dtype == nv.VT_BF16 else 4
    buf, off = _region(tensors, d.a, es_in)
    ca = off % d.a_ld + d.a_c0
    a_win = (buf, ca, ca + d.kc, None)
    r_win = None
    if d.res:
        es_res = es_out if d.epi == nv.EPI_GN else 4
        rbuf, roff = _region(tensors, d.res, es_res)
        cr = roff % d.ldres
        r_win = (rbuf, cr, cr + d.N, None)
    obuf, ooff = _region(tensors, d.out, es_out)
    co = ooff % d.ldc
    o_win = (obuf, co, co + d.N, (d.row_div, d.out_q, d.out_r, d.out_off + ooff // d.ldc))
    return a_win, r_win, o_win


def _overlap(a, b) -> bool:
    return a[0] == b[0] and a[1] < b[2] and b[1] < a[2]


def analyse(tensors: Sequence[torch.Tensor], gemms: Sequence[nv.GemmDesc], sde: Optional[nv.SdeDesc]):
    """-> (deps, lags): per layer (the Euler-Maruyama update is layer len(gemms)) up to PERSIST_MAX_DEPS producers.
    lag 1 = the producer's output of the previous step (the update's x feeding the next step's first layers)."""
    n = len(gemms)
    wins = [_windows(tensors, d) for d in gemms]
    SDE = n
    writers: List[Tuple[tuple, int, int]] = []          # (output window, layer, lag)
    if sde is not None and sde.xpad:
        es = 2 if sde.xpad_dtype == nv.VT_BF16 else 4
        xb, xo = _region(tensors, sde.xpad, es)
        writers.append(((xb, xo % sde.xpad_ld, xo % sde.xpad_ld + sde.A, None), SDE, 1))
    deps: List[List[Tuple[int, int]]] = []
    readers: List[Tuple[tuple, int]] = []               # (window read, layer) since the window's last write
    for j, (a_win, r_win, o_win) in enumerate(wins):
        dj: List[Tuple[int, int]] = []
        for win in (a_win, r_win):
            if win is None:
                continue
            for w, k, lag in writers:
                if _overlap(w, win) and (k, lag) not in dj:
                    dj.append((k, lag))
            readers.append((win, j))
        deps.append(dj)
        # this layer's output supersedes earlier writers of exactly the same rows and channels
        writers = [(w, k, lag) for (w, k, lag) in writers if not (w[0] == o_win[0] and w[1] >= o_win[1] and w[2] <= o_win[2]
                                                                 and (w[3] is None or w[3] == o_win[3]))]
        writers.append((o_win, j, 0))
    if sde is not None:
        vb, vo = _region(tensors, sde.v, 4)
        dj = [(k, lag) for (w, k, lag) in writers if w[0] == vb and lag == 0]
        if not dj:
            raise ValueError("persist: no layer writes the Euler-Maruyama update's v / s input")
        deps.append(dj)
    for j, dj in enumerate(deps):
        if len(dj) > nv.PERSIST_MAX_DEPS:
            raise ValueError(f"persist: layer {j} has {len(dj)} producers (max {nv.PERSIST_MAX_DEPS})")
    _check_hazards(wins, deps, n, sde is not None)
    return deps


def _check_hazards(wins, deps, n: int, has_sde: bool) -> None:
    """Every layer that overwrites (part of) a buffer must be ordered after all earlier readers and writers of that region by a
    chain of dependencies (inside one step); and every layer must be an ancestor of the step's last op, so that the same holds
    across consecutive steps (each step's first layers wait for the previous step's last op)."""
    anc: List[set] = []
    for j in range(len(deps)):
        s = set()
        for k, lag in deps[j]:
            if lag == 0:
                s.add(k)
                s |= anc[k]
        anc.append(s)
    for j in range(n):
        o = wins[j][2]
        for i in range(j):
            a_i, r_i, o_i = wins[i]
            touches = any(w is not None and _overlap(w, o) for w in (a_i, r_i)) or (_overlap(o_i, o) and (o_i[3] == o[3] or o_i[3] is None or o[3] is None))
            if touches and i not in anc[j]:
                raise ValueError(f"persist: layer {j} overwrites a region layer {i} reads or writes without being ordered after it")
    last = len(deps) - 1
    missing = [i for i in range(last) if i not in anc[last]]
    if missing:
        raise ValueError(f"persist: layers {missing} do not feed the step's last op: consecutive steps would race on their buffers")
    if has_sde and not any(lag == 1 for dj in deps for _, lag in dj):
        raise ValueError("persist: no layer reads the Euler-Maruyama update's output")


def make_desc(tensors: Sequence[torch.Tensor], gemms: Sequence[nv.GemmDesc], *, sde: Optional[nv.SdeDesc] = None, n_steps: int = 1,
              film_t_step: int = 0, coef: Optional[Sequence[Sequence[float]]] = None, noise_step: int = 0, sde_T: int = 0,
              tags: Optional[Sequence[str]] = None) -> nv.PersistDesc:
    """The vt_persist_desc of `gemms` (+ `sde`) repeated `n_steps` times.  coef: per step (ginv, dgg, eps, dt, nscale)."""
    n = len(gemms)
    if n > nv.PERSIST_MAX_LAYERS:
        raise ValueError(f"persist: {n} layers (max {nv.PERSIST_MAX_LAYERS})")
    deps = analyse(tensors, gemms, sde)
    rows = n + (1 if sde is not None else 0)
    dep_arr = (C.c_int32 * (rows * nv.PERSIST_MAX_DEPS))(*([-1] * (rows * nv.PERSIST_MAX_DEPS)))
    lag_arr = (C.c_int32 * (rows * nv.PERSIST_MAX_DEPS))()
    for j, dj in enumerate(deps):
        for k, (layer, lag) in enumerate(dj):
            dep_arr[j * nv.PERSIST_MAX_DEPS + k] = layer
            lag_arr[j * nv.PERSIST_MAX_DEPS + k] = lag
    garr = (nv.GemmDesc * n)(*gemms)
    d = nv.PersistDesc()
    d.gemms, d.n_gemms = C.cast(garr, C.c_void_p), n
    d.deps, d.dep_lag = C.cast(dep_arr, C.c_void_p), C.cast(lag_arr, C.c_void_p)
    d.n_steps, d.film_t_step, d.noise_step, d.sde_T = n_steps, film_t_step, noise_step, sde_T
    keep = [garr, dep_arr, lag_arr]
    if sde is not None:
        if coef is None or len(coef) != n_steps:
            raise ValueError("persist: one (ginv, dgg, eps, dt, nscale) row per step is required with the Euler-Maruyama update")
        carr = (C.c_float * (5 * n_steps))(*[float(x) for row in coef for x in row])
        d.sde, d.sde_coef = C.cast(C.pointer(sde), C.c_void_p), C.cast(carr, C.c_void_p)
        keep += [sde, carr]
    d._keep = keep
    d.layers, d.py_sde, d.coef, d.py_deps, d.tags = list(gemms), sde, coef, deps, list(tags or [])
    d.algo_flops = float(n_steps) * sum(2.0 * g.G * g.M * g.N * g.taps * g.kc for g in gemms)
    return d
