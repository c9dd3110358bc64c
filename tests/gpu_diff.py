"""TEST INFRASTRUCTURE ONLY -- runs the same plan on the GPU (native kernels) and on the CPU (plan_emu interpreter)
op by op and reports, for every op, how far the GPU output buffer is from the interpreted one."""
import torch

import plan_emu
from vla_touch_b200 import native as nv
from vla_touch_b200.plan import Plan

_OUT_FIELD = {nv.GemmDesc: "out", nv.LnDesc: "out", nv.AttnDesc: "ctx", nv.ImgStatsDesc: "flags", nv.PatchifyDesc: "out",
              nv.ClsDesc: "h", nv.PackDesc: "out", nv.AffineDesc: None, nv.TembedDesc: "out", nv.SdeDesc: "x",
              nv.LstmDesc: "y", nv.QsampleDesc: "xt", nv.SilossDesc: "out", nv.TcolDesc: "out", nv.GnbwdDesc: "draw",
              nv.ColsumDesc: "out", nv.EwiseDesc: "out", nv.SilossBwdDesc: "dvs", nv.LstmTrainDesc: "gates", nv.LstmBwdDesc: "dgates", nv.LnGeluBwdDesc: "dz0", nv.DropmaskDesc: "mask"}


def _buffer_of(plan: Plan, address: int):
    for name, t in plan.bufs.items():
        base = t.data_ptr()
        if base <= address < base + t.numel() * t.element_size():
            return name
    for i, t in enumerate(plan._reg):          # externally owned tensor: address it by registration index
        base = t.data_ptr()
        if base <= address < base + t.numel() * t.element_size():
            return i
    return None


def out_buffer_name(plan: Plan, d) -> str:
    f = _OUT_FIELD[type(d)]
    if f is None:
        f = "out" if d.out else "xpad"
    return _buffer_of(plan, getattr(d, f))


def diff_plans(cpu: Plan, gpu: Plan, first=0, count=-1, sync_inputs=True, report=None, resync=False):
    """Run ops [first, first+count) on both plans; returns a list of (index, tag, buffer, max_abs_diff, ref_max).
    Buffers with identical names must have been given identical contents beforehand (sync_inputs copies cpu -> gpu).
    resync=True overwrites the GPU buffer with the interpreted one after each op, so errors do not propagate and
    every op is checked in isolation."""
    assert len(cpu) == len(gpu)
    if sync_inputs:
        for name, t in cpu.bufs.items():
            gpu.bufs[name].copy_(t)
    prog = gpu.compile()
    last = len(cpu) if count < 0 else first + count
    rows = []
    for i in range(first, last):
        plan_emu.run(cpu, i, 1)
        prog.run(i, 1)
        torch.cuda.synchronize()
        name = out_buffer_name(cpu, cpu.descs[i])
        tc = cpu.bufs[name] if isinstance(name, str) else cpu._reg[name]
        tg = gpu.bufs[name] if isinstance(name, str) else gpu._reg[name]
        ref = tc.float()
        got = tg.float().cpu()
        err = (got - ref).abs()
        err = torch.where(torch.isfinite(err), err, torch.full_like(err, float("inf")))
        row = (i, cpu.tags[i], name, float(err.max()), float(ref.abs().max()))
        rows.append(row)
        if report:
            report(row)
        if resync:
            tg.copy_(tc)
    return rows


def format_rows(rows, tol_rel=None):
    out = []
    for i, tag, name, err, ref in rows:
        flag = ""
        if tol_rel is not None and not (err <= tol_rel * max(ref, 1e-6)):
            flag = "  <-- MISMATCH"
        out.append(f"{i:4d} {tag:48s} {str(name):16s} err {err:10.3e} ref {ref:10.3e}{flag}")
    return "\n".join(out)
