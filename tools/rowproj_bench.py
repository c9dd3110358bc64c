"""Developer tool (GPU box): rowproj_kernel (attention output projection + norm2) against the GEMM + LayerNorm pair it replaces,
at the headline row count."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vla_touch_b200 import native as nv
from vla_touch_b200.plan import Plan, linear_desc, ptr

D, rows = 384, int(os.environ.get("ROWS", 512 * 257))
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
w = (torch.randn(D, D, generator=g) / D ** 0.5).bfloat16()
vec = dict(b=torch.randn(D, generator=g) * 0.1, ls=torch.full((D,), 1e-3), lg=torch.ones(D), lb=torch.zeros(D))


def timed(plan, n_ops):
    prog = plan.compile()
    for _ in range(3):
        prog.run(0, n_ops)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        prog.run(0, n_ops)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3


def build(mode):
    plan = Plan(dev)
    ctx = plan.buf("ctx", (rows, D), torch.bfloat16)
    xn = plan.buf("xn", (rows, D), torch.bfloat16)
    h = plan.buf("h", (rows, D), torch.float32)
    ctx.normal_()
    h.normal_()
    t = {k: plan.reg(v.to(dev).contiguous()) for k, v in dict(w=w, **vec).items()}
    if mode in ("rowproj", "rowproj+ln"):
        d = nv.RowprojDesc()
        d.x, d.ld_x, d.w, d.w_ld, d.bias, d.colscale = ptr(ctx), D, ptr(t["w"]), D, ptr(t["b"]), ptr(t["ls"])
        d.h, d.ld_h, d.rows, d.D = ptr(h), D, rows, D
        if mode == "rowproj+ln":
            d.ln_gamma, d.ln_beta, d.ln_out, d.ln_ld, d.ln_eps = ptr(t["lg"]), ptr(t["lb"]), ptr(xn), D, 1e-6
        plan.add(d, "rowproj")
        return plan, 1
    plan.add(linear_desc(a=ctx, rows=rows, k=D, a_ld=D, w=t["w"], n=D, n_pad=D, w_ld=D, out=h, ldc=D, bias=t["b"],
                         colscale=t["ls"], res=h, ldres=D), "attn_out")
    if mode == "gemm":
        return plan, 1
    d = nv.LnDesc()
    d.x, d.in_ld, d.in_row_stride, d.rows, d.D = ptr(h), D, 1, rows, D
    d.gamma, d.beta, d.eps = ptr(t["lg"]), ptr(t["lb"]), 1e-6
    d.out, d.out_dtype, d.out_ld, d.out_plane, d.act = ptr(xn), nv.VT_BF16, D, 0, nv.ACT_NONE
    plan.add(d, "norm2")
    return plan, 2


for mode in ("gemm", "gemm+ln", "rowproj", "rowproj+ln"):
    print(f"{mode:14s} {timed(*build(mode)):8.1f} us")
if "dbg" in os.environ.get("VT_LIB", ""):     # knock-outs of the LayerNorm warps (debug library)
    for name, k in (("no stores", 1), ("no loads", 2), ("no loads, no stores", 3), ("handshake only", 4)):
        os.environ["VT_MLP_LN_DEBUG"] = str(k)
        print(f"  rowproj+ln, {name:20s} {timed(*build('rowproj+ln')):8.1f} us")
