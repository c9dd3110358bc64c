"""Developer tool (GPU box): the fused ViT MLP kernel at the headline row count with and without the LayerNorm warps, and
with the LayerNorm warps' loads / stores knocked out (VT_MLP_LN_DEBUG; needs the --debug-knobs library via VT_LIB)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from vla_touch_b200 import native as nv
from vla_touch_b200.plan import Plan, ptr

D, rows = 384, int(os.environ.get("ROWS", 512 * 257))
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
w1 = (torch.randn(4 * D, D, generator=g) / D ** 0.5).bfloat16()
w2 = (torch.randn(D, 4 * D, generator=g) / (4 * D) ** 0.5).bfloat16()
vec = dict(b1=torch.randn(4 * D, generator=g) * 0.1, b2=torch.randn(D, generator=g) * 0.1, ls2=torch.full((D,), 1e-3),
           lg=torch.ones(D), lb=torch.zeros(D))


def run(mode, ln, separate_out):
    if mode is None:
        os.environ.pop("VT_MLP_LN_DEBUG", None)
    else:
        os.environ["VT_MLP_LN_DEBUG"] = str(mode)
    plan = Plan(dev)
    xn = plan.buf("xn", (rows, D), torch.bfloat16)
    h = plan.buf("h", (rows, D), torch.float32)
    lo = plan.buf("lo", (rows, D), torch.bfloat16) if separate_out else xn
    xn.normal_()
    h.normal_()
    t = {k: plan.reg(v.to(dev).contiguous()) for k, v in dict(w1=w1, w2=w2, **vec).items()}
    d = nv.MlpDesc()
    d.xn, d.ld_x, d.w1, d.w1_ld, d.b1 = ptr(xn), D, ptr(t["w1"]), D, ptr(t["b1"])
    d.w2, d.w2_ld, d.b2, d.ls2 = ptr(t["w2"]), 4 * D, ptr(t["b2"]), ptr(t["ls2"])
    d.h, d.ld_h, d.rows, d.D = ptr(h), D, rows, D
    if ln:
        d.ln_gamma, d.ln_beta, d.ln_out, d.ln_ld, d.ln_eps = ptr(t["lg"]), ptr(t["lb"]), ptr(lo), D, 1e-6
    plan.add(d, "mlp")
    prog = plan.compile()
    for _ in range(3):
        prog.run(0, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        prog.run(0, 1)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3


for name, mode, ln, sep in [("no LayerNorm", None, False, False), ("LayerNorm -> xn (in place)", None, True, False),
                            ("LayerNorm -> separate buffer", None, True, True), ("  no stores", 1, True, False),
                            ("  no loads", 2, True, False), ("  no loads, no stores", 3, True, False),
                            ("  handshake only", 4, True, False)]:
    print(f"{name:34s} {run(mode, ln, sep):8.1f} us")
