"""Developer tool (GPU box): the 257-token attention op alone (cfg2 shape: 512 images x 6 heads), attn_pp_kernel against
attn_row_kernel (VT_ATTN_PP=0), CUDA-event time per launch, max |difference| between the two, tensor TFLOP/s."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from vla_touch_b200 import native as nv
from vla_touch_b200.plan import Plan, ptr


def run(images, tokens, heads, pp, qkv32, reps=20):
    os.environ["VT_ATTN_PP"] = "1" if pp else "0"
    D = heads * 64
    plan = Plan(torch.device("cuda:0"))
    qkv = plan.buf("qkv", (images * tokens, 3 * D), torch.bfloat16)
    ctx = plan.buf("ctx", (images * tokens, D), torch.bfloat16)
    qkv.copy_(qkv32)
    d = nv.AttnDesc()
    d.qkv, d.ctx, d.in_dtype, d.images, d.tokens, d.heads = ptr(qkv), ptr(ctx), nv.VT_BF16, images, tokens, heads
    d.ctx_ld, d.ctx_plane = D, 0
    plan.add(d, "attention")
    prog = plan.compile()
    for _ in range(3):
        prog.run(0, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        prog.run(0, 1)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, ctx.float().clone()


if __name__ == "__main__":
    images = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    tokens, heads = 257, 6
    g = torch.Generator().manual_seed(1)
    qkv32 = torch.randn(images * tokens, 3 * heads * 64, generator=g) * 1.5
    flops = 4.0 * tokens * tokens * 64 * heads * images
    t_row, o_row = run(images, tokens, heads, False, qkv32)
    t_pp, o_pp = run(images, tokens, heads, True, qkv32)
    print(f"attn_row_kernel {t_row:8.1f} us  {flops / t_row / 1e6:7.1f} TFLOP/s")
    print(f"attn_pp_kernel  {t_pp:8.1f} us  {flops / t_pp / 1e6:7.1f} TFLOP/s   max |pp - row| = {(o_pp - o_row).abs().max().item():.4g}")
