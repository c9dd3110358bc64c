"""Developer tool (GPU box, instrumented library): timeline of attn_pp_kernel's CTA 0 -- the first softmax thread of each query
tile and the MMA thread.   python -m vla_touch_b200.build --debug-knobs;  VT_LIB=.../libvt_b200_dbg.so python tools/attn_trace.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("VT_LIB", os.path.join(ROOT, "vla_touch_b200", "lib", "libvt_b200_dbg.so"))
os.environ["VT_ATTN_PP"] = "1"
sys.path.insert(0, ROOT)
import torch

from vla_touch_b200 import native as nv
from vla_touch_b200.plan import Plan, ptr

images, tokens, heads = 512, 257, 6
D = heads * 64
plan = Plan(torch.device("cuda:0"))
qkv = plan.buf("qkv", (images * tokens, 3 * D), torch.bfloat16)
ctx = plan.buf("ctx", (images * tokens, D), torch.bfloat16)
qkv.copy_(torch.randn(images * tokens, 3 * D, generator=torch.Generator().manual_seed(1)) * 1.5)
d = nv.AttnDesc()
d.qkv, d.ctx, d.in_dtype, d.images, d.tokens, d.heads = ptr(qkv), ptr(ctx), nv.VT_BF16, images, tokens, heads
d.ctx_ld, d.ctx_plane = D, 0
plan.add(d, "attention")
prog = plan.compile()
L = nv.lib()
L.vt_debug_timestamps.argtypes = [C.POINTER(C.c_longlong), C.c_int]
L.vt_debug_timestamps.restype = C.c_int
buf = (C.c_longlong * 2048)()
for _ in range(2):
    prog.run(0, 1)
torch.cuda.synchronize()
L.vt_debug_timestamps(buf, 2048)
prog.run(0, 1)
torch.cuda.synchronize()
n = L.vt_debug_timestamps(buf, 2048)
S = n // 4
names = {0: "wait S", 1: "S seen", 2: "max done", 3: "pub c2", 4: "pub c3", 5: "pub c0", 6: "pub c1", 7: "tail dot", 8: "O seen", 9: "read out"}
ev = []
for sl in range(4):
    for k in range(0, S, 2):
        tag, t = buf[sl * S + k], buf[sl * S + k + 1]
        if t:
            ev.append((t, sl, tag))
ev.sort()
t0 = ev[0][0]
print(f"{len(ev)} stamps; columns: tile 0 softmax | tile 1 softmax | MMA thread (20+g: S issued, 30+10g+c: P V chunk c issued)")
last = {0: t0, 1: t0, 2: t0, 3: t0}
for t, sl, tag in ev[: int(sys.argv[1]) if len(sys.argv) > 1 else 260]:
    if sl < 2:
        label = names.get(tag, str(tag))
    elif sl == 2:
        label = f"S issued g{tag - 20}" if tag < 30 else f"PV g{(tag - 30) // 10} c{(tag - 30) % 10}"
    else:
        label = {50: "tail: K seen", 51: "tail: scores", 52: "tail: done", 60: "TMA: K issued (g0)", 61: "TMA: V issued (g0)"}.get(tag, str(tag))
    print(f"{t - t0:8d}  " + "                          " * sl + f"{label:10s} +{t - last[sl]:5d}")
    last[sl] = t
