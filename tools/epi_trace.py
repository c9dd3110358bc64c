"""Developer tool (GPU box): VT_GEMM_DEBUG=128 timestamps of one epilogue warp (CTA 0, warp 2) for selected GEMM ops."""
import os, sys, ctypes as C
os.environ["VT_GEMM_DEBUG"] = os.environ.get("VT_TRACE_BITS", "128")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch, ncu_ops
from vla_touch_b200 import native as nv
ctl, eng = ncu_ops.make()
prog = eng.plan.compile()
L = nv.lib()
L.vt_debug_timestamps.argtypes = [C.POINTER(C.c_longlong), C.c_int]
L.vt_debug_timestamps.restype = C.c_int
buf = (C.c_longlong * 2048)()
names = {1: "enter", 2: "acc_full", 3: "tmem_ld", 4: "transpose", 5: "math+store", 6: "exit", 10: "chunk", 11: "h_full", 12: "tmem_ld64", 13: "gelu", 14: "h_sfree", 15: "sts+arrive", 19: "tile", 20: "gn enter", 21: "gn acc_full", 22: "gn pass1", 23: "gn stats", 24: "gn chunk", 30: "kernel entry", 31: "prologue done", 32: "pdl_wait done", 33: "first operands", 34: "tile MMAs issued", 35: "roles done", 36: "final sync"}
for i in [int(x) for x in sys.argv[1:]]:
    prog.run(i, 1); torch.cuda.synchronize()
    L.vt_debug_timestamps(buf, 2048)          # reset
    prog.run(i, 1); torch.cuda.synchronize()
    n = L.vt_debug_timestamps(buf, 2048)
    ev = [(buf[k], buf[k + 1]) for k in range(0, n, 2)]
    print(f"op {i} {eng.plan.tags[i]}: {n // 2} stamps")
    t0 = ev[0][1] if ev else 0
    prev = t0
    lo = int(os.environ.get('VT_TRACE_FROM', '0'))
    for tag, t in ev[lo:lo + 70]:
        print(f"   {names.get(tag, tag):12s} +{t - prev:6d}   (t={t - t0})")
        prev = t
