"""Developer tool (GPU box): time selected GEMM ops with parts of the kernel disabled (VT_GEMM_DEBUG)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
    import torch, ncu_ops
    ctl, eng = ncu_ops.make()
    prog = eng.plan.compile()
    for i in [int(x) for x in sys.argv[2:]]:
        ts = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); prog.run(i, 1); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"  op {i:4d} {eng.plan.tags[i]:44s} {sorted(ts)[1]:8.1f} us", flush=True)
else:
    ops = sys.argv[1:] or ["7", "9", "11", "12", "105", "111", "115", "123"]
    for dbg in os.environ.get("VT_DEBUG_SET", "0 1 2 3 6 10").split():
        print(f"VT_GEMM_DEBUG={dbg}  (1: no TMA+MMA, 2: no epilogue, 4: no TMA, 8: no MMA)", flush=True)
        env = dict(os.environ, VT_GEMM_DEBUG=dbg)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"] + ops, env=env, capture_output=True, text=True, timeout=600)
        print(r.stdout + r.stderr[-800:], flush=True)
