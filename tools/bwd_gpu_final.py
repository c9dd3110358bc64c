"""One short GPU run of the largest backward cases of tests/bwd_cases.py (no per-op interpretation: seconds, not minutes).
    python tools/bwd_gpu_final.py > gpurun_out/bwd_final.txt"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
t00 = time.time()
import torch  # noqa: E402

import bwd_cases  # noqa: E402

DEV = torch.device("cuda", 0)
cases = [("get_loss.backward A=7 T=64 (BASELINE shape)", lambda: bwd_cases.loss_case(DEV, 7, 64)),
         ("get_loss.backward A=10 T=16 (reference defaults)", lambda: bwd_cases.loss_case(DEV, 10, 16)),
         ("unet backward vs explicit oracle", lambda: bwd_cases.unet_case(DEV))]
cases += [(f"res_block {ci}->{co}", lambda ci=ci, co=co: bwd_cases.res_block_case(DEV, ci, co)) for ci, co in ((1024, 512), (7, 256), (256, 256), (256, 512))]
print(f"import {time.time() - t00:.1f}s", flush=True)
for name, fn in cases:
    t0 = time.time()
    try:
        plan, check = fn()
        prog = plan.compile()
        prog.run()
        torch.cuda.synchronize()
        t1 = time.time()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); prog.run(); ev1.record(); torch.cuda.synchronize()
        print(f"== {name}: {len(plan)} ops, {prog.num_launches()} launches, build+first run {t1 - t0:.1f}s, second run {ev0.elapsed_time(ev1):.2f} ms", flush=True)
        print(f"   PASSED {check()}  ({time.time() - t0:.1f}s)", flush=True)
    except Exception:
        print(f"== {name} FAILED\n" + traceback.format_exc()[-3000:], flush=True)
print(f"total {time.time() - t00:.1f}s")
