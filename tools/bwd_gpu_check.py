"""Runs every backward case of tests/bwd_cases.py on the GPU and on the CPU descriptor interpreter op by op (each GPU op
starts from the interpreter's inputs), prints the per-op deviations, then the comparison with the oracle.
    python tools/bwd_gpu_check.py [name substrings ...] > gpurun_out/bwd_check.txt      (e.g. `lstm` for the LSTM cases only)"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

import bwd_cases  # noqa: E402
import gpu_diff  # noqa: E402

CPU, DEV = torch.device("cpu"), torch.device("cuda", 0)
cases = [("dgrad." + k, lambda d, k=k: bwd_cases.dgrad_case(k, d)) for k in ("k5", "down", "up")]
cases += [("wgrad." + k, lambda d, k=k: bwd_cases.wgrad_case(k, d)) for k in ("k5", "down", "up", "k1in")]
cases += [("colsum", bwd_cases.colsum_case),
          ("block", lambda d: bwd_cases.block_case(d, False)),
          ("block.film", lambda d: bwd_cases.block_case(d, True)),
          ("block.film.512ch.T64", lambda d: bwd_cases.block_case(d, True, G=3, B=4, T=64, Ci=256, Co=512, seed=6))]
cases += [(f"res_block.{ci}->{co}", lambda d, ci=ci, co=co: bwd_cases.res_block_case(d, ci, co))
          for ci, co in ((256, 256), (256, 512), (1024, 512), (7, 256))]
cases += [("lstm_layers", bwd_cases.lstm_layers_case), ("lstm_loss.A7", lambda d: bwd_cases.lstm_loss_case(d, 7, 64, 32)),
          ("unet", bwd_cases.unet_case), ("loss.A10", lambda d: bwd_cases.loss_case(d, 10, 16))]
if len(sys.argv) > 1:
    cases = [c for c in cases if any(a in c[0] for a in sys.argv[1:])]
ok = True
for name, fn in cases:
    t0 = time.time()
    try:
        (pc, _), (pg, check) = fn(CPU), fn(DEV)
        rows = gpu_diff.diff_plans(pc, pg, resync=False)
        print(f"== {name}\n" + gpu_diff.format_rows(rows, tol_rel=2e-2), flush=True)
        res = check()
        print(f"   oracle check passed {res if res else ''}  ({time.time() - t0:.1f}s)", flush=True)
    except Exception:
        ok = False
        print(f"== {name} FAILED\n" + traceback.format_exc(), flush=True)
print("ALL OK" if ok else "SOME FAILED")
