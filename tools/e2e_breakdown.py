"""Developer tool (GPU box): where the end-to-end predict() step loses time against the bare graph replay (cfg2, batch 256)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench


def main():
    cx = bench.Ctx(None)
    wl = bench.WORKLOADS["cfg2"]
    batch = wl[-1]
    ctl, eng, api_step, h2d, d2h = bench.predict_harness(cx, wl, batch)
    from vla_touch_b200 import synthetic as syn
    name, hidden, heads, layers, hw, T, A, F, steps, _ = wl
    inp = syn.synth_predict_inputs(batch, T, A, F, hw, 1234)
    dev = {k: v.cuda() for k, v in inp.items()}
    dev["images_cam1"] = inp["images_cam1"][:, None].contiguous().cuda()
    dev["images_cam2"] = inp["images_cam2"][:, None].contiguous().cuda()
    out_host = torch.empty(batch, T, A, dtype=torch.float32).pin_memory()

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def graph_only():
        eng.run_predict(graph=True)

    def graph_d2h():
        eng.run_predict(graph=True)
        out_host.copy_(eng.out.clone(), non_blocking=True)

    def predict_dev():
        out = ctl.predict(dev["state"], dev["vla_actions"], dev["images_cam1"], dev["images_cam2"], dev["forces"])
        out_host.copy_(out, non_blocking=True)

    side = torch.cuda.Stream()
    host_imgs = [inp["images_cam1"][:, None].contiguous().pin_memory(), inp["images_cam2"][:, None].contiguous().pin_memory()]
    land = [torch.empty_like(h, device="cuda") for h in host_imgs]

    def graph_plus_free_upload():          # the same bytes uploaded on a side stream with NO dependency on the graph: pure interference
        with torch.cuda.stream(side):
            for l, h in zip(land, host_imgs):
                l.copy_(h, non_blocking=True)
        eng.run_predict(graph=True)

    print(f"graph replay only                       {timed(graph_only):8.3f} ms")
    print(f"graph + independent 77 MB upload        {timed(graph_plus_free_upload):8.3f} ms")
    torch.cuda.synchronize()
    print(f"graph + clone + D2H of the result       {timed(graph_d2h):8.3f} ms")
    print(f"predict(), device-resident inputs + D2H {timed(predict_dev):8.3f} ms")
    print(f"predict(), pinned host inputs + D2H     {timed(api_step):8.3f} ms   (= e2e)")
    print(f"host->device copies alone               {timed(api_step.h2d_only):8.3f} ms")
    # host time of one predict() call (launch overhead)
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        api_step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host time per api_step() call           {(t1 - t0) / 20 * 1e3:8.3f} ms (asynchronous launches)")


if __name__ == "__main__":
    main()
