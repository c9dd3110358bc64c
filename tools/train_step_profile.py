"""Developer tool (GPU box): where the time of one bridge training step goes -- CUDA-event and host wall-clock time of each phase
of trainer.DiffusionControllerTrainer.train_step at the BASELINE batch, plus a cProfile of the host side.

    python tools/train_step_profile.py [batch=256] > gpurun_out/train_step_profile.txt
"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256

    class A:
        pass
    cx = bench.Ctx(A())
    wl = bench.WORKLOADS["cfg2"]
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.trainer import DiffusionControllerTrainer
    ctl = bench.make_controller(cx, wl)
    tr = DiffusionControllerTrainer(ctl, syn.synth_stats(wl[6]), device=cx.dev)
    batch = bench.synth_train_batch(cx, wl, B, cx.dev)
    for _ in range(3):
        tr.train_step(dict(batch))
    torch.cuda.synchronize()
    # phases, each bracketed by a synchronize (serialised: the sum over-states the pipelined step)
    dm = ctl.diffusion_model
    names = []

    def phase(name, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        names.append((name, (t1 - t0) * 1e3, (t2 - t0) * 1e3))
        return r

    for it in range(2):
        names.clear()
        bd = phase("prepare_batch (normalise x2, DinoV2 x2, state encoder)", lambda: tr._prepare_batch_for_diffusion(dict(batch)))
        prog = phase("ensure program", lambda: tr._ensure(B, wl[5]))
        phase("optimizer.zero_grad", tr.optimizer.zero_grad)
        phase("sync_train_program (re-pack operands)", lambda: dm.sync_train_program(prog))
        x1, x0 = bd['expert_act'].float(), bd['vla_act'].float()
        phase("set_inputs", lambda: prog.set_inputs(x0, x1, bd['obs_cond'].detach().float().flatten(1), torch.rand(B, device=cx.dev), torch.randn_like(x1)))
        phase("program run (forward + backward, eager)", prog.run)
        phase("d_cond -> state encoder backward", lambda: bd['obs_cond'].backward(prog.d_cond.view_as(bd['obs_cond'])))
        phase("optimizer.step", tr.optimizer.step)
        print(f"--- iteration {it}: phase, host ms, host+device ms")
        for n, a, b in names:
            print(f"{n:62s} {a:9.2f} {b:9.2f}")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        tr.train_step(dict(batch))
    torch.cuda.synchronize()
    print(f"pipelined train_step: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(2):
        tr.train_step(dict(batch))
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr, stream=sys.stdout).sort_stats("cumulative").print_stats(35)


if __name__ == "__main__":
    main()
