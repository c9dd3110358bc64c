"""Developer tool (GPU box): per-op CUDA-event timing of the predict program (warm, each op replayed alone)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from vla_touch_b200 import native as nv
from vla_touch_b200 import synthetic as syn
from vla_touch_b200.bridge_controller import DiffusionController


def main():
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
    name, hidden, heads, layers, hw, T, A, F, steps, batch = wl
    if len(sys.argv) > 2:
        batch = int(sys.argv[2])
    dev = "cuda:0"
    dino_sd, enc_sd, net_sd = bench.synth_weights(hidden, layers, A, F)
    model_args = {'interpolant_type': 'linear', 'gamma_type': '2^0.5*t(t-1)', 'epsilon_type': '1-t', 'prior_policy': 'vla',
                  'beta_max': 0.03, 'sde_type': 'vs', 'action_dim': A, 'obs_dim': 256, 'obs_horizon': 1, 'net_type': 'unet1D_si',
                  'pretrain': False, 'context_frames': 2, 'horizon': T}
    ctl = DiffusionController(state_dim=A, hidden_dim=256, image_model_path=name, diffusion_steps=steps, device=dev,
                              model_args=model_args, use_force=True, force_dim=F, image_state_dict=dino_sd)
    ctl.state_encoder.load_state_dict(enc_sd)
    ctl.diffusion_model.net.load_state_dict(net_sd)
    ctl.diffusion_model.ema = type(ctl.diffusion_model.ema)(ctl.diffusion_model.net.parameters(), decay=0.75)
    ctl.stats = {k: v.to(dev) for k, v in syn.synth_stats(A).items()}
    inp = syn.synth_predict_inputs(batch, T, A, F, hw, 1234)
    ctl.predict(inp["state"].to(dev), inp["vla_actions"].to(dev), inp["images_cam1"][:, None], inp["images_cam2"][:, None],
                inp["forces"].to(dev))
    torch.cuda.synchronize()
    eng = next(iter(ctl._engines.values()))
    prog = eng.plan.compile()
    a0, b1 = eng.predict_range()
    stop = eng.step_ranges[0][1]          # dino + enc + film + first SDE step
    rows = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in list(range(a0, stop)) + [eng.ranges["denormalize"][0]]:
        d = eng.plan.descs[i]
        ts = []
        for rep in range(4):
            flush.zero_()                  # evict L2 so each launch sees HBM-resident operands like in the real sequence
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            prog.run(i, 1)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = sorted(ts)[1]
        fl = bench.gemm_flops(d) if isinstance(d, nv.GemmDesc) else 0.0
        rows.append((i, eng.plan.tags[i], us, fl))
    tot = sum(r[2] for r in rows)
    out = []
    for i, tag, us, fl in rows:
        out.append(f"{i:4d} {tag:46s} {us:9.1f} us {100 * us / tot:5.1f}%  {fl / us / 1e6 if fl else 0:8.1f} TFLOP/s")
    n_steps = eng.n_steps
    step_us = sum(r[2] for r in rows if eng.step_ranges[0][0] <= r[0] < stop)
    front_us = tot - step_us
    out.append(f"front (dino+enc+film+norm+denorm) {front_us:.0f} us, one SDE step {step_us:.0f} us, estimated whole predict "
               f"{front_us + n_steps * step_us:.0f} us")
    txt = "\n".join(out)
    print(txt)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "op_profile.txt"), "w").write(txt + "\n")


if __name__ == "__main__":
    main()
