"""Developer tool (GPU box): replay selected ops of the predict program once each inside a cudaProfiler range, for
   ncu --profile-from-start off --set full ... python tools/ncu_ops.py <op index | tag substring> [...]
   `step` replays the whole predict program once (eager, op by op) instead: the launch list of one step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch

import bench
from vla_touch_b200 import synthetic as syn
from vla_touch_b200.bridge_controller import DiffusionController


def make(workload="cfg2", batch=None):
    name, hidden, heads, layers, hw, T, A, F, steps, b = bench.WORKLOADS[workload]
    batch = batch or b
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
    return ctl, next(iter(ctl._engines.values()))


if __name__ == "__main__":
    ctl, eng = make()
    prog = eng.plan.compile()
    if sys.argv[1:] == ["step"]:
        a0, b1 = eng.predict_range()
        ops = list(range(a0, b1))
    else:
        ops = []
        for x in sys.argv[1:]:
            ops.append(int(x) if x.isdigit() else next(i for i, t in enumerate(eng.plan.tags) if x in t))
    for i in ops:
        prog.run(i, 1)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for i in ops:
        print(i, eng.plan.tags[i])
        prog.run(i, 1)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
