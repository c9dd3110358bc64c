"""TEST INFRASTRUCTURE ONLY.  Round-2 additions to the golden vectors, generated like oracle/gen_golden.py by running the UNMODIFIED
reference from /root/reference (oracle/ref_shims.py):

    python oracle/gen_golden_extra.py       # writes tests/golden/sde_bs_*.npz and tests/golden/ckpt_manifest.json

  * sde_bs_A7_T64_n10: StochasticInterpolants.sample with sde_type='bs' (bridge_model.py:271-273, 281-332), recorded noise;
  * ckpt_manifest.json: the exact structure of the three checkpoint files the reference writes -- controller.pt, bridge_model.pt
    (bridge_controller.py:203-244, bridge_model.py:435-447) and tactile_controller.pt (lstm_step_controller.py:351-379): every key
    path with shape and dtype in file order, so that the CPU test suite can hold vla_touch_b200's save() to it on a box without
    the reference.  (`ema` comes from the restated torch_ema of ref_shims: the package itself is not installed here.)
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.gen_golden import MODEL_ARGS, NoiseRecorder, OUT, quiet, save  # noqa: E402
from oracle.ref_shims import import_reference  # noqa: E402
from vla_touch_b200 import synthetic as syn  # noqa: E402


def manifest(obj, prefix=""):
    """[(key path, kind, shape, dtype)] of a checkpoint object in iteration order."""
    out = []
    if isinstance(obj, dict):
        for k, v in obj.items():
            out += manifest(v, f"{prefix}{k}/")
    elif isinstance(obj, (list, tuple)):
        out.append([prefix.rstrip("/"), type(obj).__name__, [len(obj)], ""])
        for i, v in enumerate(obj):
            out += manifest(v, f"{prefix}{i}/")
    elif torch.is_tensor(obj):
        out.append([prefix.rstrip("/"), "tensor", list(obj.shape), str(obj.dtype)])
    elif isinstance(obj, np.ndarray):
        out.append([prefix.rstrip("/"), "ndarray", list(obj.shape), str(obj.dtype)])
    else:
        out.append([prefix.rstrip("/"), type(obj).__name__, [], repr(obj) if isinstance(obj, (int, float, str, bool, type(None))) else ""])
    return out


def main():
    torch.set_num_threads(os.cpu_count())
    ref = import_reference(num_dino_layers=1)
    A, Fd, T = 7, 64, 64
    args = dict(MODEL_ARGS, action_dim=A, horizon=T, sde_type='bs')
    with quiet():
        c = ref.bridge_controller.DiffusionController(state_dim=A, hidden_dim=256, image_model_path="facebook/dinov2-small",
                                                      diffusion_steps=10, device="cpu", model_args=args, use_force=True, force_dim=Fd)
    syn.fill_named_(c.diffusion_model.net.named_parameters(), 21, prefix="net.")
    names = [n for n, _ in c.diffusion_model.net.named_parameters()]
    with torch.no_grad():
        for n, s in zip(names, c.diffusion_model.ema.shadow_params):
            s.copy_(syn.synth_param("net." + n, tuple(s.shape), 1021))
    # ---- checkpoint manifests (freshly built controllers: ema.collected_params is still None) ----
    man = {}
    stats = {k: v.numpy() for k, v in syn.synth_stats(A).items()}          # the trainer stores numpy arrays (controller_dataset.py:222-229)
    c.stats = stats
    with tempfile.TemporaryDirectory() as td:
        c.save(td)
        for f in ("controller.pt", "bridge_model.pt"):
            man[f] = manifest(torch.load(os.path.join(td, f), map_location="cpu", weights_only=False))
        with quiet():
            lc = ref.lstm_step_controller.TactileLSTMController(state_dim=A, hidden_dim=256, num_layers=2, dropout=0.1, device="cpu",
                                                                force_dim=Fd)
        lc.stats = {k: torch.as_tensor(v) for k, v in stats.items()}
        lc.save(td)
        man["tactile_controller.pt"] = manifest(torch.load(os.path.join(td, "tactile_controller.pt"), map_location="cpu", weights_only=False))
    x0 = syn.det_uniform("sde.x0", (2, T, A), 23, -1.0, 1.0)
    cond = syn.det_normal("sde.cond", (2, 256), 23)
    torch.manual_seed(321)
    with NoiseRecorder() as nr, torch.no_grad():
        out, traj = c.diffusion_model.sample(x_prior=x0, cond=cond, diffuse_step=10, recod_traj=True)
    save(f"sde_bs_A{A}_T{T}_n10", out=out, noise=torch.stack(nr.draws), x1=traj[1])

    with open(os.path.join(OUT, "ckpt_manifest.json"), "w") as f:
        json.dump({"A": A, "F": Fd, "T": T, "files": man}, f)
    print({k: len(v) for k, v in man.items()})


if __name__ == "__main__":
    main()
