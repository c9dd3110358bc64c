"""TEST INFRASTRUCTURE ONLY.  Gradient fixtures for the training rows a10/a11 of SURVEY.md 8 (the parity gate of the backward
kernels that round 2 builds): runs the UNMODIFIED reference `StochasticInterpolants.get_loss` (bridge_model.py:220-246) +
`loss.backward()` on CPU fp32 with the deterministic synthetic weights / inputs of oracle/gen_golden.py and the RNG draws
recorded there (tests/golden/loss_A*_T*.npz: `step`, `z_unit`), and writes tests/golden/loss_grads_A*_T*.npz:

  * d loss / d obs_cond in full ([B, 256]: what flows back into the state encoder, bridge_train.py:315-330),
  * per parameter tensor of net.{b,v,s}_net (438 tensors, in `named_parameters()` order): L2 norm, sum, and the first 8
    elements of the flattened gradient (the full gradients are 412 MB and are not stored).

and, for row a12 (lstm_step_controller.py:321-337), tests/golden/lstm_grads_A*_F*_T*.npz: the MSE loss of `get_loss` in eval mode
(dropout off: the train-mode masks are an RNG stream of their own) and the full gradients of the force encoder, the LSTM and
the output head (1.4 M parameters, stored as norm / sum / head digests like above) plus d loss / d obs_cond.

    python oracle/gen_golden_grads.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.gen_golden import MODEL_ARGS, OUT, quiet  # noqa: E402
from oracle.ref_shims import import_reference  # noqa: E402
from vla_touch_b200 import synthetic as syn  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ref = import_reference()
    for A, Fdim, T in ((10, 3, 16), (7, 64, 64)):
        ref.set_dino_layers(1)
        args = dict(MODEL_ARGS, action_dim=A, horizon=T)
        with quiet():
            c = ref.bridge_controller.DiffusionController(state_dim=A, hidden_dim=256, image_model_path="facebook/dinov2-small",
                                                          diffusion_steps=10, device="cpu", model_args=args, use_force=True,
                                                          force_dim=Fdim)
        syn.fill_named_(c.diffusion_model.net.named_parameters(), 21, prefix="net.")
        g = np.load(os.path.join(OUT, f"loss_A{A}_T{T}.npz"))
        B = 3
        cond = syn.det_normal("loss.cond", (B, 256), 24).requires_grad_(True)
        batch = {"obs_cond": cond, "expert_act": syn.det_uniform("loss.exp", (B, T, A), 24, -1.0, 1.0),
                 "vla_act": syn.det_uniform("loss.vla", (B, T, A), 24, -1.0, 1.0)}
        step, z_unit = torch.from_numpy(g["step"]), torch.from_numpy(g["z_unit"])
        rand_orig, randn_like_orig = torch.rand, torch.randn_like
        torch.rand = lambda *a, **k: step.clone()                # the one torch.rand(B) draw of get_loss (:236)
        torch.randn_like = lambda x, *a, **k: z_unit.clone()     # the one randn_like(x0) draw of q_sample (:105)
        try:
            c.diffusion_model.net.train()
            loss, info = c.diffusion_model.get_loss(batch, "cpu")
        finally:
            torch.rand, torch.randn_like = rand_orig, randn_like_orig
        assert abs(float(loss) - float(g["loss"])) <= 1e-5 * max(1.0, abs(float(g["loss"]))), (float(loss), float(g["loss"]))
        loss.backward()
        names, norms, sums, heads = [], [], [], []
        for n, p in c.diffusion_model.net.named_parameters():
            gr = p.grad.detach().flatten().double()
            names.append(n)
            norms.append(float(gr.norm()))
            sums.append(float(gr.sum()))
            h = torch.zeros(8, dtype=torch.float64)
            h[: min(8, gr.numel())] = gr[:8]
            heads.append(h.numpy())
        np.savez(os.path.join(OUT, f"loss_grads_A{A}_T{T}.npz"), loss=np.float32(float(loss)), d_cond=cond.grad.numpy(),
                 names=np.array(names), norm=np.array(norms), sum=np.array(sums), head=np.stack(heads))
        print(f"wrote loss_grads_A{A}_T{T}: {len(names)} tensors, |d_cond| {float(cond.grad.norm()):.4e}, "
              f"total grad norm {float(np.sqrt((np.array(norms) ** 2).sum())):.4e}")


def lstm_grads():
    ref = import_reference()
    cd = ref.controller_dataset
    for A, Fdim, T in ((10, 3, 16), (7, 64, 32)):
        ref.set_dino_layers(1)
        with quiet():
            lc = ref.lstm_step_controller.TactileLSTMController(state_dim=A, hidden_dim=256, num_layers=2, dropout=0.1,
                                                                device="cpu", force_dim=Fdim)
        mods = (("obs_encoder", lc.obs_encoder), ("force_encoder", lc.force_encoder), ("lstm", lc.lstm), ("output_head", lc.output_head))
        for nm, mod in mods:
            syn.fill_named_(mod.named_parameters(), 41, prefix=f"lstm.{nm}.")
        lc.eval()
        lc.stats = syn.synth_stats_varied(A, 41)
        B = 3
        vla = syn.det_uniform("lstm.vla", (B, T, A), 41, -1.0, 1.0)
        forces = syn.det_normal("lstm.forces", (B, T, Fdim), 41)
        cond = syn.det_normal("lstm.cond", (B, 256), 41).requires_grad_(True)
        expert = syn.det_uniform("lstm.exp", (B, T, A), 41, -1.0, 1.0)
        vla_n = cd.normalize_actions(vla, lc.stats, 'vla')
        loss = lc.get_loss({"vla_act": vla_n, "obs_cond": cond, "forces": forces, "expert_act": expert})
        loss.backward()
        names, norms, sums, heads = [], [], [], []
        for nm, mod in mods[1:]:
            for n, p in mod.named_parameters():
                gr = p.grad.detach().flatten().double()
                names.append(f"{nm}.{n}")
                norms.append(float(gr.norm()))
                sums.append(float(gr.sum()))
                h = torch.zeros(8, dtype=torch.float64)
                h[: min(8, gr.numel())] = gr[:8]
                heads.append(h.numpy())
        np.savez(os.path.join(OUT, f"lstm_grads_A{A}_F{Fdim}_T{T}.npz"), loss=np.float32(float(loss.detach())),
                 d_cond=(cond.grad.numpy() if cond.grad is not None else np.zeros((B, 256), np.float32)),
                 names=np.array(names), norm=np.array(norms), sum=np.array(sums), head=np.stack(heads))
        print(f"wrote lstm_grads_A{A}_F{Fdim}_T{T}: {len(names)} tensors, loss {float(loss.detach()):.6f}")


if __name__ == "__main__":
    main()
    lstm_grads()
