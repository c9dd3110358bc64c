"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference
modules from /root/reference (via oracle/ref_shims.py) on CPU fp32 with deterministic synthetic
weights (vla_touch_b200/synthetic.py).  Run once in the build container:

    python oracle/gen_golden.py            # writes tests/golden/*.npz

Inputs and weights are NOT stored (they are pure functions of (name, shape, seed)); only the
reference's outputs and the sampler noise the reference drew are.  tests/ rebuild the same inputs,
and compare (a) the oracle restatement and (b) the CUDA path against these outputs.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shims import import_reference  # noqa: E402
from vla_touch_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
MODEL_ARGS = {
    'interpolant_type': 'linear', 'gamma_type': '2^0.5*t(t-1)', 'epsilon_type': '1-t', 'prior_policy': 'vla',
    'beta_max': 0.03, 'sde_type': 'vs', 'action_dim': 10, 'obs_dim': 256, 'obs_horizon': 1,
    'net_type': 'unet1D_si', 'pretrain': False, 'context_frames': 2, 'horizon': 16,
}


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


class NoiseRecorder:
    """Records every torch.randn_like draw (bridge_model.py:105,372) in call order."""

    def __init__(self):
        self.draws = []
        self._orig = torch.randn_like

    def __enter__(self):
        def rec(x, *a, **k):
            z = self._orig(x, *a, **k)
            self.draws.append(z.clone())
            return z
        torch.randn_like = rec
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    np.savez(os.path.join(OUT, name + ".npz"), **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                                   for k, v in arrs.items()})
    print("wrote", name, {k: tuple(np.asarray(v.detach() if torch.is_tensor(v) else v).shape) for k, v in arrs.items()})


def fill_dino(enc, seed=0):
    syn.fill_named_(enc.model.named_parameters(), seed, prefix="dino.")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ref = import_reference()
    cd = ref.controller_dataset

    # ---------------- (vii) normalise / denormalise incl. degenerate range ----------------
    for A in (7, 10):
        st = syn.synth_stats_varied(A, seed=3)
        x = syn.det_uniform("norm.x", (3, 16, A), 3, -2.0, 2.0)
        save(f"norm_A{A}",
             vla_n=cd.normalize_actions(x, st, 'vla'), exp_n=cd.normalize_actions(x, st, 'expert'),
             exp_dn=cd.denormalize_actions(x, st, 'expert'))

    # ---------------- (i) DinoV2 CLS features ----------------
    def dino_case(tag, model, layers, hw, batch, kind, seed):
        ref.set_dino_layers(layers)
        with quiet():
            enc = ref.visual_encoder.DINOv2Encoder(model_name=model, device="cpu")
        fill_dino(enc, seed)
        u8 = syn.synth_images_u8("dino.img", batch, hw, seed, dark=(kind == "u8dark"))
        if kind in ("u8bright5d", "u8dark"):
            img = u8[:, None]                                   # [B,1,H,W,3] uint8 (deployment)
        elif kind == "f32bhwc":
            img = u8.float() / 255.0                            # [B,H,W,3] f32 in [0,1] (train/test)
        elif kind == "f32bchw":
            img = (u8.float() / 255.0).permute(0, 3, 1, 2).contiguous()   # [B,3,H,W] (smoke test)
        elif kind == "u8bhwc":
            img = u8
        else:
            raise ValueError(kind)
        extras = {}
        hooks = []
        if layers <= 2:
            hs = {}
            hooks.append(enc.model.embeddings.register_forward_hook(lambda m, i, o: hs.__setitem__("h_emb", o)))
            for li, lyr in enumerate(enc.model.encoder.layer):
                hooks.append(lyr.register_forward_hook(
                    lambda m, i, o, li=li: hs.__setitem__(f"h_l{li}", o[0] if isinstance(o, tuple) else o)))
        out = enc.forward(img)
        for h in hooks:
            h.remove()
        if layers <= 2:
            n_tok = next(iter(hs.values())).shape[1]
            idx = list(range(8)) + list(range(n_tok - 4, n_tok))  # first image, 12 tokens (fixture size)
            extras = {k: v[:1, idx] for k, v in hs.items()}
        save(tag, cls=out, **extras)

    dino_case("dino_s12_224_u8bright5d", "facebook/dinov2-small", 12, 224, 2, "u8bright5d", 11)
    dino_case("dino_s2_224_u8dark", "facebook/dinov2-small", 2, 224, 2, "u8dark", 12)
    dino_case("dino_s2_224_f32bhwc", "facebook/dinov2-small", 2, 224, 2, "f32bhwc", 13)
    dino_case("dino_s2_224_f32bchw", "facebook/dinov2-small", 2, 224, 2, "f32bchw", 14)
    dino_case("dino_s2_224_u8bhwc", "facebook/dinov2-small", 2, 224, 3, "u8bhwc", 15)
    dino_case("dino_s2_384_u8bright5d", "facebook/dinov2-small", 2, 384, 1, "u8bright5d", 16)
    dino_case("dino_b2_224_u8bright5d", "facebook/dinov2-base", 2, 224, 1, "u8bright5d", 17)

    # ---------------- controller-level fixtures ----------------
    def make_controller(A, Fdim, layers, model="facebook/dinov2-small", seed=0, horizon=16):
        ref.set_dino_layers(layers)
        args = dict(MODEL_ARGS, action_dim=A, horizon=horizon)
        with quiet():
            c = ref.bridge_controller.DiffusionController(
                state_dim=A, hidden_dim=256, image_model_path=model, diffusion_steps=10, device="cpu",
                model_args=args, use_force=True, force_dim=Fdim)
        fill_dino(c.image_encoder, seed)
        syn.fill_named_(c.state_encoder.named_parameters(), seed, prefix="enc.")
        syn.fill_named_(c.diffusion_model.net.named_parameters(), seed, prefix="net.")
        # EMA shadow != live params, so that tests see sample() really uses the EMA weights (:267)
        names = [n for n, _ in c.diffusion_model.net.named_parameters()]
        with torch.no_grad():
            for n, s in zip(names, c.diffusion_model.ema.shadow_params):
                s.copy_(syn.synth_param("net." + n, tuple(s.shape), seed + 1000))
        return c

    # (ii) state_encoder and (iii) single U-Net evals
    for A, Fdim in ((10, 3), (7, 64)):
        c = make_controller(A, Fdim, layers=1, seed=21)
        obs_dim = 2 * 384 + A + Fdim
        xin = syn.det_normal("enc.in", (4, obs_dim), 21)
        save(f"enc_A{A}_F{Fdim}", out=c.state_encoder(xin))
        for T in (16, 32, 48, 64):
            B = 3
            x = syn.det_uniform("unet.x", (B, T, A), 22, -1.0, 1.0)
            cond = syn.det_normal("unet.cond", (B, 256), 22)
            t = torch.tensor([0.3, 0.001, 0.999])
            with torch.no_grad():
                v = c.diffusion_model.net.v_net(x, t, global_cond=cond)
                s = c.diffusion_model.net.s_net(x, t[:1].expand(B), global_cond=cond)
            save(f"unet_A{A}_T{T}", v=v, s=s)

        # (iv) full sde_vs with recorded noise, EMA weights, n in {10, 50 (A=7 only)}
        for n in ((10, 50) if A == 7 else (10,)):
            T, B = (64, 2) if A == 7 else (16, 2)
            x0 = syn.det_uniform("sde.x0", (B, T, A), 23, -1.0, 1.0)
            cond = syn.det_normal("sde.cond", (B, 256), 23)
            torch.manual_seed(100 + n)
            with NoiseRecorder() as nr, torch.no_grad():
                out, traj = c.diffusion_model.sample(x_prior=x0, cond=cond, diffuse_step=n, recod_traj=True)
            save(f"sde_A{A}_T{T}_n{n}", out=out, noise=torch.stack(nr.draws), x1=traj[1])
        # beta_max = 0 (deterministic) variant
        c.diffusion_model.d = 0.0
        T, B = (64, 2) if A == 7 else (16, 2)
        x0 = syn.det_uniform("sde.x0", (B, T, A), 23, -1.0, 1.0)
        cond = syn.det_normal("sde.cond", (B, 256), 23)
        with torch.no_grad():
            out = c.diffusion_model.sample(x_prior=x0, cond=cond, diffuse_step=10)
        save(f"sde_A{A}_T{T}_n10_beta0", out=out)
        c.diffusion_model.d = 0.03

        # (v) get_loss with recorded t and z
        T, B = (64, 3) if A == 7 else (16, 3)
        batch = {"obs_cond": syn.det_normal("loss.cond", (B, 256), 24),
                 "expert_act": syn.det_uniform("loss.exp", (B, T, A), 24, -1.0, 1.0),
                 "vla_act": syn.det_uniform("loss.vla", (B, T, A), 24, -1.0, 1.0)}
        torch.manual_seed(77)
        rand_orig = torch.rand
        steps = []

        def rec_rand(*a, **k):
            r = rand_orig(*a, **k)
            steps.append(r.clone())
            return r
        torch.rand = rec_rand
        try:
            with NoiseRecorder() as nr:
                loss, info = c.diffusion_model.get_loss(batch, "cpu")
        finally:
            torch.rand = rand_orig
        save(f"loss_A{A}_T{T}", loss=loss, v_loss=info['v_loss'], s_loss=info['s_loss'], b_loss=info['b_loss'],
             step=steps[0], z_unit=nr.draws[0])

    # ---------------- (full) predict: BASELINE config 1 + small config-2 shape ----------------
    def predict_case(tag, A, Fdim, T, hw, B, layers, model, seed, steps=10, dark=False, kind="u8_5d"):
        c = make_controller(A, Fdim, layers, model, seed, horizon=T)
        c.stats = syn.synth_stats_varied(A, seed) if tag.endswith("varstats") else syn.synth_stats(A)
        c.diffusion_steps = steps
        inp = syn.synth_predict_inputs(B, T, A, Fdim, hw, seed, dark)
        i1, i2 = inp["images_cam1"], inp["images_cam2"]
        if kind == "u8_5d":
            i1, i2 = i1[:, None], i2[:, None]
        elif kind == "f32_bhwc":
            i1, i2 = i1.float() / 255.0, i2.float() / 255.0
        torch.manual_seed(seed)
        with NoiseRecorder() as nr, quiet():
            out = c.predict(inp["state"], inp["vla_actions"], i1, i2, inp["forces"])
            cond = c.encode_observation(inp["state"], i1, i2, inp["forces"])
        save(tag, out=out, noise=torch.stack(nr.draws), cond=cond)

    # BASELINE.json configs[0]: batch 1, 10 steps, fp32, reference defaults (T16 A10 F3 384^2 small)
    predict_case("predict_cfg1", 10, 3, 16, 384, 1, 12, "facebook/dinov2-small", 31)
    # configs[1] shapes at a CPU-affordable batch
    predict_case("predict_cfg2_B2", 7, 64, 64, 224, 2, 12, "facebook/dinov2-small", 32)
    predict_case("predict_cfg2_B3_dark_varstats", 7, 64, 64, 224, 3, 2, "facebook/dinov2-small", 33, dark=True)
    predict_case("predict_T48_f32_varstats", 10, 3, 48, 224, 2, 2, "facebook/dinov2-small", 34, kind="f32_bhwc")
    # configs[2] model family (DinoV2-B/14, 50 steps) at B=1, 2 layers
    predict_case("predict_cfg3_B1_base", 7, 64, 64, 224, 1, 2, "facebook/dinov2-base", 35, steps=50)

    # ---------------- (vi) LSTM controller ----------------
    for A, Fdim, T in ((10, 3, 16), (7, 64, 32)):
        ref.set_dino_layers(1)
        with quiet():
            lc = ref.lstm_step_controller.TactileLSTMController(state_dim=A, hidden_dim=256, num_layers=2, dropout=0.1,
                                                                device="cpu", force_dim=Fdim)
        for nm, mod in (("obs_encoder", lc.obs_encoder), ("force_encoder", lc.force_encoder), ("lstm", lc.lstm),
                        ("output_head", lc.output_head)):
            syn.fill_named_(mod.named_parameters(), 41, prefix=f"lstm.{nm}.")
        lc.eval()
        lc.stats = syn.synth_stats_varied(A, 41)
        B = 3
        vla = syn.det_uniform("lstm.vla", (B, T, A), 41, -1.0, 1.0)
        forces = syn.det_normal("lstm.forces", (B, T, Fdim), 41)
        cond = syn.det_normal("lstm.cond", (B, 256), 41)
        expert = syn.det_uniform("lstm.exp", (B, T, A), 41, -1.0, 1.0)
        with torch.no_grad():
            vla_n = cd.normalize_actions(vla, lc.stats, 'vla')
            fwd = lc.forward({"vla_act": vla_n, "obs_cond": cond, "forces": forces})
            loss = lc.get_loss({"vla_act": vla_n, "obs_cond": cond, "forces": forces, "expert_act": expert})
            seq = lc.predict_sequence(cond, vla, forces)
        save(f"lstm_A{A}_F{Fdim}_T{T}", fwd=fwd, loss=loss, seq=seq)


if __name__ == "__main__":
    main()
