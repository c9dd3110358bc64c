"""Shared helpers for the tests: golden fixtures + deterministic weights (same recipe as
oracle/gen_golden.py, which produced the fixtures from the reference itself)."""
import os

import numpy as np
import torch

from vla_touch_b200 import shapes as shp
from vla_touch_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOK_IDX = lambda n: list(range(8)) + list(range(n - 4, n))


def golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: (torch.from_numpy(np.asarray(z[k])) if z[k].dtype.kind in "fiub" else z[k]) for k in z.files}   # names stay numpy


def dino_sd(hidden, layers, seed):
    return syn.synth_state_dict(shp.dinov2_shapes(hidden, layers), seed, prefix="dino.")


def enc_sd(obs_dim, seed, hidden=256):
    return syn.synth_state_dict(shp.mlp_shapes([obs_dim, hidden, hidden, hidden]), seed, prefix="enc.")


def net_sd(A, seed, which=None, cond=256):
    """Full InterpolantsConditionalUnet1D state dict ('b_net.'/'v_net.'/'s_net.' prefixes), or one
    sub-net's un-prefixed dict when which in {'v_net','s_net','b_net'}."""
    full = syn.synth_state_dict(shp.si_net_shapes(A, cond), seed, prefix="net.")
    if which is None:
        return full
    p = which + "."
    return {k[len(p):]: v for k, v in full.items() if k.startswith(p)}


def images_for(kind, name, batch, hw, seed):
    u8 = syn.synth_images_u8(name, batch, hw, seed, dark=(kind == "u8dark"))
    if kind in ("u8bright5d", "u8dark"):
        return u8[:, None]
    if kind == "f32bhwc":
        return u8.float() / 255.0
    if kind == "f32bchw":
        return (u8.float() / 255.0).permute(0, 3, 1, 2).contiguous()
    if kind == "u8bhwc":
        return u8
    raise ValueError(kind)


DINO_CASES = [
    # tag, hidden, heads, layers, hw, batch, kind, seed
    ("dino_s12_224_u8bright5d", 384, 6, 12, 224, 2, "u8bright5d", 11),
    ("dino_s2_224_u8dark", 384, 6, 2, 224, 2, "u8dark", 12),
    ("dino_s2_224_f32bhwc", 384, 6, 2, 224, 2, "f32bhwc", 13),
    ("dino_s2_224_f32bchw", 384, 6, 2, 224, 2, "f32bchw", 14),
    ("dino_s2_224_u8bhwc", 384, 6, 2, 224, 3, "u8bhwc", 15),
    ("dino_s2_384_u8bright5d", 384, 6, 2, 384, 1, "u8bright5d", 16),
    ("dino_b2_224_u8bright5d", 768, 12, 2, 224, 1, "u8bright5d", 17),
]

PREDICT_CASES = {
    # tag: A, F, T, hw, B, layers, hidden, heads, seed, steps, dark, kind, varstats
    "predict_cfg1": (10, 3, 16, 384, 1, 12, 384, 6, 31, 10, False, "u8_5d", False),
    "predict_cfg2_B2": (7, 64, 64, 224, 2, 12, 384, 6, 32, 10, False, "u8_5d", False),
    "predict_cfg2_B3_dark_varstats": (7, 64, 64, 224, 3, 2, 384, 6, 33, 10, True, "u8_5d", True),
    "predict_T48_f32_varstats": (10, 3, 48, 224, 2, 2, 384, 6, 34, 10, False, "f32_bhwc", True),
    "predict_cfg3_B1_base": (7, 64, 64, 224, 1, 2, 768, 12, 35, 50, False, "u8_5d", False),
}


def predict_case(tag):
    A, Fd, T, hw, B, layers, hidden, heads, seed, steps, dark, kind, varstats = PREDICT_CASES[tag]
    inp = syn.synth_predict_inputs(B, T, A, Fd, hw, seed, dark)
    i1, i2 = inp["images_cam1"], inp["images_cam2"]
    if kind == "u8_5d":
        i1, i2 = i1[:, None], i2[:, None]
    elif kind == "f32_bhwc":
        i1, i2 = i1.float() / 255.0, i2.float() / 255.0
    stats = syn.synth_stats_varied(A, seed) if varstats else syn.synth_stats(A)
    return dict(A=A, F=Fd, T=T, hw=hw, B=B, layers=layers, hidden=hidden, heads=heads, seed=seed, steps=steps,
                state=inp["state"], vla=inp["vla_actions"], img1=i1, img2=i2, forces=inp["forces"], stats=stats,
                dino=dino_sd(hidden, layers, seed), enc=enc_sd(2 * hidden + A + Fd, seed),
                v_ema=net_sd(A, seed + 1000, "v_net"), s_ema=net_sd(A, seed + 1000, "s_net"),
                gold=golden(tag))


def make_controller(c, device, precise=False):
    """DiffusionController (public API) holding exactly the weights of a PREDICT_CASES entry: live net = seed, EMA
    shadow = seed + 1000 (so that a test passes only if sample() really uses the EMA weights, bridge_model.py:267)."""
    from vla_touch_b200.bridge_controller import DiffusionController
    model_args = {'interpolant_type': 'linear', 'gamma_type': '2^0.5*t(t-1)', 'epsilon_type': '1-t', 'prior_policy': 'vla',
                  'beta_max': 0.03, 'sde_type': 'vs', 'action_dim': c["A"], 'obs_dim': 256, 'obs_horizon': 1,
                  'net_type': 'unet1D_si', 'pretrain': False, 'context_frames': 2, 'horizon': c["T"]}
    name = "facebook/dinov2-base" if c["hidden"] == 768 else "facebook/dinov2-small"
    ctl = DiffusionController(state_dim=c["A"], hidden_dim=256, image_model_path=name, diffusion_steps=c["steps"], device=device,
                              model_args=model_args, use_force=True, force_dim=c["F"], image_state_dict=c["dino"], precise=precise)
    ctl.state_encoder.load_state_dict(c["enc"])
    ctl.diffusion_model.net.load_state_dict(net_sd(c["A"], c["seed"]))
    ema_full = net_sd(c["A"], c["seed"] + 1000)
    names = [n for n, _ in ctl.diffusion_model.net.named_parameters()]
    with torch.no_grad():
        for n, s in zip(names, ctl.diffusion_model.ema.shadow_params):
            s.copy_(ema_full[n])
    ctl.diffusion_model.ema.version += 1
    ctl.stats = {k: v.to(device) for k, v in c["stats"].items()}
    return ctl
