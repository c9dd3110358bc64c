"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the VLA-Touch action-refinement hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file; the product path (vla_touch_b200/) never does and fails loudly without its
CUDA library.

Every function restates one reference function in plain functional PyTorch (CPU fp32), working
from a flat state_dict so it needs neither /root/reference nor HF `transformers` at run time.
Citations are file:line into /root/reference/VLA/residual_controller/ unless prefixed `HF:`
(= transformers 5.5.0 models/dinov2/modeling_dinov2.py, the un-vendored dependency that holds
the DinoV2 arithmetic; the reference's call site is visual_encoder.py:27,87-91).

PARITY PINNING: the reference ships no golden vectors / KATs for this path, so
the oracle is pinned against outputs of the *reference itself* run in the build container:
oracle/gen_golden.py imports the unmodified reference modules (oracle/ref_shims.py) with
deterministic synthetic weights and writes tests/golden/*.npz; tests/test_oracle_golden.py
checks every function below against those fixtures (bit-level or <=2e-6).

`quant` (None | "bf16" | "tf32") optionally rounds the *operands* of every contraction the way
the tensor-core path does; it is an error model for choosing test tolerances, not a code path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------
# operand rounding emulation
# ----------------------------------------------------------------------------------------------
def _q(x: torch.Tensor, quant: Optional[str]) -> torch.Tensor:
    if quant is None:
        return x
    if quant == "bf16":
        return x.to(torch.bfloat16).to(torch.float32)
    if quant == "tf32":  # round-to-nearest-even onto a 10-bit mantissa
        i = x.contiguous().view(torch.int32)
        lsb = (i >> 13) & 1
        i = (i + 0x0FFF + lsb) & ~0x1FFF
        return i.view(torch.float32)
    if quant == "tf32t":  # truncation (what the MMA datapath does to raw fp32 operands)
        i = x.contiguous().view(torch.int32) & ~0x1FFF
        return i.view(torch.float32)
    raise ValueError(quant)


def _linear(x, w, b, quant=None):
    return F.linear(_q(x, quant), _q(w, quant), b)


# ----------------------------------------------------------------------------------------------
# controller_dataset.py:303-384
# ----------------------------------------------------------------------------------------------
def normalize_actions(actions, stats, action_type="expert", padding_factor=1.4):
    """controller_dataset.py:303-346"""
    if action_type == "expert":
        mins, maxs = stats["action_mins"], stats["action_maxs"]
    elif action_type == "vla":
        mins, maxs = stats["vla_mins"], stats["vla_maxs"]
    else:
        raise ValueError(f"Unknown action_type: {action_type}")
    orig_range = maxs - mins
    padded_range = orig_range * padding_factor
    center = (mins + maxs) / 2
    padded_mins = center - padded_range / 2
    padded_maxs = center + padded_range / 2
    safe_range = padded_maxs - padded_mins
    safe_range = torch.where(safe_range < 1e-6, torch.ones_like(safe_range), safe_range)
    return 2.0 * (actions - padded_mins) / safe_range - 1.0


def denormalize_actions(normalized_actions, stats, action_type="expert", padding_factor=1.4):
    """controller_dataset.py:349-384 (note: no safe_range guard on this side)"""
    if action_type == "expert":
        mins, maxs = stats["action_mins"], stats["action_maxs"]
    elif action_type == "vla":
        mins, maxs = stats["vla_mins"], stats["vla_maxs"]
    else:
        raise ValueError(f"Unknown action_type: {action_type}")
    orig_range = maxs - mins
    padded_range = orig_range * padding_factor
    center = (mins + maxs) / 2
    padded_mins = center - padded_range / 2
    padded_maxs = center + padded_range / 2
    safe_range = padded_maxs - padded_mins
    return (normalized_actions + 1.0) / 2.0 * safe_range + padded_mins


# ----------------------------------------------------------------------------------------------
# visual_encoder.py:56-106  +  HF Dinov2Model
# ----------------------------------------------------------------------------------------------
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def dinov2_preprocess(images: torch.Tensor) -> torch.Tensor:
    """visual_encoder.py:66-81,95-106 -> pixel_values [B,3,H,W] fp32.
    Both predicates are over the WHOLE call tensor (batch-global)."""
    if not torch.is_tensor(images):            # numpy path :66-67
        images = torch.from_numpy(images).float() / 255.0
    if images.dim() == 5:                      # [B,T,H,W,C]  :70-73
        B, T, H, W, C = images.shape
        images = images.reshape(B * T, H, W, C).permute(0, 3, 1, 2)
    elif images.dim() == 4 and images.shape[-1] == 3:   # [B,H,W,C]  :74-75
        images = images.permute(0, 3, 1, 2)
    if images.max() > 1.0:                     # :78-79 (uint8 / 255.0 -> float32 true division)
        images = images / 255.0
    if images.mean() < 0.5:                    # :100  (whole-call-tensor predicate)
        return images
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    return (images - mean) / std               # :104-106


def dinov2_pos_embed(sd: SD, height: int, width: int, patch: int = 14) -> torch.Tensor:
    """HF:57-95 interpolate_pos_encoding -> [1, 1+h*w, D]"""
    pos = sd["embeddings.position_embeddings"]
    n_pos = pos.shape[1] - 1
    nh, nw = height // patch, width // patch
    if nh * nw == n_pos and height == width:
        return pos
    cls_pos, patch_pos = pos[:, :1], pos[:, 1:]
    dim = pos.shape[-1]
    s = int(n_pos ** 0.5)
    patch_pos = patch_pos.reshape(1, s, s, dim).permute(0, 3, 1, 2)
    patch_pos = F.interpolate(patch_pos.float(), size=(nh, nw), mode="bicubic", align_corners=False)
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((cls_pos, patch_pos), dim=1)


def dinov2_embeddings(sd: SD, pixel_values: torch.Tensor, quant=None) -> torch.Tensor:
    """HF:97-116 + HF:139-149 (Conv2d k14 s14 patch projection, CLS prepend, + pos-emb)"""
    w = sd["embeddings.patch_embeddings.projection.weight"]
    if pixel_values.shape[1] != w.shape[1]:
        raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the "
                         f"configuration. Expected {w.shape[1]} but got {pixel_values.shape[1]}.")
    B, _, H, W = pixel_values.shape
    x = F.conv2d(_q(pixel_values, quant), _q(w, quant), sd["embeddings.patch_embeddings.projection.bias"],
                 stride=w.shape[-1])
    x = x.flatten(2).transpose(1, 2)
    cls = sd["embeddings.cls_token"].expand(B, -1, -1)
    x = torch.cat((cls, x), dim=1)
    return x + dinov2_pos_embed(sd, H, W, w.shape[-1])


def dinov2_layer(sd: SD, i: int, h: torch.Tensor, num_heads: int, eps: float = 1e-6, quant=None) -> torch.Tensor:
    """HF:367-386 Dinov2Layer.forward (attention HF:199-235, output dense HF:272-278,
    LayerScale HF:289-295, MLP HF:312-328)"""
    p = f"encoder.layer.{i}."
    B, N, D = h.shape
    hd = D // num_heads
    x = F.layer_norm(h, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    q = _linear(x, sd[p + "attention.attention.query.weight"], sd[p + "attention.attention.query.bias"], quant)
    k = _linear(x, sd[p + "attention.attention.key.weight"], sd[p + "attention.attention.key.bias"], quant)
    v = _linear(x, sd[p + "attention.attention.value.weight"], sd[p + "attention.attention.value.bias"], quant)
    q = q.view(B, N, num_heads, hd).transpose(1, 2)
    k = k.view(B, N, num_heads, hd).transpose(1, 2)
    v = v.view(B, N, num_heads, hd).transpose(1, 2)
    att = torch.matmul(_q(q, quant), _q(k, quant).transpose(-1, -2)) * (hd ** -0.5)
    att = torch.softmax(att, dim=-1)
    ctx = torch.matmul(_q(att, quant), _q(v, quant)).transpose(1, 2).reshape(B, N, D)
    a = _linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"], quant)
    h = a * sd[p + "layer_scale1.lambda1"] + h
    x = F.layer_norm(h, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    x = _linear(x, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"], quant)
    x = F.gelu(x)  # exact erf GELU (hidden_act="gelu")
    x = _linear(x, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"], quant)
    return x * sd[p + "layer_scale2.lambda1"] + h


def dinov2_num_layers(sd: SD) -> int:
    n = 0
    while f"encoder.layer.{n}.norm1.weight" in sd:
        n += 1
    return n


def dinov2_pooler(sd: SD, pixel_values: torch.Tensor, num_heads: int, quant=None) -> torch.Tensor:
    """HF:463-485 Dinov2Model.forward -> pooler_output = layernorm(h)[:,0]"""
    h = dinov2_embeddings(sd, pixel_values, quant)
    for i in range(dinov2_num_layers(sd)):
        h = dinov2_layer(sd, i, h, num_heads, quant=quant)
    D = h.shape[-1]
    h = F.layer_norm(h, (D,), sd["layernorm.weight"], sd["layernorm.bias"], 1e-6)
    return h[:, 0, :]


def dino_encoder_forward(sd: SD, images: torch.Tensor, num_heads: int, quant=None) -> torch.Tensor:
    """visual_encoder.py:56-93 DINOv2Encoder.forward"""
    return dinov2_pooler(sd, dinov2_preprocess(images), num_heads, quant)


# ----------------------------------------------------------------------------------------------
# bridge_controller.py:42-48,112-134
# ----------------------------------------------------------------------------------------------
def mlp3_gelu(sd: SD, x: torch.Tensor, prefix: str = "", quant=None) -> torch.Tensor:
    """nn.Sequential(Linear, GELU, Linear, GELU, Linear)  bridge_controller.py:42-48"""
    x = F.gelu(_linear(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"], quant))
    x = F.gelu(_linear(x, sd[prefix + "2.weight"], sd[prefix + "2.bias"], quant))
    return _linear(x, sd[prefix + "4.weight"], sd[prefix + "4.bias"], quant)


def encode_observation(dino_sd: SD, enc_sd: SD, num_heads: int, state, img1, img2, forces=None, quant=None):
    """bridge_controller.py:112-134 (two separate encoder calls, :106-107)"""
    f1 = dino_encoder_forward(dino_sd, img1, num_heads, quant)
    f2 = dino_encoder_forward(dino_sd, img2, num_heads, quant)
    if forces is not None:
        state = torch.cat((state, forces), dim=-1)
    return mlp3_gelu(enc_sd, torch.cat((f1, f2, state), dim=-1), quant=quant)


# ----------------------------------------------------------------------------------------------
# bridge/networks/conditional_unet_1D.py
# ----------------------------------------------------------------------------------------------
def sinusoidal_pos_emb(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """conditional_unet_1D.py:7-19 (cat(sin, cos); t is a float in (0,1) on this path)"""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half) * -e)
    e = t[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def _conv1d(x, w, b, stride=1, padding=0, quant=None):
    return F.conv1d(_q(x, quant), _q(w, quant), b, stride=stride, padding=padding)


def _conv_block(sd: SD, p: str, x: torch.Tensor, n_groups: int = 8, quant=None) -> torch.Tensor:
    """Conv1dBlock :40-55  Conv1d(k, pad k//2) -> GroupNorm(8) -> Mish"""
    w = sd[p + "block.0.weight"]
    x = _conv1d(x, w, sd[p + "block.0.bias"], padding=w.shape[-1] // 2, quant=quant)
    x = F.group_norm(x, n_groups, sd[p + "block.1.weight"], sd[p + "block.1.bias"], 1e-5)
    return F.mish(x)


def _res_block(sd: SD, p: str, x: torch.Tensor, gf: torch.Tensor, quant=None) -> torch.Tensor:
    """ConditionalResidualBlock1D.forward :86-105"""
    out = _conv_block(sd, p + "blocks.0.", x, quant=quant)
    emb = _linear(F.mish(gf), sd[p + "cond_encoder.1.weight"], sd[p + "cond_encoder.1.bias"], quant)
    C = out.shape[1]
    emb = emb.reshape(emb.shape[0], 2, C, 1)
    out = emb[:, 0] * out + emb[:, 1]
    out = _conv_block(sd, p + "blocks.1.", out, quant=quant)
    if p + "residual_conv.weight" in sd:
        res = _conv1d(x, sd[p + "residual_conv.weight"], sd[p + "residual_conv.bias"], quant=quant)
    else:
        res = x
    return out + res


def unet_forward(sd: SD, sample: torch.Tensor, timestep: torch.Tensor, global_cond: torch.Tensor,
                 prefix: str = "", quant=None) -> torch.Tensor:
    """DiffusionConditionalUnet1D.forward :194-247.  sample [B,T,A] -> [B,T,A]"""
    g = lambda k: sd[prefix + k]
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)} if prefix else sd
    x = sample.moveaxis(-1, -2)
    t = timestep.expand(sample.shape[0])
    temb = sinusoidal_pos_emb(t, g("diffusion_step_encoder.1.weight").shape[1])
    temb = _linear(temb, g("diffusion_step_encoder.1.weight"), g("diffusion_step_encoder.1.bias"), quant)
    temb = _linear(F.mish(temb), g("diffusion_step_encoder.3.weight"), g("diffusion_step_encoder.3.bias"), quant)
    gf = torch.cat([temb, global_cond], dim=-1)
    h: List[torch.Tensor] = []
    L = 0
    while f"down_modules.{L}.0.blocks.0.block.0.weight" in sub:
        x = _res_block(sub, f"down_modules.{L}.0.", x, gf, quant)
        x = _res_block(sub, f"down_modules.{L}.1.", x, gf, quant)
        h.append(x)
        if f"down_modules.{L}.2.conv.weight" in sub:   # Downsample1d :22-28 (last level: Identity)
            x = _conv1d(x, sub[f"down_modules.{L}.2.conv.weight"], sub[f"down_modules.{L}.2.conv.bias"],
                        stride=2, padding=1, quant=quant)
        L += 1
    for m in range(2):
        x = _res_block(sub, f"mid_modules.{m}.", x, gf, quant)
    U = 0
    while f"up_modules.{U}.0.blocks.0.block.0.weight" in sub:
        x = torch.cat((x, h.pop()), dim=1)
        x = _res_block(sub, f"up_modules.{U}.0.", x, gf, quant)
        x = _res_block(sub, f"up_modules.{U}.1.", x, gf, quant)
        if f"up_modules.{U}.2.conv.weight" in sub:     # Upsample1d :31-37 ConvTranspose1d(4,2,1)
            x = F.conv_transpose1d(_q(x, quant), _q(sub[f"up_modules.{U}.2.conv.weight"], quant),
                                   sub[f"up_modules.{U}.2.conv.bias"], stride=2, padding=1)
        U += 1
    x = _conv_block(sub, "final_conv.0.", x, quant=quant)
    x = _conv1d(x, sub["final_conv.1.weight"], sub["final_conv.1.bias"], quant=quant)
    return x.moveaxis(-1, -2)


# ----------------------------------------------------------------------------------------------
# bridge/bridge_model.py  (only the configuration every script selects: SURVEY App. B)
# ----------------------------------------------------------------------------------------------
T_MIN = 0.001
GAMMA_INV_MAX = 200.0


def sde_schedule(n_steps_arg: int):
    """bridge_model.py:259-279,334-348.  Returns (n_steps, delta_t, [t_k]) exactly as the
    reference derives them: delta_t = float(1.0/diffuse_step); n = int(1.0/delta_t)."""
    delta_t = float(1.0 / n_steps_arg)
    n_steps = int(1.0 / delta_t)
    ts = []
    for k in range(1, n_steps + 1):
        t = torch.full((1,), k / n_steps).float()
        ts.append(torch.clip(t, T_MIN, 1.0 - T_MIN))
    return n_steps, delta_t, ts


def sde_coefficients(t: torch.Tensor, delta_t: float, d: float):
    """bridge_model.py:59-101,347-385 for gamma '2^0.5*t(t-1)', epsilon '1-t' (fp32 tensor math,
    same operation order as the reference).  Returns python floats
    (gamma_inv, dot_gamma*gamma, eps, noise_scale=dt*sqrt(2 eps), d)."""
    gamma = 1.4142 * t * (1 - t)
    dgamma = 1.4142 * (1 - 2 * t)
    ginv = torch.clamp(1 / (1.4142 * t * (1 - t) + 1e-4), 0.0, GAMMA_INV_MAX)
    eps = (1 - t) * 1.0
    noise_scale = delta_t * torch.sqrt(2 * eps)
    return ginv, dgamma * gamma, eps, noise_scale


def sde_vs(v_sd: SD, s_sd: SD, x_initial: torch.Tensor, cond: torch.Tensor, diffuse_step: int = 10,
           beta_max: float = 0.03, noise: Optional[torch.Tensor] = None, quant=None,
           return_traj: bool = False):
    """StochasticInterpolants.sample -> sde_vs (forward direction, score_weight=1)
    bridge_model.py:259-279,334-387.  `noise` = [n_steps,B,T,A] standard-normal draws in loop
    order (what torch.randn_like returns at :372); None -> draws with torch.randn_like."""
    n_steps, delta_t, ts = sde_schedule(diffuse_step)
    B = x_initial.shape[0]
    x = x_initial
    traj = [x]
    for k in range(n_steps):
        t = ts[k].expand(B)
        ginv, dgg, eps, noise_scale = sde_coefficients(t, delta_t, beta_max)
        v = unet_forward(v_sd, x, t, cond, quant=quant)
        s = unet_forward(s_sd, x, t, cond, quant=quant)
        s = s * ginv[:, None, None]
        b = v - dgg[:, None, None] * s * eps[0]
        z = torch.randn_like(x) if noise is None else noise[k]
        dW = beta_max * z
        new_x = x + (b + 1.0 * eps[0] * s) * delta_t
        new_x = new_x + noise_scale[0] * dW
        x = new_x
        traj.append(x)
    return (x, traj) if return_traj else x


def sde_bs(b_sd: SD, s_sd: SD, x_initial: torch.Tensor, cond: torch.Tensor, diffuse_step: int = 10,
           beta_max: float = 0.03, noise: Optional[torch.Tensor] = None, quant=None):
    """StochasticInterpolants.sample with sde_type='bs' -> sde_bs (forward direction, score_weight=1), bridge_model.py:281-332:
    the drift comes straight from b_net instead of v_net - dot_gamma gamma eps s."""
    n_steps, delta_t, ts = sde_schedule(diffuse_step)
    B = x_initial.shape[0]
    x = x_initial
    for k in range(n_steps):
        t = ts[k].expand(B)
        ginv, _, eps, noise_scale = sde_coefficients(t, delta_t, beta_max)
        b = unet_forward(b_sd, x, t, cond, quant=quant)
        s = unet_forward(s_sd, x, t, cond, quant=quant) * ginv[:, None, None]
        z = torch.randn_like(x) if noise is None else noise[k]
        new_x = x + (b + 1.0 * eps[0] * s) * delta_t
        x = new_x + noise_scale[0] * (beta_max * z)
    return x


def predict(dino_sd: SD, enc_sd: SD, v_sd: SD, s_sd: SD, stats, num_heads: int, state, vla_actions, img1, img2,
            forces, diffuse_step: int = 10, beta_max: float = 0.03, noise=None, quant=None):
    """DiffusionController.predict  bridge_controller.py:149-182 (v_sd/s_sd = EMA weights, :267)"""
    cond = encode_observation(dino_sd, enc_sd, num_heads, state, img1, img2, forces, quant)
    x0 = normalize_actions(vla_actions, stats, "vla")
    x = sde_vs(v_sd, s_sd, x0, cond, diffuse_step, beta_max, noise, quant)
    return denormalize_actions(x, stats, "expert")


# ----------------------------------------------------------------------------------------------
# losses  bridge_model.py:183-257 (given recorded t and z)
# ----------------------------------------------------------------------------------------------
def bridge_losses(net_sd: SD, obs_cond, expert_act, vla_act, step: torch.Tensor, z_unit: torch.Tensor,
                  beta_max: float = 0.03):
    """get_loss :220-246 with the two RNG draws injected: step=torch.rand(B) (:236),
    z_unit=torch.randn_like(x0) (:105).  Returns (loss, v_loss, s_loss, b_loss)."""
    x0, x1 = vla_act, expert_act
    tb = torch.clip(step[:, None, None], T_MIN, 1.0 - T_MIN)
    gamma = 1.4142 * tb * (1 - tb)
    z = beta_max * z_unit
    xt = (1 - tb) * x0 + tb * x1 + gamma * z
    t = torch.clip(step, T_MIN, 1.0 - T_MIN)
    sub = lambda n: {k[len(n):]: v for k, v in net_sd.items() if k.startswith(n)}
    v = unet_forward(sub("v_net."), xt, t, obs_cond).flatten(-2)
    s = unet_forward(sub("s_net."), xt, t, obs_cond).flatten(-2)
    b = unet_forward(sub("b_net."), xt, t, obs_cond).flatten(-2)
    pt = (x1 - x0).flatten(-2)
    zr = z.flatten(-2)
    v_loss = torch.mean(0.5 * torch.norm(v, dim=-1) ** 2 - torch.sum(pt * v, dim=-1))
    s_loss = torch.mean(0.5 * torch.norm(s, dim=-1) ** 2 + torch.sum(zr * s, dim=-1))
    gd = (1.4142 * (1 - 2 * t))[:, None]
    b_loss = torch.mean(0.5 * torch.norm(b, dim=-1) ** 2 - torch.sum((pt + gd * zr) * b, dim=-1))
    return v_loss + s_loss + b_loss, v_loss, s_loss, b_loss


# ----------------------------------------------------------------------------------------------
# lstm_step_controller.py:148-319 (eval mode: dropout off)
# ----------------------------------------------------------------------------------------------
def lstm_forward(mods: Dict[str, SD], vla_act_n, obs_cond, forces, state=None, return_state=False):
    """TactileLSTMController.forward :170-213 (eval).  mods = {force_encoder, lstm, output_head}
    state = (h[L,B,H], c[L,B,H]) or None (zeros, :196-197)"""
    fe, lstm, head = mods["force_encoder"], mods["lstm"], mods["output_head"]
    B, T, _ = vla_act_n.shape
    f = forces.reshape(B * T, -1)
    f = _linear(F.gelu(_linear(f, fe["0.weight"], fe["0.bias"])), fe["2.weight"], fe["2.bias"]).reshape(B, T, -1)
    x = torch.cat([f, vla_act_n], dim=-1)
    L = 0
    while f"weight_ih_l{L}" in lstm:
        L += 1
    H = lstm["weight_hh_l0"].shape[1]
    if state is None:
        h = [torch.zeros(B, H) for _ in range(L)]
        c = [torch.zeros(B, H) for _ in range(L)]
    else:
        h = [state[0][l] for l in range(L)]
        c = [state[1][l] for l in range(L)]
    outs = []
    for t in range(T):
        inp = x[:, t]
        for l in range(L):  # PyTorch gate order i,f,g,o
            g = (F.linear(inp, lstm[f"weight_ih_l{l}"], lstm[f"bias_ih_l{l}"])
                 + F.linear(h[l], lstm[f"weight_hh_l{l}"], lstm[f"bias_hh_l{l}"]))
            i_, f_, g_, o_ = g.chunk(4, dim=-1)
            c[l] = torch.sigmoid(f_) * c[l] + torch.sigmoid(i_) * torch.tanh(g_)
            h[l] = torch.sigmoid(o_) * torch.tanh(c[l])
            inp = h[l]
        outs.append(inp)
    y = torch.stack(outs, dim=1)
    comb = torch.cat([y, obs_cond.unsqueeze(1).repeat(1, T, 1)], dim=-1)
    z = F.linear(comb, head["0.weight"], head["0.bias"])
    z = F.layer_norm(z, (z.shape[-1],), head["1.weight"], head["1.bias"], 1e-5)
    delta = F.linear(F.gelu(z), head["4.weight"], head["4.bias"])
    out = vla_act_n + delta
    if return_state:
        return out, (torch.stack(h), torch.stack(c))
    return out
