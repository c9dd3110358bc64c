"""State-dict key/shape contracts of the reference modules on the hot path.

These tables are the checkpoint-interchange contract: `controller.pt`, `bridge_model.pt`
(`net` + torch_ema `shadow_params` in `net.parameters()` ORDER) and `tactile_controller.pt`
written by the reference load here and vice versa.  Orders follow the reference's module
registration order (conditional_unet_1D.py:143-190: mid_modules first, then
diffusion_step_encoder, up_modules, down_modules, final_conv; conditional_unet_1D_si.py:25-49:
b_net, v_net, s_net) and are verified against the reference in tests/test_contract.py.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Sequence, Tuple

Shapes = "OrderedDict[str, Tuple[int, ...]]"

DINO_VARIANTS = {
    # name fragment -> (hidden, heads, layers)   visual_encoder.py:31-46, HF configs
    "small": (384, 6, 12),
    "base": (768, 12, 12),
    "large": (1024, 16, 24),
}


def dino_variant(model_name: str) -> Tuple[int, int, int]:
    for k, v in DINO_VARIANTS.items():
        if k in model_name:
            return v
    if "giant" in model_name:
        raise NotImplementedError("dinov2-giant (SwiGLU MLP, HF:331-345) is not implemented")
    return DINO_VARIANTS["small"]  # visual_encoder.py:43-46 default


def dinov2_shapes(hidden: int, layers: int, patch: int = 14, pos_grid: int = 37, mlp_ratio: int = 4) -> Shapes:
    D = hidden
    s: Shapes = OrderedDict()
    s["embeddings.cls_token"] = (1, 1, D)
    s["embeddings.mask_token"] = (1, D)
    s["embeddings.position_embeddings"] = (1, pos_grid * pos_grid + 1, D)
    s["embeddings.patch_embeddings.projection.weight"] = (D, 3, patch, patch)
    s["embeddings.patch_embeddings.projection.bias"] = (D,)
    for i in range(layers):
        p = f"encoder.layer.{i}."
        s[p + "norm1.weight"] = (D,)
        s[p + "norm1.bias"] = (D,)
        for n in ("query", "key", "value"):
            s[p + f"attention.attention.{n}.weight"] = (D, D)
            s[p + f"attention.attention.{n}.bias"] = (D,)
        s[p + "attention.output.dense.weight"] = (D, D)
        s[p + "attention.output.dense.bias"] = (D,)
        s[p + "layer_scale1.lambda1"] = (D,)
        s[p + "norm2.weight"] = (D,)
        s[p + "norm2.bias"] = (D,)
        s[p + "mlp.fc1.weight"] = (mlp_ratio * D, D)
        s[p + "mlp.fc1.bias"] = (mlp_ratio * D,)
        s[p + "mlp.fc2.weight"] = (D, mlp_ratio * D)
        s[p + "mlp.fc2.bias"] = (D,)
        s[p + "layer_scale2.lambda1"] = (D,)
    s["layernorm.weight"] = (D,)
    s["layernorm.bias"] = (D,)
    return s


def mlp_shapes(dims: Sequence[int]) -> Shapes:
    """nn.Sequential(Linear, GELU, Linear, GELU, ...): Linear i sits at index 2*i."""
    s: Shapes = OrderedDict()
    for i in range(len(dims) - 1):
        s[f"{2 * i}.weight"] = (dims[i + 1], dims[i])
        s[f"{2 * i}.bias"] = (dims[i + 1],)
    return s


def _crb(s: Shapes, p: str, cin: int, cout: int, cond_dim: int, k: int) -> None:
    """ConditionalResidualBlock1D (conditional_unet_1D.py:58-84) registration order."""
    for b, ci in ((0, cin), (1, cout)):
        s[p + f"blocks.{b}.block.0.weight"] = (cout, ci, k)
        s[p + f"blocks.{b}.block.0.bias"] = (cout,)
        s[p + f"blocks.{b}.block.1.weight"] = (cout,)
        s[p + f"blocks.{b}.block.1.bias"] = (cout,)
    s[p + "cond_encoder.1.weight"] = (2 * cout, cond_dim)
    s[p + "cond_encoder.1.bias"] = (2 * cout,)
    if cin != cout:
        s[p + "residual_conv.weight"] = (cout, cin, 1)
        s[p + "residual_conv.bias"] = (cout,)


def unet_shapes(action_dim: int, global_cond_dim: int = 256, dsed: int = 256,
                down_dims: Sequence[int] = (256, 512, 512), k: int = 5) -> Shapes:
    """DiffusionConditionalUnet1D (conditional_unet_1D.py:108-190)."""
    all_dims = [action_dim] + list(down_dims)
    in_out = list(zip(all_dims[:-1], all_dims[1:]))
    cond_dim = dsed + global_cond_dim
    mid = all_dims[-1]
    s: Shapes = OrderedDict()
    for m in range(2):
        _crb(s, f"mid_modules.{m}.", mid, mid, cond_dim, k)
    s["diffusion_step_encoder.1.weight"] = (dsed * 4, dsed)
    s["diffusion_step_encoder.1.bias"] = (dsed * 4,)
    s["diffusion_step_encoder.3.weight"] = (dsed, dsed * 4)
    s["diffusion_step_encoder.3.bias"] = (dsed,)
    for ind, (din, dout) in enumerate(reversed(in_out[1:])):
        _crb(s, f"up_modules.{ind}.0.", dout * 2, din, cond_dim, k)
        _crb(s, f"up_modules.{ind}.1.", din, din, cond_dim, k)
        # `is_last = ind >= len(in_out)-1` (:169) can never be true here -> always Upsample1d
        s[f"up_modules.{ind}.2.conv.weight"] = (din, din, 4)      # ConvTranspose1d layout (Cin,Cout,k)
        s[f"up_modules.{ind}.2.conv.bias"] = (din,)
    for ind, (din, dout) in enumerate(in_out):
        _crb(s, f"down_modules.{ind}.0.", din, dout, cond_dim, k)
        _crb(s, f"down_modules.{ind}.1.", dout, dout, cond_dim, k)
        if ind < len(in_out) - 1:
            s[f"down_modules.{ind}.2.conv.weight"] = (dout, dout, 3)
            s[f"down_modules.{ind}.2.conv.bias"] = (dout,)
    start = down_dims[0]
    s["final_conv.0.block.0.weight"] = (start, start, k)
    s["final_conv.0.block.0.bias"] = (start,)
    s["final_conv.0.block.1.weight"] = (start,)
    s["final_conv.0.block.1.bias"] = (start,)
    s["final_conv.1.weight"] = (action_dim, start, 1)
    s["final_conv.1.bias"] = (action_dim,)
    return s


def si_net_shapes(action_dim: int, global_cond_dim: int = 256) -> Shapes:
    """InterpolantsConditionalUnet1D: b_net, v_net, s_net (conditional_unet_1D_si.py:25-49)."""
    one = unet_shapes(action_dim, global_cond_dim)
    s: Shapes = OrderedDict()
    for net in ("b_net", "v_net", "s_net"):
        for k_, v in one.items():
            s[f"{net}.{k_}"] = v
    return s


def lstm_shapes(input_dim: int, hidden: int = 256, layers: int = 2) -> Shapes:
    """nn.LSTM parameter order: per layer weight_ih, weight_hh, bias_ih, bias_hh (gates i,f,g,o)."""
    s: Shapes = OrderedDict()
    for l in range(layers):
        s[f"weight_ih_l{l}"] = (4 * hidden, input_dim if l == 0 else hidden)
        s[f"weight_hh_l{l}"] = (4 * hidden, hidden)
        s[f"bias_ih_l{l}"] = (4 * hidden,)
        s[f"bias_hh_l{l}"] = (4 * hidden,)
    return s


def lstm_head_shapes(hidden: int, action_dim: int) -> Shapes:
    """output_head = Sequential(Linear(2H,H), LayerNorm(H), GELU, Dropout, Linear(H,A))
    lstm_step_controller.py:76-82"""
    s: Shapes = OrderedDict()
    s["0.weight"] = (hidden, 2 * hidden)
    s["0.bias"] = (hidden,)
    s["1.weight"] = (hidden,)
    s["1.bias"] = (hidden,)
    s["4.weight"] = (action_dim, hidden)
    s["4.bias"] = (action_dim,)
    return s
