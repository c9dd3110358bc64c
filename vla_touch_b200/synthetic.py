"""Deterministic synthetic weights and inputs (no network => no pretrained checkpoints).

Everything here is a pure function of (name, shape, seed) through numpy's PCG64 stream, whose
output is stable across numpy versions, so the golden-vector generator (oracle/gen_golden.py,
run once next to the reference), the parity tests and bench.py all see bit-identical tensors
without having to commit ~100 MB of weights.

Input recipes follow SURVEY.md section 8(d) ("Synthetic inputs").
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import numpy as np
import torch


def _rng(name: str, seed: int) -> np.random.Generator:
    key = (zlib.crc32(name.encode("utf-8")) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFFFFFF
    return np.random.Generator(np.random.PCG64(key))


def det_normal(name: str, shape: Tuple[int, ...], seed: int = 0, scale: float = 1.0) -> torch.Tensor:
    """N(0, scale^2) tensor that depends only on (name, shape, seed)."""
    a = _rng(name, seed).standard_normal(size=tuple(shape), dtype=np.float32) * np.float32(scale)
    return torch.from_numpy(np.ascontiguousarray(a))


def det_uniform(name: str, shape: Tuple[int, ...], seed: int = 0, lo: float = 0.0, hi: float = 1.0) -> torch.Tensor:
    a = _rng(name, seed).random(size=tuple(shape), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)
    return torch.from_numpy(np.ascontiguousarray(a))


def synth_param(name: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    """Well-conditioned synthetic value for a parameter called `name` of shape `shape`.

    matrices / conv kernels : N(0, 1/fan_in)      (fan_in = prod(shape[1:]))
    norm scales (1-D weight): 1 + 0.1 N(0,1)
    biases                  : 0.05 N(0,1)
    LayerScale lambda1      : 1 + 0.1 N(0,1)      (HF random-init uses layerscale_value=1.0)
    cls / position tokens   : 0.02 N(0,1)         (HF trunc-normal(0.02) init)
    """
    shape = tuple(int(s) for s in shape)
    leaf = name.rsplit(".", 1)[-1]
    if leaf in ("cls_token", "position_embeddings"):
        return det_normal(name, shape, seed, 0.02)
    if leaf == "mask_token":
        return torch.zeros(shape, dtype=torch.float32)
    if len(shape) >= 2:
        fan_in = int(np.prod(shape[1:]))
        return det_normal(name, shape, seed, 1.0 / np.sqrt(max(fan_in, 1)))
    if leaf.startswith("bias"):
        return det_normal(name, shape, seed, 0.05)
    # 1-D "weight" (LayerNorm / GroupNorm affine) and LayerScale "lambda1"
    return 1.0 + det_normal(name, shape, seed, 0.1)


@torch.no_grad()
def fill_named_(named_tensors: Iterable[Tuple[str, torch.Tensor]], seed: int = 0, prefix: str = "") -> None:
    """In-place deterministic fill of (name, tensor) pairs, e.g. module.named_parameters()."""
    for name, p in named_tensors:
        p.copy_(synth_param(prefix + name, tuple(p.shape), seed).to(p.dtype))


def synth_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, prefix: str = "") -> Dict[str, torch.Tensor]:
    return {k: synth_param(prefix + k, s, seed) for k, s in shapes.items()}


# --------------------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------------------

def synth_images_u8(name: str, batch: int, hw: int, seed: int = 0, dark: bool = False) -> torch.Tensor:
    """[B,H,W,3] uint8.  bright: 128 + floor(64 U) (mean ~0.62 -> ImageNet-normalise branch of
    visual_encoder.py:100); dark: floor(100 U) (mean ~0.19 -> skip branch)."""
    u = _rng(name, seed).random(size=(batch, hw, hw, 3), dtype=np.float32)
    a = np.floor(u * 100.0) if dark else 128.0 + np.floor(u * 64.0)
    return torch.from_numpy(a.astype(np.uint8))


def synth_predict_inputs(batch: int, horizon: int, action_dim: int, force_dim: int, hw: int, seed: int = 0,
                         dark: bool = False) -> Dict[str, torch.Tensor]:
    return {
        "state": det_normal("in.state", (batch, action_dim), seed),
        "forces": det_normal("in.forces", (batch, force_dim), seed),
        "vla_actions": det_uniform("in.vla", (batch, horizon, action_dim), seed, -1.0, 1.0),
        "images_cam1": synth_images_u8("in.cam1", batch, hw, seed, dark),
        "images_cam2": synth_images_u8("in.cam2", batch, hw, seed, dark),
    }


def synth_stats(action_dim: int) -> Dict[str, torch.Tensor]:
    """mins=-1, maxs=+1 per dim -> normalise is a pure x/1.4."""
    mins = -torch.ones(action_dim)
    maxs = torch.ones(action_dim)
    return {
        "action_mins": mins.clone(), "action_maxs": maxs.clone(), "action_range": (maxs - mins).clone(),
        "vla_mins": mins.clone(), "vla_maxs": maxs.clone(), "vla_range": (maxs - mins).clone(),
    }


def synth_stats_varied(action_dim: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Non-trivial stats (different per dim and per action type), incl. one degenerate range."""
    a_lo = -1.0 - det_uniform("st.alo", (action_dim,), seed)
    a_hi = 0.5 + det_uniform("st.ahi", (action_dim,), seed)
    v_lo = -0.8 - det_uniform("st.vlo", (action_dim,), seed)
    v_hi = 0.7 + det_uniform("st.vhi", (action_dim,), seed)
    v_hi[-1] = v_lo[-1]  # degenerate vla range -> safe_range:=1 (controller_dataset.py:341-342)
    return {
        "action_mins": a_lo, "action_maxs": a_hi, "action_range": a_hi - a_lo,
        "vla_mins": v_lo, "vla_maxs": v_hi, "vla_range": v_hi - v_lo,
    }


def synth_episode(seed: int, n_frames: int, image_size: int = 28, vla_T: int = 64, still_frames: int = 3, moving: bool = True,
                  dark: bool = False, images: bool = True) -> Dict[str, object]:
    """One episode with the reference's HDF5 schema (data/create_controller_dataset_episode.py:179-188) as a nested dict of numpy
    arrays: the end effector stands still for `still_frames` frames (the dataset skips those, controller_dataset.py:80-92), then
    random-walks; gripper in [0, 255]; `moving=False` gives an episode the dataset must skip entirely."""
    g = np.random.default_rng([seed, n_frames])
    step = g.normal(0.0, 0.02, (n_frames, 3))
    step[:still_frames + 1] = 0.0
    if not moving:
        step[:] = 0.0
    pos = np.array([0.4, 0.0, 0.3]) + np.cumsum(step, axis=0)
    quat = g.normal(0.0, 1.0, (n_frames, 4))
    quat[:still_frames + 1] = quat[0]
    if not moving:
        quat[:] = quat[0]
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    epi = {
        "ee_poses": np.concatenate((pos, quat), axis=1),
        "gripper_pos": np.floor(g.uniform(0.0, 256.0, n_frames)),
        "vla_action": np.concatenate((g.normal(0.0, 0.5, (n_frames, vla_T, 9)), g.uniform(0.0, 255.0, (n_frames, vla_T, 1))), axis=2).astype(np.float32),
        "gelsight_force": {"forces": g.normal(0.0, 1.0, (n_frames, 3)), "displacement": g.normal(0.0, 2.0, (n_frames, 63, 2))},
    }
    if images:
        lo, hi = (0, 100) if dark else (96, 256)
        epi["camera1_resized"] = g.integers(lo, hi, (n_frames, image_size, image_size, 3), dtype=np.uint8)
        epi["camera2_resized"] = g.integers(lo, hi, (n_frames, image_size, image_size, 3), dtype=np.uint8)
    return epi
