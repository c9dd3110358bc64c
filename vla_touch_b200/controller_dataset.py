"""normalize_actions / denormalize_actions with the reference's signature and error behaviour
(controller_dataset.py:303-384): affine map onto [-1, 1] over the min/max range padded by `padding_factor`.
On CUDA tensors the map runs in the AFFINE kernel of libvt_b200 (bit-identical to the PyTorch fp32 arithmetic:
same operation order, no FMA contraction).  Dataset loading (ControllerDataset / HDF5) is out of scope."""
from __future__ import annotations

import torch

from . import native as nv
from .plan import ptr


def _stats(stats, action_type, device):
    if action_type == 'expert':
        mins, maxs = stats['action_mins'], stats['action_maxs']
    elif action_type == 'vla':
        mins, maxs = stats['vla_mins'], stats['vla_maxs']
    else:
        raise ValueError(f"Unknown action_type: {action_type}")
    conv = lambda v: torch.as_tensor(v, dtype=torch.float32).to(device).contiguous()
    return conv(mins), conv(maxs)


def _affine(actions, stats, action_type, padding_factor, denorm):
    if not torch.is_tensor(actions):
        actions = torch.as_tensor(actions, dtype=torch.float32)
    if actions.device.type != "cuda":
        raise nv.NativeError("vla_touch_b200 runs on a CUDA (B200) device only: move `actions` to the GPU")
    mins, maxs = _stats(stats, action_type, actions.device)
    x = actions.to(torch.float32).contiguous()
    A = x.shape[-1]
    if mins.numel() != A:
        raise ValueError(f"stats have {mins.numel()} dims, actions have {A}")
    out = torch.empty_like(x)
    d = nv.AffineDesc()
    d.x, d.out, d.mins, d.maxs, d.rows, d.A, d.denorm, d.pad = ptr(x), ptr(out), ptr(mins), ptr(maxs), x.numel() // A, A, denorm, padding_factor
    prog = nv.Program()
    prog.add(d)
    prog.run()
    return out


def normalize_actions(actions, stats, action_type='expert', padding_factor=1.4):
    return _affine(actions, stats, action_type, padding_factor, 0)


def denormalize_actions(normalized_actions, stats, action_type='expert', padding_factor=1.4):
    return _affine(normalized_actions, stats, action_type, padding_factor, 1)
