"""The step in front of the controller: handing RDT's action chunk to `DiffusionController.predict`.

Reference: `RoboticDiffusionTransformerModel.step` ends with `_unformat_action_to_joint(trajectory).to(torch.float32)`
(scripts/franka_model_eef.py:199-222,312) -- an index-select of the robot's 10 dims out of the 128-wide unified action vector, a
multiply by `[1,...,1,255]` in the policy's dtype (bf16), a widening copy; `inference_fn` then copies the chunk to the host
(scripts/franka_inference_eef.py:186, a stream synchronisation), and the control loop divides the gripper column by 255 in place
and slices the first `act_chunk_execute_step` rows for `controller.predict` (:546, :552-554).  Five launches, three temporaries
and a device->host wait between the policy's last kernel and the controller's first.

Here: ONE launch on the stream the policy ran on reads the unified action vector where the policy left it (no copy, no
synchronisation) and writes both fp32 tensors the script uses; the host copy for the action buffer can be made afterwards,
off the critical path.  Values are bit-identical to the reference's (`tests/test_gpu_parity.py::test_rdt_handoff_*`).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import native as nv

# Positions of [eef_pos_x, eef_pos_y, eef_pos_z, eef_angle_0..5, right_gripper_open] in RDT's unified vector.  The table
# (configs/state_vec.py, STATE_VEC_IDX_MAPPING) belongs to RDT upstream and is NOT in the reference tree: these are upstream's
# published right-arm values; pass `indices=` built from the deployment's own mapping (scripts/franka_model_eef.py:14-24).
RDT_EEF_INDICES = (30, 31, 32, 33, 34, 35, 36, 37, 38, 10)
GRIPPER_SCALE = 255.0

_cache = {}


def _tables(indices: Sequence[int], device) -> Tuple[torch.Tensor, torch.Tensor]:
    key = (tuple(int(i) for i in indices), str(device))
    if key not in _cache:
        idx = torch.tensor(key[0], dtype=torch.int32, device=device)
        scale = torch.ones(len(key[0]), dtype=torch.float32, device=device)
        scale[-1] = GRIPPER_SCALE                                   # franka_model_eef.py:216-219
        _cache[key] = (idx, scale)
    return _cache[key]


def handoff_action_chunk(trajectory: torch.Tensor, act_chunk_execute_step: Optional[int] = None,
                         indices: Sequence[int] = RDT_EEF_INDICES, want_raw: bool = True):
    """trajectory: [B, N, S] bf16 / fp32 CUDA tensor, the policy's unified action vectors (`self.policy.predict_action(...)`).
    Returns (vla_tensor, chunk): `vla_tensor` [B, N, A] fp32 = what `policy.step` returns (None if not want_raw), `chunk`
    [B, act_chunk_execute_step, A] fp32 = what the control loop passes to `controller.predict` (gripper column / 255)."""
    if not trajectory.is_cuda:
        raise nv.NativeError("handoff_action_chunk runs on the policy's CUDA tensor; vla_touch_b200 has no host path")
    if trajectory.dim() != 3:
        raise ValueError(f"expected [B, N, S] unified action vectors, got shape {tuple(trajectory.shape)}")
    if trajectory.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError(f"bf16 or fp32 action vectors expected, got {trajectory.dtype}")
    B, N, S = trajectory.shape
    A = len(indices)
    if max(indices) >= S or min(indices) < 0:
        raise IndexError(f"state index outside the {S}-wide unified vector")
    T = N if act_chunk_execute_step is None else int(act_chunk_execute_step)
    if not 0 < T <= N:
        raise ValueError(f"act_chunk_execute_step {T} outside the chunk of {N} steps")
    x = trajectory.contiguous()
    idx, scale = _tables(indices, x.device)
    raw = torch.empty((B, N, A), dtype=torch.float32, device=x.device) if want_raw else None
    chunk = torch.empty((B, T, A), dtype=torch.float32, device=x.device)
    nv.check(nv.lib().vt_chunk_handoff(x.data_ptr(), nv.VT_BF16 if x.dtype == torch.bfloat16 else nv.VT_F32, B, N, S, idx.data_ptr(),
                                       scale.data_ptr(), A, GRIPPER_SCALE, raw.data_ptr() if want_raw else None, chunk.data_ptr(), T,
                                       nv.current_stream_ptr()))
    return raw, chunk
