"""Camera-frame preprocessing in front of the controller: `pad_and_resize_for_siglip` of the deployment
script (scripts/utils_eef.py:44-77, called at scripts/franka_inference_eef.py:329-330) on the GPU.

The reference pads the H x W x C uint8 frame to a centred square on the host and calls cv2.resize(..., INTER_AREA); here the raw
frame is uploaded once and one kernel (csrc/vt_resize.cuh) writes the target x target x C result -- bit-identical to OpenCV --
which `DiffusionController.predict` accepts directly as a device uint8 image, so the resized frame never exists on the host."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import native as nv


def pad_and_resize_for_siglip(image, target_size: int = 384, device="cuda") -> Optional[torch.Tensor]:
    """image: uint8 [H, W, C] or a batch [N, H, W, C] (numpy array, CPU tensor or CUDA tensor); None passes through like the
    reference.  Returns a uint8 CUDA tensor [target, target, C] ([N, ...] for a batch).  Raises NotImplementedError for frames whose
    longer side is smaller than `target_size` (INTER_AREA up-scaling is a different OpenCV code path)."""
    if image is None:
        return None
    t = torch.from_numpy(np.ascontiguousarray(image)) if isinstance(image, np.ndarray) else image
    if t.dtype != torch.uint8:
        raise TypeError(f"pad_and_resize_for_siglip expects uint8 frames, got {t.dtype}")
    if t.dim() not in (3, 4):
        raise ValueError(f"pad_and_resize_for_siglip expects [H, W, C] or [N, H, W, C], got shape {tuple(t.shape)}")
    batched = t.dim() == 4
    if not batched:
        t = t[None]
    n, h, w, c = t.shape
    if max(h, w) < target_size:
        raise NotImplementedError(f"INTER_AREA up-scaling ({max(h, w)} -> {target_size}) is not built")
    if not 1 <= c <= 4:
        raise ValueError(f"1..4 channels supported, got {c}")
    src = t.to(device, non_blocking=True).contiguous()
    dst = torch.empty((n, target_size, target_size, c), dtype=torch.uint8, device=src.device)
    nv.check(nv.lib().vt_pad_resize_area(src.data_ptr(), n, h, w, c, dst.data_ptr(), target_size, nv.current_stream_ptr()))
    return dst if batched else dst[0]


def pad_and_resize_for_siglip_batch(images, target_size: int = 384, device="cuda") -> Optional[torch.Tensor]:
    """scripts/utils_eef.py:5-41: the same operation on [N, H, W, C]."""
    return pad_and_resize_for_siglip(images, target_size, device)
