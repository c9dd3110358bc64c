"""DINOv2Encoder with the reference's interface (visual_encoder.py:9-106): a frozen DinoV2 ViT whose `forward(images)`
returns the CLS `pooler_output` [B, hidden].  Layout fix-ups, `/255 iff max > 1`, `ImageNet-normalise iff mean >= 0.5`
and the whole ViT run in one native program per input shape (vla_touch_b200.dino); nothing synchronises with the host."""
from __future__ import annotations

import os
import warnings
from typing import Dict, Optional

import numpy as np
import torch

from . import native as nv
from . import shapes as shp
from . import synthetic as syn
from .dino import DinoProgram, DinoWeights
from .params import ParamTree
from .plan import Plan


def _load_local_state_dict(model_name: str) -> Optional[Dict[str, torch.Tensor]]:
    """HF-format weights from a local directory / file, or from the local HF cache (no network access here)."""
    cands = []
    if os.path.isdir(model_name):
        cands += [os.path.join(model_name, f) for f in ("model.safetensors", "pytorch_model.bin")]
    elif os.path.isfile(model_name):
        cands.append(model_name)
    for f in cands:
        if not os.path.exists(f):
            continue
        if f.endswith(".safetensors"):
            from safetensors.torch import load_file
            return load_file(f)
        return torch.load(f, map_location="cpu", weights_only=True)
    try:
        from transformers import Dinov2Model
        return Dinov2Model.from_pretrained(model_name, local_files_only=True).state_dict()
    except Exception:
        return None


def prepare_images(images, device, keep_pinned_host: bool = False):
    """Host-side part of visual_encoder.py:66-75: numpy -> float/255, [B,T,H,W,C] -> [B*T,H,W,C]; returns
    (contiguous tensor on `device`, layout).  keep_pinned_host: a contiguous pinned host tensor is returned as it is, for
    the caller's own asynchronous upload."""
    if isinstance(images, np.ndarray):
        images = torch.from_numpy(images).float() / 255.0
    if images.dim() == 5:
        B, T, H, W, C = images.shape
        images = images.reshape(B * T, H, W, C)
        layout = nv.LAYOUT_BHWC
    elif images.dim() == 4 and images.shape[-1] == 3:
        layout = nv.LAYOUT_BHWC
    elif images.dim() == 4:
        layout = nv.LAYOUT_BCHW
    else:
        raise ValueError(f"unsupported image tensor shape {tuple(images.shape)}")
    if images.dtype not in (torch.uint8, torch.float32):
        images = images.float()
    if layout == nv.LAYOUT_BCHW and images.shape[1] != 3:
        raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the "
                         f"configuration. Expected 3 but got {images.shape[1]}.")
    if keep_pinned_host and images.device.type == "cpu" and images.is_pinned() and images.is_contiguous():
        return images, layout
    return images.to(device, non_blocking=True).contiguous(), layout


_side_streams: Dict[str, "torch.cuda.Stream"] = {}


def forward_two_cameras(enc: "DINOv2Encoder", images_cam1, images_cam2):
    """(enc.forward(cam1), enc.forward(cam2)) as the reference's encode_images calls them (bridge_controller.py:99-110).  When both
    batches are pinned host tensors (a DataLoader with pin_memory=True, controller_dataset.py:451-459) the second camera's
    host -> device copy runs on a side stream behind the first one's and overlaps the first camera's ViT forward; the values
    and the order of the two forward calls on the caller's stream are unchanged."""
    pinned = lambda t: torch.is_tensor(t) and t.device.type == "cpu" and t.is_pinned()
    if not (pinned(images_cam1) and pinned(images_cam2)) or not torch.cuda.is_available():
        return enc.forward(images_cam1), enc.forward(images_cam2)
    dev = torch.device(enc.device)
    main = torch.cuda.current_stream(dev)
    side = _side_streams.get(str(dev))
    if side is None:
        side = _side_streams[str(dev)] = torch.cuda.Stream(device=dev)
    d1 = images_cam1.to(dev, non_blocking=True)
    up1 = torch.cuda.Event()
    up1.record(main)
    side.wait_event(up1)                        # one copy at a time on the link: camera 2 follows camera 1
    with torch.cuda.stream(side):
        d2 = images_cam2.to(dev, non_blocking=True)
        up2 = torch.cuda.Event()
        up2.record(side)
    f1 = enc.forward(d1)
    main.wait_event(up2)
    d2.record_stream(main)                      # allocated on the side stream, consumed on the caller's
    return f1, enc.forward(d2)


class DINOv2Encoder:
    def __init__(self, model_name="facebook/dinov2-small", device="cuda", state_dict=None, precise: bool = False,
                 allow_synthetic_weights: bool = False, num_layers: Optional[int] = None):
        self.device = device
        self.precise = precise
        hidden, heads, layers = shp.dino_variant(model_name)
        self.patch_size = 14
        self.hidden_size = hidden
        self.num_heads = heads
        sd = state_dict if state_dict is not None else _load_local_state_dict(model_name)
        if sd is None:
            if not (allow_synthetic_weights or os.environ.get("VT_ALLOW_SYNTHETIC_DINO") == "1"):
                raise FileNotFoundError(
                    f"no local DinoV2 weights for '{model_name}' (no network): pass state_dict=..., a local directory, or "
                    "allow_synthetic_weights=True for seeded synthetic weights (benchmarks/tests)")
            warnings.warn("DINOv2Encoder: using seeded synthetic weights")
            sd = syn.synth_state_dict(shp.dinov2_shapes(hidden, num_layers or layers), 0, prefix="dino.")
        n_layers = 0
        while f"encoder.layer.{n_layers}.norm1.weight" in sd:
            n_layers += 1
        # frozen parameter container with the HF key names (reference: self.model, requires_grad False)
        self.model = ParamTree(shp.dinov2_shapes(hidden, n_layers, pos_grid=int((sd["embeddings.position_embeddings"].shape[1] - 1) ** 0.5)),
                               requires_grad=False)
        self.model.load_state_dict({k: v for k, v in sd.items() if k in self.model.state_dict()}, strict=False)
        self.model.to(device)
        self.model.eval()
        self._weights: Optional[DinoWeights] = None
        self._programs: Dict[tuple, tuple] = {}

    def weights(self) -> DinoWeights:
        if self._weights is None:
            self._weights = DinoWeights(self.model.state_dict(), self.num_heads, self.device, self.precise)
        return self._weights

    @torch.no_grad()
    def forward(self, images):
        images, layout = prepare_images(images, self.device)
        B = images.shape[0]
        H, W = (images.shape[1], images.shape[2]) if layout == nv.LAYOUT_BHWC else (images.shape[2], images.shape[3])
        key = (B, H, W, images.dtype, layout)
        if key not in self._programs:
            plan = Plan(self.device)
            self._programs[key] = (plan, DinoProgram(plan, self.weights(), 1, B, H, W, images.dtype, layout))
        plan, prog = self._programs[key]
        prog.img[0].copy_(images)
        plan.compile().run()
        return prog.feat[0].clone()

    __call__ = forward
