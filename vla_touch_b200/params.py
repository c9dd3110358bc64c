"""Parameter containers whose state_dict keys and parameter order equal the reference modules', so that checkpoints
(`controller.pt`, `bridge_model.pt`, `tactile_controller.pt`) and optimizers interchange with the reference
(SURVEY.md App. C).  They hold weights only: the arithmetic runs in libvt_b200.so."""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterable, Tuple

import torch
import torch.nn as nn


class ParamTree(nn.Module):
    """nn.Module tree built from dotted names -> shapes, registered in the given order."""

    def __init__(self, shapes: "OrderedDict[str, Tuple[int, ...]]", requires_grad: bool = True, init=None):
        super().__init__()
        for name, shape in shapes.items():
            node = self
            parts = name.split(".")
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, _Node())
                node = node._modules[p]
            t = torch.zeros(tuple(shape)) if init is None else init(name, tuple(shape))
            node.register_parameter(parts[-1], nn.Parameter(t, requires_grad=requires_grad))

    def forward(self, *a, **k):
        raise RuntimeError("parameter container: the computation runs in the vla_touch_b200 native engine")


class _Node(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("parameter container: the computation runs in the vla_touch_b200 native engine")


def torch_default_init(seed: int):
    """PyTorch-like default init (uniform +-1/sqrt(fan_in) for matrices/biases, ones/zeros for norm affine)."""
    g = torch.Generator().manual_seed(seed)

    def init(name: str, shape):
        leaf = name.rsplit(".", 1)[-1]
        if len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            b = 1.0 / max(fan_in, 1) ** 0.5
            return (torch.rand(shape, generator=g) * 2 - 1) * b
        if "block.1." in name or ".norm" in name or name.startswith("1.") and leaf == "weight":
            return torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return init


def sub_state_dict(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
