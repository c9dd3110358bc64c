"""InterpolantsConditionalUnet1D: the three conditional 1-D U-Nets b_net, v_net, s_net of the stochastic interpolant
(reference: bridge/networks/conditional_unet_1D_si.py:4-50, conditional_unet_1D.py:108-247).

This class is the parameter container (reference state_dict keys and parameter order, so `bridge_model.pt` and the
torch_ema shadow list interchange); evaluating a net runs the grouped tcgen05 implicit-GEMM program of
vla_touch_b200.unet."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from ... import shapes as shp
from ...params import ParamTree, sub_state_dict, torch_default_init
from ...unet import UnetProgram


class DiffusionConditionalUnet1D(ParamTree):
    """One net (conditional_unet_1D.py:108-247).  forward(sample [B,T,A], timestep, global_cond [B,cond]) -> [B,T,A]."""

    def __init__(self, input_dim: int, global_cond_dim: int, seed: int = 0, precise: bool = False):
        super().__init__(shp.unet_shapes(input_dim, global_cond_dim), init=torch_default_init(seed))
        self.input_dim, self.global_cond_dim, self.precise = input_dim, global_cond_dim, precise
        self._progs: Dict[tuple, UnetProgram] = {}

    @torch.no_grad()
    def forward(self, sample: torch.Tensor, timestep, global_cond: Optional[torch.Tensor] = None) -> torch.Tensor:
        if global_cond is None:
            raise NotImplementedError("global_cond=None is not used by the reference controller")
        B, T, A = sample.shape
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.float32, device=sample.device)
        timestep = timestep.to(sample.device, torch.float32).reshape(-1).expand(B)
        version = sum(p._version for p in self.parameters())
        key = (B, T, str(sample.device))
        ent = self._progs.get(key)
        if ent is None or ent[1] != version:
            ent = (UnetProgram([{k: v.detach() for k, v in self.state_dict().items()}], A, B, T, sample.device, self.precise), version)
            self._progs[key] = ent
        return ent[0](sample.float(), timestep, global_cond.float())[0].clone()


class InterpolantsConditionalUnet1D(nn.Module):
    def __init__(self, input_dim: int, global_cond_dim: int, precise: bool = False):
        super().__init__()
        self.b_net = DiffusionConditionalUnet1D(input_dim, global_cond_dim, seed=1, precise=precise)
        self.v_net = DiffusionConditionalUnet1D(input_dim, global_cond_dim, seed=2, precise=precise)
        self.s_net = DiffusionConditionalUnet1D(input_dim, global_cond_dim, seed=3, precise=precise)
        self.input_dim, self.global_cond_dim = input_dim, global_cond_dim

    def forward(self, *a, **k):
        raise RuntimeError("call b_net / v_net / s_net")
