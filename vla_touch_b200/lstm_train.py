"""Training plans of the residual-LSTM controller (row a12 of SURVEY 8, lstm_train.py:70-130: get_loss(...).backward()).

First slice: one nn.LSTM layer (lstm_step_controller.py:66-73,196-204) forward-for-training and back-propagation through time.

    forward   xw = W_ih x + b_ih + b_hh for all steps (one GEMM)  ->  lstm_seq_train_kernel (recurrence; keeps gates and c)
    backward  lstm_bwd_kernel: the sequential part (d gates of every step; dh_{t-1} = d gates_t W_hh)
              d W_ih = d gates^T x,  d W_hh = d gates^T h_{t-1}   two weight-gradient GEMMs over all B*T rows (unet_bwd.conv_wgrad:
                                                                  h_{t-1} is the hidden output read with tap offset -1)
              d b_ih = d b_hh = column sum of d gates;  d x = d gates W_ih   (one GEMM)

`lstm_loss_backward` of oracle/vt_oracle_bwd.py (pinned to the reference's gradient digests) is the checker.  The head
(Linear, LayerNorm, GELU, Linear), the force encoder and inter-layer dropout are not built yet.  Written after the round's GPU
budget ended: the two kernels have been compiled for sm_100a and checked on the CPU descriptor interpreter only.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import native as nv
from . import unet_bwd as ub
from .plan import Plan, linear_desc, ptr
from .unet import _View

H_LSTM = 256


class LstmLayerTrain:
    """One LSTM layer of hidden size 256 on (B, T): training forward + BPTT as plan ops.

    x: bf16 [B][T][k_pad] input (zero padded to a multiple of 64), weights in nn.LSTM layout (weight_ih [4H, k_in],
    weight_hh [4H, H], bias_ih, bias_hh [4H]).  After `forward(plan)`: self.y (bf16 [B][T][H]), self.gates, self.c.
    After `backward(plan, dy)`: self.grads = {weight_ih, weight_hh, bias_ih, bias_hh} and self.dx (fp32 [B][T][k_pad])."""

    def __init__(self, plan: Plan, x: torch.Tensor, k_in: int, w_ih: torch.Tensor, w_hh: torch.Tensor, b_ih: torch.Tensor,
                 b_hh: torch.Tensor, B: int, T: int, tag: str = "lstm.l0"):
        H, dev, f32, bf = H_LSTM, plan.device, torch.float32, torch.bfloat16
        assert w_hh.shape == (4 * H, H) and x.dtype == bf and x.shape[-1] % 64 == 0
        self.plan, self.x, self.k_in, self.k_pad, self.B, self.T, self.tag = plan, x, k_in, x.shape[-1], B, T, tag
        wp = torch.zeros(4 * H, self.k_pad, device=dev)
        wp[:, :k_in] = w_ih.detach().to(dev, f32)
        self.w_ih = plan.reg(wp.to(bf).contiguous())                                   # forward operand [4H][k_pad]
        self.w_ih_t = plan.reg(wp.t().contiguous().to(bf))                             # d x operand     [k_pad][4H]
        self.b_sum = plan.reg((b_ih + b_hh).detach().to(dev, f32).contiguous())
        self.w_hh = plan.reg(w_hh.detach().to(dev, f32).contiguous())                  # [4H][H]  (backward recurrence)
        self.w_hh_t = plan.reg(w_hh.detach().to(dev, f32).t().contiguous())            # [H][4H]  (forward recurrence)
        R = B * T
        self.xw = plan.buf(f"{tag}.xw", (R, 4 * H), f32)
        self.y = plan.buf(f"{tag}.y", (B, T, H), bf)
        self.gates = plan.buf(f"{tag}.gates", (B, T, 4 * H), f32)
        self.c = plan.buf(f"{tag}.c", (B, T, H), f32)
        self.grads: Dict[str, torch.Tensor] = {}
        self.dx: Optional[torch.Tensor] = None

    def forward(self) -> torch.Tensor:
        p, H, R = self.plan, H_LSTM, self.B * self.T
        p.add(linear_desc(a=self.x, rows=R, k=self.k_pad, a_ld=self.k_pad, w=self.w_ih, n=4 * H, n_pad=4 * H, w_ld=self.k_pad,
                          out=self.xw, ldc=4 * H, bias=self.b_sum), f"{self.tag}.input_proj")
        d = nv.LstmTrainDesc()
        d.xw, d.w_hh, d.y, d.y_dtype, d.y_ld = ptr(self.xw), ptr(self.w_hh_t), ptr(self.y), nv.VT_BF16, H
        d.gates, d.c, d.B, d.T, d.H = ptr(self.gates), ptr(self.c), self.B, self.T, H
        p.add(d, f"{self.tag}.recurrence(train)")
        return self.y

    def backward(self, dy: torch.Tensor, need_dx: bool = True) -> Optional[torch.Tensor]:
        """dy: fp32 [B][T][H] gradient of the hidden outputs."""
        p, H, B, T, tag = self.plan, H_LSTM, self.B, self.T, self.tag
        R = B * T
        dg = p.buf(f"{tag}.dgates", (B, T, 4 * H), torch.float32)
        d = nv.LstmBwdDesc()
        d.gates, d.c, d.dy, d.dy_ld, d.w_hh, d.dgates = ptr(self.gates), ptr(self.c), ptr(dy), dy.shape[-1], ptr(self.w_hh), ptr(dg)
        d.B, d.T, d.H = B, T, H
        p.add(d, f"{tag}.bptt")
        ctx = ub.DgradCtx(1, precise=False)
        V = _View
        dgv = V(dg.view(1, B, T, 4 * H), T, 4 * H)                                     # fp32 source: tcol converts to bf16
        dw_ih = ub.conv_wgrad(p, ctx, B, dgv, V(self.x.view(1, B, T, self.k_pad), T, self.k_pad), tap_off=[0], t_out=T,
                              tag=f"{tag}.weight_ih.wgrad")
        dw_hh = ub.conv_wgrad(p, ctx, B, dgv, V(self.y.view(1, B, T, H), T, H), tap_off=[-1], t_out=T,
                              tag=f"{tag}.weight_hh.wgrad")                            # h_{t-1}: zero at t = 0
        db = ub.colsum(p, 1, B, dgv, T, f"{tag}.bias.colsum")
        self.grads = {"weight_ih": dw_ih[0, :, : self.k_in], "weight_hh": dw_hh[0], "bias_ih": db[0], "bias_hh": db[0]}
        if need_dx:
            dgb = ub.cast_bf16(p, 1, B, dgv, T, f"{tag}.dgates.bf16")
            self.dx = p.buf(f"{tag}.dx", (B, T, self.k_pad), torch.float32)
            p.add(linear_desc(a=dgb.t, rows=R, k=4 * H, a_ld=4 * H, w=self.w_ih_t, n=self.k_pad, n_pad=self.k_pad,
                              w_ld=4 * H, out=self.dx, ldc=self.k_pad), f"{tag}.dx")
        return self.dx


def lstm_layers_train(plan: Plan, x: torch.Tensor, k_in: int, lstm_sd: Dict[str, torch.Tensor], B: int, T: int,
                      num_layers: int = 2) -> Sequence[LstmLayerTrain]:
    """The stacked layers of nn.LSTM(num_layers) in eval-equivalent training (no inter-layer dropout yet): forward ops of all
    layers are appended to `plan`; call `.backward(dy)` on them in reverse order."""
    layers = []
    for l in range(num_layers):
        lay = LstmLayerTrain(plan, x, k_in, lstm_sd[f"weight_ih_l{l}"], lstm_sd[f"weight_hh_l{l}"], lstm_sd[f"bias_ih_l{l}"],
                             lstm_sd[f"bias_hh_l{l}"], B, T, tag=f"lstm.l{l}")
        x = lay.forward()
        k_in = H_LSTM
        layers.append(lay)
    return layers
