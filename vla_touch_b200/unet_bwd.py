"""Backward of the interpolant U-Net, part 1: the DATA gradients (dgrad) of every convolution as implicit GEMMs that run on
the forward kernel unchanged (`gemm_tc_kernel`, LINEAR epilogue) -- only the tap tables and the weight packing differ.

    Conv1d(k, s=1, p)            dX[u]    = sum_k W_k^T dY[u + p - k]                 one GEMM, taps shifted by (p - k)
    Conv1d(k3, s=2, p=1)         dX[2v]   = W_1^T dY[v]                               one GEMM per output phase, rows interleaved
    (Downsample1d :22-28)        dX[2v+1] = W_0^T dY[v+1] + W_2^T dY[v]
    ConvTranspose1d(k4, s2, p1)  dX[t]    = sum_k W_k dY[2t + k - 1]                  a stride-2 conv over dY: even / odd phase taps
    (Upsample1d :31-37)

(conditional_unet_1D.py:22-55).  W_k is the [C_out, C_in] slice of tap k (ConvTranspose: [C_in, C_out]).  The weight
gradients (wgrad: K = rows, MN-major operands), the fused GroupNorm+Mish+FiLM backward and the training-mode forward that keeps
the raw conv outputs are the round-2 kernels; `oracle/vt_oracle_bwd.py` is the checker all of them are held to.  These plan
builders are host logic: verified on the CPU by interpreting the descriptors (tests/test_plan_cpu.py) against that oracle.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from .unet import Mode, _View, _conv, _pack_conv


class DgradCtx:
    """The two attributes of UnetWeights the conv builder needs (mode, number of nets)."""

    def __init__(self, G: int, precise: bool):
        self.G = G
        self.mode = Mode(precise)


def pack_dgrad_conv(ws: Sequence[torch.Tensor], taps_k: Sequence[int], cout_pad: int, mode: Mode) -> torch.Tensor:
    """Forward Conv1d weights [C_out, C_in, K] of G nets -> dgrad B operand [G][round128(C_in)][len(taps_k) * cout_pad]:
    row ci holds, tap by tap (in the order of taps_k), W[:, ci, k] over the output channels (the GEMM's K dimension)."""
    return _pack_conv([w.permute(1, 0, 2)[:, :, list(taps_k)] for w in ws], cout_pad, mode)


def pack_dgrad_convT(ws: Sequence[torch.Tensor], taps_k: Sequence[int], cout_pad: int, mode: Mode) -> torch.Tensor:
    """ConvTranspose1d weights [C_in, C_out, K] are already [n = C_in][channel = C_out][tap]."""
    return _pack_conv([w[:, :, list(taps_k)] for w in ws], cout_pad, mode)


def conv_dgrad(plan, ctx: DgradCtx, B: int, dy: _View, dx: _View, ws: Sequence[torch.Tensor], pad: int, tag: str = "") -> List[torch.Tensor]:
    """dX of a stride-1 Conv1d(k, padding=pad).  dy: [G][B][T][C_out] view, dx: [G][B][T][C_in] view.  Returns the tensors the
    plan must keep alive (packed weights, zero bias)."""
    co, ci, K = ws[0].shape
    m = ctx.mode
    wd = plan.reg(pack_dgrad_conv(ws, range(K), dy.C, m).to(plan.device))
    zb = plan.reg(torch.zeros(ctx.G, wd.shape[1], dtype=torch.float32, device=plan.device))
    _conv(plan, ctx, B, dy, dx, wd, zb, taps=[(0, pad - k) for k in range(K)], cin_pad=dy.C, n=ci, t_out=dy.T,
          tag=tag or "conv.dgrad")
    return [wd, zb]


def downsample_dgrad(plan, ctx: DgradCtx, B: int, dy: _View, dx: _View, ws: Sequence[torch.Tensor], tag: str = "") -> List[torch.Tensor]:
    """dX of Conv1d(k3, stride 2, padding 1): dy has T/2 positions, dx T; one GEMM per parity of the dX position."""
    co, ci, K = ws[0].shape
    assert K == 3 and dx.T == 2 * dy.T
    m, keep = ctx.mode, []
    for ph, taps_k, taps in ((0, (1,), [(0, 0)]), (1, (0, 2), [(0, 1), (0, 0)])):
        wd = plan.reg(pack_dgrad_conv(ws, taps_k, dy.C, m).to(plan.device))
        zb = plan.reg(torch.zeros(ctx.G, wd.shape[1], dtype=torch.float32, device=plan.device))
        _conv(plan, ctx, B, dy, dx, wd, zb, taps=taps, cin_pad=dy.C, n=ci, t_out=dy.T, out_rows=(dx.T, 2, ph, dx.T),
              tag=(tag or "downsample.dgrad") + f".phase{ph}")
        keep += [wd, zb]
    return keep


def upsample_dgrad(plan, ctx: DgradCtx, B: int, dy: _View, dx: _View, ws: Sequence[torch.Tensor], tag: str = "") -> List[torch.Tensor]:
    """dX of ConvTranspose1d(k4, stride 2, padding 1): dy has 2T positions (read as even / odd phases), dx T."""
    ci, co, K = ws[0].shape
    assert K == 4 and dy.T == 2 * dx.T
    m = ctx.mode
    wd = plan.reg(pack_dgrad_convT(ws, range(K), dy.C, m).to(plan.device))
    zb = plan.reg(torch.zeros(ctx.G, wd.shape[1], dtype=torch.float32, device=plan.device))
    # dY[2t + k - 1]: k = 0 -> odd phase, index t-1;  k = 1 -> even, t;  k = 2 -> odd, t;  k = 3 -> even, t+1
    _conv(plan, ctx, B, dy, dx, wd, zb, taps=[(1, -1), (0, 0), (1, 0), (0, 1)], cin_pad=dy.C, n=ci, t_out=dx.T, phases=2,
          tag=tag or "upsample.dgrad")
    return [wd, zb]
