"""Backward of the interpolant U-Net's convolution blocks (row a10 of SURVEY 8: get_loss().backward()).

Part 1: the DATA gradients (dgrad) of every convolution as implicit GEMMs that run on the forward kernel unchanged (`gemm_tc_kernel`, LINEAR epilogue) -- only the tap tables and the weight packing differ.

    Conv1d(k, s=1, p)            dX[u]    = sum_k W_k^T dY[u + p - k]                 one GEMM, taps shifted by (p - k)
    Conv1d(k3, s=2, p=1)         dX[2v]   = W_1^T dY[v]                               one GEMM per output phase, rows interleaved
    (Downsample1d :22-28)        dX[2v+1] = W_0^T dY[v+1] + W_2^T dY[v]
    ConvTranspose1d(k4, s2, p1)  dX[t]    = sum_k W_k dY[2t + k - 1]                  a stride-2 conv over dY: even / odd phase taps
    (Upsample1d :31-37)

(conditional_unet_1D.py:22-55).  W_k is the [C_out, C_in] slice of tap k (ConvTranspose: [C_in, C_out]).

Part 2: the WEIGHT gradients as plain row-major GEMMs on the same kernel (K = B*T positions) over K-major operand copies made
by `tcol_kernel` (transposed im2col), the fused GroupNorm + Mish (+ FiLM) backward `gn_mish_bwd_kernel` working from the raw
conv output (which is recomputed with the LINEAR epilogue instead of being saved by the forward pass), bias gradients as
column sums, `conv_block_backward` (one whole Conv1dBlock) and `res_block_backward` (one ConditionalResidualBlock1D).
`unet_train.py` assembles them with the FiLM / time-MLP linears and the loss derivative into the full get_loss backward;
`oracle/vt_oracle_bwd.py` is the checker all of it is held to.  The plan builders are host logic: verified on the CPU by
interpreting the descriptors (tests/test_plan_cpu.py) against that oracle; tests/test_zz_backward_gpu.py runs the same plans
on the B200 (tests/bwd_cases.py holds the cases both share).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch

from . import native as nv
from .plan import VT_DT, linear_desc, ptr, round_up
from .unet import N_GROUPS, Mode, _View, _conv, _pack_conv, _pack_vec


class DgradCtx:
    """The two attributes of UnetWeights the conv builder needs (mode, number of nets)."""

    def __init__(self, G: int, precise: bool):
        self.G = G
        self.mode = Mode(precise)


class Ws(list):
    """The same parameter of the G nets (a list of tensors) that remembers its state-dict key, so that the operand copies
    packed from it at plan-build time can be re-packed after an optimizer step (`repack_all`)."""

    def __init__(self, sds: Sequence[dict], key: str, device):
        super().__init__(sd[key].detach().to(device) for sd in sds)
        self.key, self.device = key, device


def track(ctx, dest: torch.Tensor, ws, pack) -> torch.Tensor:
    """Remember how `dest` was packed from the parameter list `ws` (only when the parameters carry their key)."""
    key = getattr(ws, "key", None)
    if key is not None:
        if not hasattr(ctx, "repack"):
            ctx.repack = []
        ctx.repack.append((dest, lambda sds, key=key, dev=ws.device: pack(Ws(sds, key, dev))))
    return dest


def repack_all(ctx, sds: Sequence[dict]) -> None:
    """Re-pack every tracked operand copy from new parameter values into the same device tensors (addresses unchanged)."""
    for dest, fn in getattr(ctx, "repack", []):
        dest.copy_(fn(sds))


def pack_dgrad_conv(ws: Sequence[torch.Tensor], taps_k: Sequence[int], cout_pad: int, mode: Mode) -> torch.Tensor:
    """Forward Conv1d weights [C_out, C_in, K] of G nets -> dgrad B operand [G][round128(C_in)][len(taps_k) * cout_pad]:
    row ci holds, tap by tap (in the order of taps_k), W[:, ci, k] over the output channels (the GEMM's K dimension)."""
    return _pack_conv([torch.stack([w.permute(1, 0, 2)[:, :, k] for k in taps_k], dim=-1) for w in ws], cout_pad, mode)


def pack_dgrad_convT(ws: Sequence[torch.Tensor], taps_k: Sequence[int], cout_pad: int, mode: Mode) -> torch.Tensor:
    """ConvTranspose1d weights [C_in, C_out, K] are already [n = C_in][channel = C_out][tap]."""
    return _pack_conv([torch.stack([w[:, :, k] for k in taps_k], dim=-1) for w in ws], cout_pad, mode)


def conv_dgrad(plan, ctx: DgradCtx, B: int, dy: _View, dx: _View, ws: Sequence[torch.Tensor], pad: int, tag: str = "",
               res: Optional[_View] = None) -> List[torch.Tensor]:
    """dX of a stride-1 Conv1d(k, padding=pad).  dy: [G][B][T][C_out] view, dx: [G][B][T][C_in] view (bf16, or fp32 when the
    gradient feeds an elementwise backward); res: fp32 view added to the result (gradient fan-in: residual / skip paths).
    Returns the tensors the plan must keep alive (packed weights, zero bias)."""
    co, ci, K = ws[0].shape
    m = ctx.mode
    pk = lambda w: pack_dgrad_conv(w, range(K), dy.C, m).to(plan.device)
    wd = track(ctx, plan.reg(pk(ws)), ws, pk)
    zb = plan.reg(torch.zeros(ctx.G, wd.shape[1], dtype=torch.float32, device=plan.device))
    _conv(plan, ctx, B, dy, dx, wd, zb, taps=[(0, pad - k) for k in range(K)], cin_pad=dy.C, n=ci, t_out=dy.T, res=res,
          tag=tag or "conv.dgrad")
    return [wd, zb]


def downsample_dgrad(plan, ctx: DgradCtx, B: int, dy: _View, dx: _View, ws: Sequence[torch.Tensor], tag: str = "",
                     res: Optional[_View] = None) -> List[torch.Tensor]:
    """dX of Conv1d(k3, stride 2, padding 1): dy has T/2 positions, dx T; one GEMM per parity of the dX position."""
    co, ci, K = ws[0].shape
    assert K == 3 and dx.T == 2 * dy.T
    m, keep = ctx.mode, []
    for ph, taps_k, taps in ((0, (1,), [(0, 0)]), (1, (0, 2), [(0, 1), (0, 0)])):
        pk = lambda w, taps_k=taps_k: pack_dgrad_conv(w, taps_k, dy.C, m).to(plan.device)
        wd = track(ctx, plan.reg(pk(ws)), ws, pk)
        zb = plan.reg(torch.zeros(ctx.G, wd.shape[1], dtype=torch.float32, device=plan.device))
        _conv(plan, ctx, B, dy, dx, wd, zb, taps=taps, cin_pad=dy.C, n=ci, t_out=dy.T, out_rows=(dx.T, 2, ph, dx.T), res=res,
              tag=(tag or "downsample.dgrad") + f".phase{ph}")
        keep += [wd, zb]
    return keep


def upsample_dgrad(plan, ctx: DgradCtx, B: int, dy: _View, dx: _View, ws: Sequence[torch.Tensor], tag: str = "",
                   res: Optional[_View] = None) -> List[torch.Tensor]:
    """dX of ConvTranspose1d(k4, stride 2, padding 1): dy has 2T positions (read as even / odd phases), dx T."""
    ci, co, K = ws[0].shape
    assert K == 4 and dy.T == 2 * dx.T
    m = ctx.mode
    pk = lambda w: pack_dgrad_convT(w, range(K), dy.C, m).to(plan.device)
    wd = track(ctx, plan.reg(pk(ws)), ws, pk)
    zb = plan.reg(torch.zeros(ctx.G, wd.shape[1], dtype=torch.float32, device=plan.device))
    # dY[2t + k - 1]: k = 0 -> odd phase, index t-1;  k = 1 -> even, t;  k = 2 -> odd, t;  k = 3 -> even, t+1
    _conv(plan, ctx, B, dy, dx, wd, zb, taps=[(1, -1), (0, 0), (1, 0), (0, 1)], cin_pad=dy.C, n=ci, t_out=dx.T, phases=2, res=res,
          tag=tag or "upsample.dgrad")
    return [wd, zb]


# ------------------------------------------------------------------------------------------------
# part 2: weight gradients and the fused GroupNorm + Mish (+ FiLM) backward
# ------------------------------------------------------------------------------------------------
def _tcol(src: _View, B: int, G: int, out: torch.Tensor, *, tap_off: Sequence[int], stride: int, t_out: int, c_pad: int) -> nv.TcolDesc:
    d = nv.TcolDesc()
    d.src, d.src_dtype = ptr(src.t, src.c0), VT_DT[src.t.dtype]
    d.ld, d.sB, d.sG = src.ld, src.T * src.ld, 0 if src.shared else B * src.T * src.ld
    d.G, d.B, d.T_src, d.C, d.taps = G, B, src.T, src.C, len(tap_off)
    for i, o in enumerate(tap_off):
        d.tap_off[i] = o
    d.stride, d.t_out, d.out, d.c_pad = stride, t_out, ptr(out), c_pad
    d.k_ld, d.out_g = out.shape[-1], out.shape[1] * out.shape[2]
    return d


def conv_wgrad(plan, ctx: DgradCtx, B: int, rows_src: _View, cols_src: _View, *, tap_off: Sequence[int], stride: int = 1,
               t_out: int, tag: str = "conv.wgrad", split_k: Optional[int] = None, rows_t_memo: Optional[dict] = None) -> torch.Tensor:
    """Weight gradient of a convolution as ONE plain row-major GEMM on the forward kernel (K = B * t_out positions):

        dW[g][r][tap * c_pad + c] = sum_{b, t < t_out} rows_src[g][b][t][r] * cols_src[g][b][t * stride + tap_off[tap]][c]

    Conv1d(k, s, p):         rows_src = dY (one tap, offset 0), cols_src = X  (tap_off[k] = k - p, stride s)  -> [C_out][k][C_in]
    ConvTranspose1d(4,2,1):  rows_src = X,                      cols_src = dY (tap_off[k] = k - 1, stride 2)  -> [C_in][k][C_out]
    (conv1d_bwd / convT1d_bwd of oracle/vt_oracle_bwd.py; conditional_unet_1D.py:22-55).  Both operands are first copied
    into K-major (transposed) bf16 buffers by tcol_kernel.  Returns dW fp32 [G][rows][taps * c_pad]: the forward packing of
    the weight (`_pack_conv`), so `unpack_wgrad` is a view.

    split_k = S > 1 (default: env VT_WGRAD_SPLITK, 1): the batch is cut into S slices that run as S x G GEMM groups (S times the
    tiles: a 256 -> 256 k5 conv has only 5 x G tile pairs for 148 SMs otherwise) and one column-sum launch adds the partial
    gradients.  Needs B % S == 0 and per-net (not shared) operands; otherwise it silently stays at 1."""
    G = ctx.G
    assert not ctx.mode.precise, "training runs in the bf16 mode"
    R, c_pad, taps = rows_src.C, cols_src.C, len(tap_off)
    n = taps * c_pad
    # MN-major path (csrc/vt_wgrad.cuh): both operands are read by TMA straight from the channels-last activations -- no
    # transposed copies.  Needs bf16 per-net operands, 64-column tap blocks and a K tile of whole samples (t_out | 64).
    direct = (os.environ.get("VT_WGRAD_DIRECT", "1") != "0" and rows_src.t.dtype == torch.bfloat16 and cols_src.t.dtype == torch.bfloat16
              and not rows_src.shared and not cols_src.shared and 64 % t_out == 0 and c_pad % 64 == 0 and stride in (1, 2)
              and (stride == 1 or cols_src.T % 2 == 0))
    if direct:
        S = split_k if split_k is not None else int(os.environ.get("VT_WGRAD_SPLITK", "0"))
        if S <= 0:      # automatic: enough tiles for the 74 CTA pairs, at least 256 positions per slice
            tiles = G * ((R + 255) // 256) * ((n + 255) // 256)
            S = 1
            while tiles * S < 64 and S < 16 and B % (2 * S) == 0 and (B // (2 * S)) * t_out >= 256:
                S *= 2
        if B % S != 0:
            S = 1
        Gs, Bs = G * S, B // S
        nm = tag + f"#{len(plan)}"
        dw = plan.buf(nm + ".dw", (G, R, n), torch.float32, arena="grads")
        part = dw if S == 1 else plan.buf(nm + ".dw_part", (Gs, R, n), torch.float32)
        d = nv.WgradDesc()
        d.rows, d.rows_C, d.rows_P, d.rows_T = ptr(rows_src.t, rows_src.c0), rows_src.ld - rows_src.c0, 1, rows_src.T
        d.rows_ld, d.rows_sB, d.rows_sG, d.rows_p, d.rows_t = rows_src.ld, rows_src.T * rows_src.ld, Bs * rows_src.T * rows_src.ld, 0, 0
        d.cols, d.cols_C, d.cols_P, d.cols_T = ptr(cols_src.t, cols_src.c0), cols_src.ld - cols_src.c0, stride, cols_src.T // stride
        d.cols_ld, d.cols_sB, d.cols_sG = cols_src.ld, cols_src.T * cols_src.ld, Bs * cols_src.T * cols_src.ld
        d.taps, d.c_pad = taps, c_pad
        for i, o in enumerate(tap_off):
            d.tap_p[i] = o % stride
            d.tap_t[i] = (o - o % stride) // stride
        d.G, d.B, d.t_out, d.R, d.out, d.ldc, d.out_g = Gs, Bs, t_out, R, ptr(part), n, R * n
        plan.add(d, tag + ".gemm(MN-major)")
        if S > 1:
            c = nv.ColsumDesc()
            c.x, c.ld, c.x_g, c.G, c.rows, c.C, c.out, c.out_ld = ptr(part), R * n, S * R * n, G, S, R * n, ptr(dw), R * n
            plan.add(c, tag + ".splitk_sum")
        return dw
    S = split_k if split_k is not None else int(os.environ.get("VT_WGRAD_SPLITK", "0"))
    if S <= 0:          # automatic (as on the MN-major path): enough tiles for the 74 CTA pairs, at least 256 positions per slice
        bn_ = 256 if n % 256 == 0 else 128
        tiles = G * ((R + 255) // 256) * ((n + bn_ - 1) // bn_)
        S = 1
        while tiles * S < 64 and S < 16 and B % (2 * S) == 0 and (B // (2 * S)) * t_out >= 256:
            S *= 2
    if S < 1 or B % S != 0 or rows_src.shared or cols_src.shared:
        S = 1
    Gs, Bs = G * S, B // S                       # split-K: slice s of net g = samples [s Bs, (s+1) Bs) -> group g S + s
    kp = round_up(Bs * t_out, 64)
    bn = 256 if n % 256 == 0 else 128
    n_pad = round_up(n, bn)
    nm = tag + f"#{len(plan)}"
    # rows_t_memo (opt-in, caller's promise that rows_src is not rewritten between the calls that share the dict): consecutive
    # weight gradients with the SAME rows operand (d gates of an LSTM layer feeds d W_ih and d W_hh) share one transposed copy
    key = (ptr(rows_src.t, rows_src.c0), rows_src.ld, rows_src.T, Bs, Gs, t_out, R, kp)
    rT = rows_t_memo.get(key) if rows_t_memo is not None else None
    new_rows = rT is None
    if new_rows:
        rT = plan.buf(nm + ".rowsT", (Gs, round_up(R, 128), kp), torch.bfloat16)
        if rows_t_memo is not None:
            rows_t_memo[key] = rT
    cT = plan.buf(nm + ".colsT", (Gs, n_pad, kp), torch.bfloat16)
    dw = plan.buf(nm + ".dw", (G, R, n), torch.float32, arena="grads")
    part = dw if S == 1 else plan.buf(nm + ".dw_part", (Gs, R, n), torch.float32)
    if new_rows:
        plan.add(_tcol(rows_src, Bs, Gs, rT, tap_off=[0], stride=1, t_out=t_out, c_pad=R), tag + ".rowsT")
    plan.add(_tcol(cols_src, Bs, Gs, cT, tap_off=list(tap_off), stride=stride, t_out=t_out, c_pad=c_pad), tag + ".colsT")
    plan.add(linear_desc(a=rT, rows=R, k=kp, a_ld=kp, w=cT, n=n, n_pad=n_pad, w_ld=kp, out=part, ldc=n, G=Gs, a_G=Gs,
                         a_sG=rT.shape[1] * kp, out_g=R * n, bn=bn), tag + ".gemm")
    if S > 1:                                    # dW[g] = sum_s part[g S + s]: a column sum over S rows of R n columns
        d = nv.ColsumDesc()
        d.x, d.ld, d.x_g, d.G, d.rows, d.C, d.out, d.out_ld = ptr(part), R * n, S * R * n, G, S, R * n, ptr(dw), R * n
        plan.add(d, tag + ".splitk_sum")
    return dw


def unpack_wgrad(dw: torch.Tensor, c_valid: int, taps: int) -> torch.Tensor:
    """[G][R][taps * c_pad] -> the nn.Module layout: Conv1d [G][C_out][C_in][k]; ConvTranspose1d [G][C_in][C_out][k]
    (both are [rows][channels][k] in the respective layer's own weight order)."""
    G, R, n = dw.shape
    return dw.view(G, R, taps, n // taps)[:, :, :, :c_valid].permute(0, 1, 3, 2)


def as_view(t, T: int) -> _View:
    """fp32 gradient given as a compact [G][B][T][C] tensor or as a channel window (_View) of a wider buffer."""
    return t if isinstance(t, _View) else _View(t, T, t.shape[-1])


def gn_mish_backward(plan, ctx: DgradCtx, B: int, T: int, C: int, raw: torch.Tensor, dout: torch.Tensor, gamma: torch.Tensor,
                     beta: torch.Tensor, *, film=None, tag: str = "gn.bwd"):
    """GroupNorm(8) + Mish (+ FiLM) backward of one Conv1dBlock for G nets (gn_mish_bwd + the FiLM lines of _res_block_bwd in
    oracle/vt_oracle_bwd.py).  raw fp32 [G][B][T][C], dout fp32 tensor or channel window of the same extent; gamma / beta fp32 [G][C];
    film = (film table [G][B][ld], d film table [G][B][ld], column offset) or None.
    Returns (draw bf16 [G][B][T][C], dgamma, dbeta, dbias fp32 [G][C])."""
    G = ctx.G
    nm = tag + f"#{len(plan)}"
    draw = plan.buf(nm + ".draw", (G, B, T, C), torch.bfloat16)
    part = plan.buf(nm + ".part", (G, B, 3, C), torch.float32)
    dg, db, dbias = (plan.buf(nm + "." + k, (G, C), torch.float32, arena="grads") for k in ("dgamma", "dbeta", "dbias"))
    d = nv.GnbwdDesc()
    dv = as_view(dout, T)
    d.raw, d.dout, d.dout_ld, d.dout_g = ptr(raw), ptr(dv.t, dv.c0), dv.ld, B * T * dv.ld
    d.gamma, d.beta, d.p_ld = ptr(gamma), ptr(beta), gamma.shape[-1]
    if film is not None:
        ft, dft, off = film
        d.film, d.dfilm, d.film_g, d.film_ld, d.film_off = ptr(ft), ptr(dft), B * ft.shape[-1], ft.shape[-1], off
    d.draw, d.part, d.dgamma, d.dbeta, d.dbias = ptr(draw), ptr(part), ptr(dg), ptr(db), ptr(dbias)
    d.G, d.B, d.T, d.C, d.groups, d.eps = G, B, T, C, N_GROUPS, 1e-5
    plan.add(d, tag)
    return draw, dg, db, dbias


def conv_block_backward(plan, ctx: DgradCtx, B: int, x: _View, ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor],
                        gammas: Sequence[torch.Tensor], betas: Sequence[torch.Tensor], dout: torch.Tensor, dx: _View, *,
                        film=None, res: Optional[_View] = None, raw: Optional[torch.Tensor] = None, tag: str = "block"):
    """Backward of Conv1dBlock = Conv1d(k, padding k//2) -> GroupNorm(8) -> Mish [-> FiLM] (conditional_unet_1D.py:40-55,
    97-102) for G nets.  raw = fp32 [G][B][T][C_out] conv + bias saved by the training forward (vt_gemm_desc.raw_out); None: it is
    RECOMPUTED here with the forward kernel (LINEAR epilogue, fp32) -- bit-identical, one more GEMM.  Then gn_mish_backward ->
    conv_wgrad -> conv_dgrad.
    x: bf16 [G][B][T][C_in] view (block input), dout fp32 [G][B][T][C_out], dx: bf16 or fp32 view receiving d x (+ res).
    Returns dict(dw [G][C_out][k * cin_pad], dbias, dgamma, dbeta [G][C_out])."""
    co, ci, K = ws[0].shape
    m, T = ctx.mode, x.T
    dev = plan.device
    pc = lambda v: _pack_vec([t.to(dev) for t in v], co)
    gm = track(ctx, plan.reg(pc(gammas)), gammas, pc)
    bt = track(ctx, plan.reg(pc(betas)), betas, pc)
    if raw is None:
        pw = lambda w: _pack_conv([t.to(dev) for t in w], x.C, m)
        wf = track(ctx, plan.reg(pw(ws)), ws, pw)
        pb = lambda v: _pack_vec([t.to(dev) for t in v], wf.shape[1])
        bf = track(ctx, plan.reg(pb(bs)), bs, pb)
        raw = plan.buf(tag + f"#{len(plan)}.raw", (ctx.G, B, T, co), torch.float32)
        _conv(plan, ctx, B, x, None, wf, bf, taps=[(0, k - K // 2) for k in range(K)], cin_pad=x.C, n=co, t_out=T, out_f32=raw,
              tag=tag + ".conv_raw(recompute)")
    draw, dg, db, dbias = gn_mish_backward(plan, ctx, B, T, co, raw, dout, gm, bt, film=film, tag=tag + ".gn+mish.bwd")
    vy = _View(draw, T, co)
    dw = conv_wgrad(plan, ctx, B, vy, x, tap_off=[k - K // 2 for k in range(K)], t_out=T, tag=tag + ".wgrad")
    if dx is not None:                     # the first block's input gradient (d sample) is not needed by training
        conv_dgrad(plan, ctx, B, vy, dx, ws, pad=K // 2, tag=tag + ".dgrad", res=res)
    return dict(dw=dw, dbias=dbias, dgamma=dg, dbeta=db, raw=raw, draw=draw)


def cast_bf16(plan, G: int, B: int, src, T: int, tag: str, c_pad: Optional[int] = None) -> _View:
    """fp32 gradient (tensor or channel window) -> compact bf16 copy [G][B][T][c_pad] (zero-filled beyond C): the GEMM operand
    of the dgrad / wgrad that consume it."""
    v = as_view(src, T)
    cp = c_pad or v.C
    out = plan.buf(tag + f"#{len(plan)}.bf16", (G, B, T, cp), torch.bfloat16)
    d = nv.PackDesc()
    d.src, d.src_ld, d.rows, d.cols, d.act = ptr(v.t, v.c0), v.ld, G * B * T, v.C, nv.ACT_NONE
    d.out, d.out_dtype, d.out_ld, d.dst_c0, d.out_plane, d.zero_to = ptr(out), nv.VT_BF16, cp, 0, 0, cp if cp > v.C else 0
    plan.add(d, tag)
    return _View(out, T, cp)


def colsum(plan, G: int, B: int, x, T: int, tag: str) -> torch.Tensor:
    """fp32 [G][B*T][C] (tensor or channel window) -> [G][C]: the bias gradient of a convolution without GroupNorm."""
    v = as_view(x, T)
    out = plan.buf(tag + f"#{len(plan)}.out", (G, v.C), torch.float32, arena="grads")
    d = nv.ColsumDesc()
    d.x, d.ld, d.x_g, d.G, d.rows, d.C, d.out, d.out_ld = ptr(v.t, v.c0), v.ld, B * T * v.ld, G, B * T, v.C, ptr(out), v.C
    plan.add(d, tag)
    return out


def res_block_backward(plan, ctx: DgradCtx, B: int, sds: Sequence[dict], pfx: str, x: _View, y1: _View, dout: torch.Tensor,
                       dx: _View, film, raws=(None, None), tag: str = "") -> dict:
    """Backward of ConditionalResidualBlock1D (conditional_unet_1D.py:58-105; `_res_block_bwd` of the oracle) for G nets.

    sds: the nets' state dicts, pfx the block's key prefix; x: bf16 view of the block input, y1: bf16 view of the FiLM output
    (= input of blocks[1], kept by the training forward), dout fp32 [G][B][T][C_out], dx: fp32 view receiving d x (None: skip),
    film = (FiLM table, d FiLM table, column offset of this block).  The gradient of the cond_encoder Linear is taken from the
    d FiLM table for all 12 blocks at once by the caller.  Returns {reference parameter key suffix: gradient buffer}."""
    tag = tag or pfx
    g = lambda k: Ws(sds, pfx + k, plan.device)
    T = x.T
    dout = as_view(dout, T)
    co = dout.C
    out = {}
    # blocks[1]: Conv1d -> GN -> Mish, no FiLM; its d x is d y1 (fp32: it is the d out of blocks[0]'s elementwise backward)
    dy1 = plan.buf(tag + f"#{len(plan)}.dy1", (ctx.G, B, T, co), torch.float32)
    b1 = conv_block_backward(plan, ctx, B, y1, g("blocks.1.block.0.weight"), g("blocks.1.block.0.bias"),
                             g("blocks.1.block.1.weight"), g("blocks.1.block.1.bias"), dout, _View(dy1, T, co), raw=raws[1],
                             tag=tag + "blocks.1")
    # residual path: d x += dout (identity) or W_r^T dout (1x1 conv)
    res = dout
    if pfx + "residual_conv.weight" in sds[0]:
        wr = g("residual_conv.weight")
        dob = cast_bf16(plan, ctx.G, B, dout, T, tag + "dout.bf16")
        if dx is not None:
            dxr = plan.buf(tag + f"#{len(plan)}.dxr", (ctx.G, B, T, dx.C), torch.float32)
            conv_dgrad(plan, ctx, B, dob, _View(dxr, T, dx.C), wr, pad=0, tag=tag + "residual_conv.dgrad")
            res = _View(dxr, T, dx.C)
        out["residual_conv.weight"] = (conv_wgrad(plan, ctx, B, dob, x, tap_off=[0], t_out=T, tag=tag + "residual_conv.wgrad"), 1)
        out["residual_conv.bias"] = colsum(plan, ctx.G, B, dout, T, tag + "residual_conv.dbias")
    b0 = conv_block_backward(plan, ctx, B, x, g("blocks.0.block.0.weight"), g("blocks.0.block.0.bias"),
                             g("blocks.0.block.1.weight"), g("blocks.0.block.1.bias"), dy1, dx, film=film, res=res, raw=raws[0],
                             tag=tag + "blocks.0")
    for i, b in ((0, b0), (1, b1)):
        K = sds[0][pfx + f"blocks.{i}.block.0.weight"].shape[-1]
        out[f"blocks.{i}.block.0.weight"] = (b["dw"], K)
        out[f"blocks.{i}.block.0.bias"] = b["dbias"]
        out[f"blocks.{i}.block.1.weight"] = b["dgamma"]
        out[f"blocks.{i}.block.1.bias"] = b["dbeta"]
    return out
