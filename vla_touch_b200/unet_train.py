"""Training-mode forward and the explicit backward of the interpolant U-Nets (row a10 of SURVEY 8:
StochasticInterpolants.get_loss(...).backward(), bridge_model.py:220-246, through DiffusionConditionalUnet1D.forward,
bridge/networks/conditional_unet_1D.py:194-247) as ONE plan for the G = 3 nets (b_net, v_net, s_net).

Forward: the same fused kernels as inference (implicit-GEMM conv + GroupNorm + Mish + FiLM + residual in the epilogue), but
every block writes its own buffers, because the backward reads each block's input and FiLM output again.  The raw conv outputs
(conv + bias, what GroupNorm sees) are stored as fp32 by the forward GEMM's GroupNorm epilogue itself (vt_gemm_desc.raw_out), as
autograd would keep them; VT_TRAIN_RECOMPUTE=1 restores the older form in which the backward recomputes them
(unet_bwd.conv_block_backward without `raw`).

Backward, in reverse execution order (`unet_backward` of oracle/vt_oracle_bwd.py is the checker):
  final 1x1 conv, final Conv1dBlock, then per level: ConvTranspose1d (up) / 12 x ConditionalResidualBlock1D / Conv1d stride 2
  (down); torch.cat splits the gradient into channel windows (free: consumers read windows of the fp32 gradient buffer), the
  skip connections sum gradients through the dgrad GEMM's residual input or one `ewise` add; the unused level-0 skip
   gets no gradient.  Then the FiLM Linear of all 12 blocks at once (their d scale / d shift rows are columns
  of one [B][11264] table), Mish'(gf), the time-embedding MLP, and d global_cond.
Gradient activations are fp32 where an elementwise backward consumes them and are cast to bf16 copies where a GEMM does.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import native as nv
from . import unet_bwd as ub
from .plan import Plan, linear_desc, ptr
from .unet import DOWN_DIMS, DSED, FILM_ROWS, Mode, UnetWeights, _View, _conv, _tembed, block_names

SD = Dict[str, torch.Tensor]


def _operand_dtype():
    """bf16, or fp32 while the packing functions are traced with index-valued tensors (GatherRepack)"""
    return torch.float32 if Mode.TRACE else torch.bfloat16
K5 = [(0, dt) for dt in (-2, -1, 0, 1, 2)]
K1 = [(0, 0)]


class _Block:
    def __init__(self, pfx: str, x: _View, y1: _View, out: _View, r: Optional[_View], raw0=None, raw1=None):
        self.pfx, self.x, self.y1, self.out, self.r = pfx, x, y1, out, r
        self.raw0, self.raw1 = raw0, raw1          # fp32 [G][B][T][C_out]: conv + bias of blocks[0] / blocks[1] (None: recomputed)


class UnetTrainBuffers:
    """Activations of one training forward of G nets on (B, T): one buffer per block output / FiLM output."""

    def __init__(self, plan: Plan, W: UnetWeights, B: int, T: int, tag: str = "tr", save_raw: Optional[bool] = None):
        if save_raw is None:
            save_raw = os.environ.get("VT_TRAIN_RECOMPUTE") != "1"
        if T % 4 != 0 or T > 128 or T < 4:
            raise ValueError(f"horizon T={T} must be a multiple of 4 in [4, 128]")
        if W.mode.precise:
            raise ValueError("the training plans run in the bf16 mode")
        G = W.G
        self.B, self.T, self.G = B, T, G
        T0, T1, T2 = self.Ts = (T, T // 2, T // 4)
        d0, d1, d2 = DOWN_DIMS
        bf = torch.bfloat16
        mk = lambda name, t, c: plan.buf(f"{tag}.{name}", (G, B, t, c), bf)
        self.xpad = plan.buf(f"{tag}.xpad", (B, T0, W.cin0), bf)
        self.out = plan.buf(f"{tag}.vs", (G, B, T0, W.A), torch.float32)
        A0, A1, D1, A2 = mk("A0", T0, d0), mk("A1", T0, d0), mk("D1", T1, d0), mk("A2", T1, d1)
        cat1, D2, A4, cat0 = mk("cat1", T1, 2 * d1), mk("D2", T2, d1), mk("A4", T2, d2), mk("cat0", T2, 2 * d2)
        A6, A8, A9 = mk("A6", T2, d2), mk("A8", T2, d1), mk("A9", T2, d1)
        A10, A11, F0, F1 = mk("A10", T1, d0), mk("A11", T1, d0), mk("F0", T0, d0), mk("F1", T0, d0)
        V = _View
        self.xin = V(self.xpad, T0, W.cin0, shared=True)
        self.A1, self.A9, self.A11, self.F0, self.F1 = V(A1, T0, d0), V(A9, T2, d1), V(A11, T1, d0), V(F0, T0, d0), V(F1, T0, d0)
        self.D1, self.D2 = V(D1, T1, d0), V(D2, T2, d1)
        self.h1 = V(cat1, T1, d1, c0=d1, ctot=2 * d1)            # skip of level 1 = second half of up_modules.1's input
        self.h2 = V(cat0, T2, d2, c0=d2, ctot=2 * d2)            # skip of level 2 = second half of up_modules.0's input
        self.cat0, self.cat1 = V(cat0, T2, 2 * d2), V(cat1, T1, 2 * d1)
        self.up0_out = V(cat1, T1, d1, c0=0, ctot=2 * d1)
        mid_out = V(cat0, T2, d2, c0=0, ctot=2 * d2)
        chain = [(self.xin, V(A0, T0, d0)), (V(A0, T0, d0), self.A1), (self.D1, V(A2, T1, d1)), (V(A2, T1, d1), self.h1),
                 (self.D2, V(A4, T2, d2)), (V(A4, T2, d2), self.h2), (self.h2, V(A6, T2, d2)), (V(A6, T2, d2), mid_out),
                 (self.cat0, V(A8, T2, d1)), (V(A8, T2, d1), self.A9), (self.cat1, V(A10, T1, d0)), (V(A10, T1, d0), self.A11)]
        self.blocks: List[_Block] = []
        for i, ((pfx, _, co), (x, out)) in enumerate(zip(block_names(), chain)):
            y1 = V(mk(f"U{i}", out.T, co), out.T, co)
            r = V(mk(f"R{i}", out.T, co), out.T, co) if pfx + "r.w" in W.t else None
            raws = [plan.buf(f"{tag}.raw{i}.{j}", (G, B, out.T, co), torch.float32, zero=False) if save_raw else None for j in (0, 1)]
            self.blocks.append(_Block(pfx, x, y1, out, r, *raws))
        self.rawF = plan.buf(f"{tag}.rawF", (G, B, T0, d0), torch.float32, zero=False) if save_raw else None


def build_unet_train_forward(plan: Plan, W: UnetWeights, tb: UnetTrainBuffers, film: torch.Tensor, tag: str = "fwd") -> None:
    """tb.xpad (+ the per-sample FiLM table film fp32 [G][B][11264]) -> tb.out [G][B][T][A]; 42 GEMM launches."""
    T_, B = W.t, tb.B
    T0, T1, T2 = tb.Ts
    d0, d1, d2 = DOWN_DIMS
    ds_taps = [(1, -1), (0, 0), (1, 0)]

    def crb(k: int) -> None:
        b = tb.blocks[k]
        pfx, co, t = b.pfx, b.out.C, b.out.T
        _conv(plan, W, B, b.x, b.y1, T_[pfx + "c0.w"], T_[pfx + "c0.b"], taps=K5, cin_pad=b.x.C, n=co, t_out=t,
              gn=(T_[pfx + "g0.w"], T_[pfx + "g0.b"]), film=(film, None, 0, W.film_off[pfx]), raw=b.raw0,
              tag=f"{tag}.{pfx}conv0+gn+mish+film")
        res = b.x
        if b.r is not None:
            _conv(plan, W, B, b.x, b.r, T_[pfx + "r.w"], T_[pfx + "r.b"], taps=K1, cin_pad=b.x.C, n=co, t_out=t,
                  tag=f"{tag}.{pfx}residual_conv")
            res = b.r
        _conv(plan, W, B, b.y1, b.out, T_[pfx + "c1.w"], T_[pfx + "c1.b"], taps=K5, cin_pad=co, n=co, t_out=t,
              gn=(T_[pfx + "g1.w"], T_[pfx + "g1.b"]), res=res, raw=b.raw1, tag=f"{tag}.{pfx}conv1+gn+mish+res")

    def up(U: int, src: _View, dst: _View, t_in: int, c: int) -> None:
        for ph, taps in ((0, [(0, 0), (0, -1)]), (1, [(0, 1), (0, 0)])):
            _conv(plan, W, B, src, dst, T_[f"us{U}.w{ph}"], T_[f"us{U}.b"], taps=taps, cin_pad=c, n=c, t_out=t_in,
                  out_rows=(2 * t_in, 2, ph, 2 * t_in), tag=f"{tag}.up{U}.upsample.phase{ph}")

    crb(0); crb(1)
    _conv(plan, W, B, tb.A1, tb.D1, T_["ds0.w"], T_["ds0.b"], taps=ds_taps, cin_pad=d0, n=d0, t_out=T1, phases=2, tag=f"{tag}.down0.downsample")
    crb(2); crb(3)
    _conv(plan, W, B, tb.h1, tb.D2, T_["ds1.w"], T_["ds1.b"], taps=ds_taps, cin_pad=d1, n=d1, t_out=T2, phases=2, tag=f"{tag}.down1.downsample")
    for k in (4, 5, 6, 7, 8, 9):
        crb(k)
    up(0, tb.A9, tb.up0_out, T2, d1)
    crb(10); crb(11)
    up(1, tb.A11, tb.F0, T1, d0)
    _conv(plan, W, B, tb.F0, tb.F1, T_["final0.w"], T_["final0.b"], taps=K5, cin_pad=d0, n=d0, t_out=T0,
          gn=(T_["final0.gw"], T_["final0.gb"]), raw=tb.rawF, tag=f"{tag}.final_conv.0+gn+mish")
    _conv(plan, W, B, tb.F1, None, T_["final1.w"], T_["final1.b"], taps=K1, cin_pad=d0, n=W.A, t_out=T0, bn=32,
          out_f32=tb.out, tag=f"{tag}.final_conv.1")


Grad = Tuple[torch.Tensor, int]     # (buffer, taps): taps > 0 -> packed conv weight gradient (unet_bwd.unpack_wgrad), 0 -> as is


def build_unet_backward(plan: Plan, W: UnetWeights, sds: Sequence[SD], tb: UnetTrainBuffers, dvs: torch.Tensor,
                        film: torch.Tensor, dfilm: torch.Tensor, tag: str = "bwd") -> Dict[str, Grad]:
    """Gradients of every convolution / GroupNorm parameter of the G nets from dvs = d loss / d tb.out (fp32 [G][B][T][A]);
    fills dfilm (fp32 [G][B][11264]: d scale | d shift of every block).  Returns {reference state-dict key: (buffer, taps)}."""
    G, B = W.G, tb.B
    T0, T1, T2 = tb.Ts
    d0, d1, d2 = DOWN_DIMS
    dev = plan.device
    grads: Dict[str, Grad] = {}
    g = lambda k: ub.Ws(sds, k, dev)
    f32 = lambda name, t, c: plan.buf(f"{tag}.{name}", (G, B, t, c), torch.float32)
    V = _View

    # final_conv.1: Conv1d(d0 -> A, k1)
    dvb = ub.cast_bf16(plan, G, B, dvs, T0, f"{tag}.dvs.bf16", c_pad=W.mode.ke)
    grads["final_conv.1.weight"] = (ub.conv_wgrad(plan, W, B, dvb, tb.F1, tap_off=[0], t_out=T0, tag=f"{tag}.final_conv.1.wgrad"), 1)
    grads["final_conv.1.bias"] = (ub.colsum(plan, G, B, dvs, T0, f"{tag}.final_conv.1.dbias"), 0)
    dF1 = f32("dF1", T0, d0)
    ub.conv_dgrad(plan, W, B, dvb, V(dF1, T0, d0), g("final_conv.1.weight"), pad=0, tag=f"{tag}.final_conv.1.dgrad")
    # final_conv.0: Conv1dBlock
    dF0 = f32("dF0", T0, d0)
    b = ub.conv_block_backward(plan, W, B, tb.F0, g("final_conv.0.block.0.weight"), g("final_conv.0.block.0.bias"),
                               g("final_conv.0.block.1.weight"), g("final_conv.0.block.1.bias"), dF1, V(dF0, T0, d0),
                               raw=tb.rawF, tag=f"{tag}.final_conv.0")
    grads.update({"final_conv.0.block.0.weight": (b["dw"], 5), "final_conv.0.block.0.bias": (b["dbias"], 0),
                  "final_conv.0.block.1.weight": (b["dgamma"], 0), "final_conv.0.block.1.bias": (b["dbeta"], 0)})

    def up_bwd(U: int, x: _View, dy, t_in: int, c: int, dx: torch.Tensor) -> None:
        """ConvTranspose1d(k4, s2, p1): x [t_in] -> y [2 t_in]; dy fp32 (tensor / window over 2 t_in positions)"""
        key = f"up_modules.{U}.2.conv."
        dyb = ub.cast_bf16(plan, G, B, dy, 2 * t_in, f"{tag}.{key}dy.bf16")
        grads[key + "weight"] = (ub.conv_wgrad(plan, W, B, x, dyb, tap_off=[-1, 0, 1, 2], stride=2, t_out=t_in, tag=f"{tag}.{key}wgrad"), 4)
        grads[key + "bias"] = (ub.colsum(plan, G, B, dy, 2 * t_in, f"{tag}.{key}dbias"), 0)
        ub.upsample_dgrad(plan, W, B, dyb, V(dx, t_in, c), g(key + "weight"), tag=f"{tag}.{key}dgrad")

    def down_bwd(L: int, x: _View, dy: torch.Tensor, t_out: int, c: int, dx: torch.Tensor, res: Optional[_View]) -> None:
        """Conv1d(k3, s2, p1): x [2 t_out] -> y [t_out]"""
        key = f"down_modules.{L}.2.conv."
        dyb = ub.cast_bf16(plan, G, B, dy, t_out, f"{tag}.{key}dy.bf16")
        grads[key + "weight"] = (ub.conv_wgrad(plan, W, B, dyb, x, tap_off=[-1, 0, 1], stride=2, t_out=t_out, tag=f"{tag}.{key}wgrad"), 3)
        grads[key + "bias"] = (ub.colsum(plan, G, B, dy, t_out, f"{tag}.{key}dbias"), 0)
        ub.downsample_dgrad(plan, W, B, dyb, V(dx, 2 * t_out, c), g(key + "weight"), tag=f"{tag}.{key}dgrad", res=res)

    def crb_bwd(k: int, dout, dx: Optional[_View]) -> None:
        blk = tb.blocks[k]
        pfx = blk.pfx
        out = ub.res_block_backward(plan, W, B, sds, pfx, blk.x, blk.y1, dout, dx, (film, dfilm, W.film_off[pfx]), raws=(blk.raw0, blk.raw1),
                                    tag=f"{tag}.{pfx}")
        for key, v in out.items():
            grads[pfx + key] = v if isinstance(v, tuple) else (v, 0)

    dA11 = f32("dA11", T1, d0)
    up_bwd(1, tb.A11, dF0, T1, d0, dA11)
    dA10, dcat1 = f32("dA10", T1, d0), f32("dcat1", T1, 2 * d1)
    crb_bwd(11, dA11, V(dA10, T1, d0))
    crb_bwd(10, dA10, V(dcat1, T1, 2 * d1))
    dA9 = f32("dA9", T2, d1)
    up_bwd(0, tb.A9, V(dcat1, T1, d1, c0=0), T2, d1, dA9)
    dA8, dcat0 = f32("dA8", T2, d1), f32("dcat0", T2, 2 * d2)
    crb_bwd(9, dA9, V(dA8, T2, d1))
    crb_bwd(8, dA8, V(dcat0, T2, 2 * d2))
    dA6, dh2a, dh2 = f32("dA6", T2, d2), f32("dh2a", T2, d2), f32("dh2", T2, d2)
    crb_bwd(7, V(dcat0, T2, d2, c0=0), V(dA6, T2, d2))
    crb_bwd(6, dA6, V(dh2a, T2, d2))
    e = nv.EwiseDesc()                                           # level-2 skip: d h2 = (through mid / up path) + (cat half)
    e.a, e.a_ld, e.b, e.b_ld, e.out, e.out_ld = ptr(dh2a), d2, ptr(dcat0, d2), 2 * d2, ptr(dh2), d2
    e.rows, e.cols, e.op = G * B * T2, d2, nv.EW_ADD
    plan.add(e, f"{tag}.skip2.add")
    dA4, dD2 = f32("dA4", T2, d2), f32("dD2", T2, d1)
    crb_bwd(5, dh2, V(dA4, T2, d2))
    crb_bwd(4, dA4, V(dD2, T2, d1))
    dh1 = f32("dh1", T1, d1)
    down_bwd(1, tb.h1, dD2, T2, d1, dh1, res=V(dcat1, T1, d1, c0=d1))          # level-1 skip summed in the dgrad epilogue
    dA2, dD1 = f32("dA2", T1, d1), f32("dD1", T1, d0)
    crb_bwd(3, dh1, V(dA2, T1, d1))
    crb_bwd(2, dA2, V(dD1, T1, d0))
    dA1 = f32("dA1", T0, d0)
    down_bwd(0, tb.A1, dD1, T1, d0, dA1, res=None)                              # the level-0 skip is never consumed
    dA0 = f32("dA0", T0, d0)
    crb_bwd(1, dA1, V(dA0, T0, d0))
    crb_bwd(0, dA0, None)                                                       # d sample is not needed
    return grads


def grad_tensor(v: Grad, ref_shape) -> torch.Tensor:
    """-> [G, *ref_shape] view/copy of a gradient buffer in the reference parameter's own layout."""
    buf, taps = v
    if taps:
        return ub.unpack_wgrad(buf, ref_shape[1], taps)[:, : ref_shape[0]]
    return buf[:, : ref_shape[0]] if buf.dim() == 2 else buf


# ------------------------------------------------------------------------------------------------
# time embedding + FiLM (diffusion_step_encoder, cond_encoder): training forward keeps the pre-activations
# ------------------------------------------------------------------------------------------------
def _pack_rows(plan: Plan, src: torch.Tensor, src_ld: int, rows: int, cols: int, out: torch.Tensor, out_ld: int, dst_c0: int,
               act: int, tag: str, out_off: int = 0) -> None:
    d = nv.PackDesc()
    d.src, d.src_ld, d.rows, d.cols, d.act = ptr(src), src_ld, rows, cols, act
    d.out, d.out_dtype, d.out_ld, d.dst_c0 = ptr(out, out_off), nv.VT_BF16 if out.dtype == torch.bfloat16 else nv.VT_F32, out_ld, dst_c0
    d.out_plane, d.zero_to = 0, 0
    plan.add(d, tag)


def build_time_film_train(plan: Plan, W: UnetWeights, t_rows: torch.Tensor, B: int, cond: torch.Tensor, film: torch.Tensor,
                          tag: str = "film") -> Dict[str, torch.Tensor]:
    """film[G][B][11264] = cond_encoder(Mish(cat(diffusion_step_encoder(t), cond))) for G nets like unet.build_time_film, but
    with the Mish inputs t1 = Linear(256 -> 1024)(pe) and gf = cat(temb, cond) kept in fp32 for the backward
    (conditional_unet_1D.py:186-191, 209-213, 76-80)."""
    m, G, T_ = W.mode, W.G, W.t
    kd = DSED + W.cond_dim
    emb = plan.buf(f"{tag}.emb", (B, DSED), m.tdt)
    t1 = plan.buf(f"{tag}.t1", (G, B, 4 * DSED), torch.float32)
    hid = plan.buf(f"{tag}.hid", (G, B, 4 * DSED), m.tdt)
    gf = plan.buf(f"{tag}.gf", (G, B, kd), torch.float32)
    mgf = plan.buf(f"{tag}.mgf", (G, B, kd), m.tdt)
    plan.add(_tembed(t_rows, B, emb, m), f"{tag}.sinusoid")
    plan.add(linear_desc(a=emb, rows=B, k=DSED, a_ld=DSED, w=T_["time1.w"], n=4 * DSED, n_pad=4 * DSED, w_ld=DSED, out=t1,
                         ldc=4 * DSED, bias=T_["time1.b"], G=G, a_G=1, out_g=B * 4 * DSED), f"{tag}.time_mlp.0")
    _pack_rows(plan, t1, 4 * DSED, G * B, 4 * DSED, hid, 4 * DSED, 0, nv.ACT_MISH, f"{tag}.mish(t1)")
    plan.add(linear_desc(a=hid, rows=B, k=4 * DSED, a_ld=4 * DSED, w=T_["time2.w"], n=DSED, n_pad=DSED, w_ld=4 * DSED, out=gf,
                         ldc=kd, bias=T_["time2.b"], G=G, a_G=G, a_sG=B * 4 * DSED, out_g=B * kd), f"{tag}.time_mlp.1")
    for g in range(G):
        _pack_rows(plan, cond, cond.shape[-1], B, W.cond_dim, gf, kd, DSED, nv.ACT_NONE, f"{tag}.cat(cond).g{g}", out_off=g * B * kd)
    _pack_rows(plan, gf, kd, G * B, kd, mgf, kd, 0, nv.ACT_MISH, f"{tag}.mish(gf)")
    plan.add(linear_desc(a=mgf, rows=B, k=kd, a_ld=kd, w=T_["film.w_full"], n=FILM_ROWS, n_pad=FILM_ROWS, w_ld=kd, out=film,
                         ldc=FILM_ROWS, bias=T_["film.b"], G=G, a_G=G, a_sG=B * kd, out_g=B * FILM_ROWS), f"{tag}.film_gemm")
    return dict(emb=emb, t1=t1, hid=hid, gf=gf, mgf=mgf)


def _ewise(plan: Plan, a, a_ld: int, b, b_ld: int, out: torch.Tensor, rows: int, cols: int, op: int, tag: str) -> None:
    e = nv.EwiseDesc()
    e.a, e.a_ld, e.b, e.b_ld, e.out, e.out_ld, e.rows, e.cols, e.op = a, a_ld, b, b_ld, ptr(out), out.shape[-1], rows, cols, op
    plan.add(e, tag)


def build_film_time_backward(plan: Plan, W: UnetWeights, sds: Sequence[SD], tf: Dict[str, torch.Tensor], B: int,
                             film: torch.Tensor, dfilm: torch.Tensor, grads: Dict[str, Grad], tag: str = "bwd.film") -> Dict[str, torch.Tensor]:
    """From the d FiLM table of all 12 blocks: gradients of every cond_encoder Linear (one wgrad GEMM for the stacked
    [11264 x 512] matrix), d Mish(gf) -> d gf, the diffusion_step_encoder MLP, and d global_cond per net
    (tail of `unet_backward` in oracle/vt_oracle_bwd.py).  Adds the parameter gradients to `grads`; returns {"dcond": [G][B][cond]}."""
    G, dev = W.G, plan.device
    kd = DSED + W.cond_dim
    V = _View
    # cond_encoder.1 of all blocks: rows [off, off + 2 C) of the stacked matrix
    dwf = ub.conv_wgrad(plan, W, B, V(dfilm, 1, FILM_ROWS), V(tf["mgf"], 1, kd), tap_off=[0], t_out=1, tag=f"{tag}.cond_encoder.wgrad")
    dbf = ub.colsum(plan, G, B, dfilm, 1, f"{tag}.cond_encoder.dbias")
    for pfx, _, co in block_names():
        off = W.film_off[pfx]
        grads[pfx + "cond_encoder.1.weight"] = (dwf[:, off: off + 2 * co], 0)
        grads[pfx + "cond_encoder.1.bias"] = (dbf[:, off: off + 2 * co], 0)
    # d Mish(gf) = d film @ W_full
    pack_wt = lambda sds_: torch.stack([torch.cat([sd[pfx + "cond_encoder.1.weight"].detach().to(dev).float()
                                                   for pfx, _, _ in block_names()]).t() for sd in sds_]).to(_operand_dtype()).contiguous()
    wt = plan.reg(pack_wt(sds))                                                                       # [G][512][11264]
    dfb = ub.cast_bf16(plan, G, B, dfilm, 1, f"{tag}.dfilm.bf16")
    dmgf = plan.buf(f"{tag}.dmgf", (G, B, kd), torch.float32)
    plan.add(linear_desc(a=dfb.t, rows=B, k=FILM_ROWS, a_ld=FILM_ROWS, w=wt, n=kd, n_pad=kd, w_ld=FILM_ROWS, out=dmgf, ldc=kd,
                         G=G, a_G=G, a_sG=B * FILM_ROWS, out_g=B * kd), f"{tag}.cond_encoder.dgrad")
    dgf = plan.buf(f"{tag}.dgf", (G, B, kd), torch.float32)
    _ewise(plan, ptr(dmgf), kd, ptr(tf["gf"]), kd, dgf, G * B, kd, nv.EW_MISH_BWD, f"{tag}.mish'(gf)")
    dtemb = V(dgf, 1, DSED, c0=0)
    # diffusion_step_encoder.3: Linear(1024 -> 256)
    grads["diffusion_step_encoder.3.weight"] = (ub.conv_wgrad(plan, W, B, dtemb, V(tf["hid"], 1, 4 * DSED), tap_off=[0], t_out=1,
                                                              tag=f"{tag}.time_mlp.1.wgrad"), 0)
    grads["diffusion_step_encoder.3.bias"] = (ub.colsum(plan, G, B, dtemb, 1, f"{tag}.time_mlp.1.dbias"), 0)
    pack_w3t = lambda sds_: torch.stack([sd["diffusion_step_encoder.3.weight"].detach().to(dev).float().t()
                                         for sd in sds_]).to(_operand_dtype()).contiguous()
    w3t = plan.reg(pack_w3t(sds))
    if not hasattr(W, "repack"):
        W.repack = []
    W.repack += [(wt, pack_wt), (w3t, pack_w3t)]
    dtb = ub.cast_bf16(plan, G, B, dtemb, 1, f"{tag}.dtemb.bf16")
    dhid = plan.buf(f"{tag}.dhid", (G, B, 4 * DSED), torch.float32)
    plan.add(linear_desc(a=dtb.t, rows=B, k=DSED, a_ld=DSED, w=w3t, n=4 * DSED, n_pad=4 * DSED, w_ld=DSED, out=dhid, ldc=4 * DSED,
                         G=G, a_G=G, a_sG=B * DSED, out_g=B * 4 * DSED), f"{tag}.time_mlp.1.dgrad")
    dt1 = plan.buf(f"{tag}.dt1", (G, B, 4 * DSED), torch.float32)
    _ewise(plan, ptr(dhid), 4 * DSED, ptr(tf["t1"]), 4 * DSED, dt1, G * B, 4 * DSED, nv.EW_MISH_BWD, f"{tag}.mish'(t1)")
    # diffusion_step_encoder.1: Linear(256 -> 1024) on the sinusoidal embedding (shared by the nets)
    grads["diffusion_step_encoder.1.weight"] = (ub.conv_wgrad(plan, W, B, V(dt1, 1, 4 * DSED), V(tf["emb"], 1, DSED, shared=True),
                                                              tap_off=[0], t_out=1, tag=f"{tag}.time_mlp.0.wgrad"), 0)
    grads["diffusion_step_encoder.1.bias"] = (ub.colsum(plan, G, B, dt1, 1, f"{tag}.time_mlp.0.dbias"), 0)
    return dict(dcond=dgf[:, :, DSED:], dgf=dgf)


class LossBackwardProgram:
    """StochasticInterpolants.get_loss(...) and its backward (bridge_model.py:220-257, 183-218; bridge_train.py:315-334) as one
    program for [b_net, v_net, s_net]:  q_sample -> time embedding / FiLM -> training forward of the three U-Nets -> the three
    losses -> their derivative -> the explicit backward (build_unet_backward, build_film_time_backward).

    Inputs: x0 (vla_act), x1 (expert_act) [B,T,A], cond [B,cond] (obs_cond), step [B] ~ U(0,1), z_unit [B,T,A] ~ N(0,1).
    Outputs: .out fp32 [4] = (loss, v_loss, s_loss, b_loss); .grads {'b_net.'/'v_net.'/'s_net.' + reference key: tensor in the
    reference parameter's layout}; .d_cond [B,cond] = d loss / d obs_cond (summed over the nets: the gradient that trains the
    state encoder)."""
    NETS = ("b_net.", "v_net.", "s_net.")

    def __init__(self, sds_bvs: Sequence[SD], action_dim: int, B: int, T: int, beta_max: float, device):
        assert len(sds_bvs) == 3, "expects [b_net, v_net, s_net]"
        self.sds = [dict(sd) for sd in sds_bvs]
        self.W = W = UnetWeights(sds_bvs, action_dim, device, precise=False)
        self.plan = p = Plan(device)
        W.register(p)
        m, A, f32 = W.mode, action_dim, torch.float32
        self.B, self.T, self.A = B, T, A
        self.x0, self.x1 = p.buf("in.x0", (B, T, A), f32), p.buf("in.x1", (B, T, A), f32)
        self.cond = p.buf("in.cond", (B, W.cond_dim), f32)
        self.step, self.z = p.buf("in.step", (B,), f32), p.buf("in.z_unit", (B, T, A), f32)
        self.xt, self.tclip = p.buf("xt", (B, T, A), f32), p.buf("tclip", (B,), f32)
        self.film, self.dfilm = p.buf("film", (3, B, FILM_ROWS), f32), p.buf("dfilm", (3, B, FILM_ROWS), f32)
        self.tb = tb = UnetTrainBuffers(p, W, B, T)
        self.per_sample, self.out = p.buf("loss.per_sample", (3, B), f32), p.buf("loss.out", (4,), f32)
        self.dvs = p.buf("loss.dvs", (3, B, T, A), f32)
        d = nv.QsampleDesc()
        d.x0, d.x1, d.step, d.z_unit, d.d = ptr(self.x0), ptr(self.x1), ptr(self.step), ptr(self.z), beta_max
        d.B, d.n, d.A, d.xt, d.tclip = B, T * A, A, ptr(self.xt), ptr(self.tclip)
        d.xpad, d.xpad_dtype, d.xpad_ld, d.xpad_plane = ptr(tb.xpad), m.dt, W.cin0, 0
        p.add(d, "q_sample")
        self.tf = build_time_film_train(p, W, self.tclip, B, self.cond, self.film)
        build_unet_train_forward(p, W, tb, self.film)
        d = nv.SilossDesc()
        d.bvs, d.x0, d.x1, d.z_unit, d.tclip, d.d = ptr(tb.out), ptr(self.x0), ptr(self.x1), ptr(self.z), ptr(self.tclip), beta_max
        d.B, d.n, d.per_sample, d.out = B, T * A, ptr(self.per_sample), ptr(self.out)
        p.add(d, "si_losses")
        self.n_forward_ops = len(p)
        d = nv.SilossBwdDesc()
        d.bvs, d.x0, d.x1, d.z_unit, d.tclip, d.d = ptr(tb.out), ptr(self.x0), ptr(self.x1), ptr(self.z), ptr(self.tclip), beta_max
        d.B, d.n, d.dvs = B, T * A, ptr(self.dvs)
        p.add(d, "si_losses.bwd")
        # every parameter gradient lives in ONE contiguous fp32 arena, in the order the backward produces them: the
        # data-parallel all-reduce runs in place on slices of it and the optimizer reads it in place (`grad_sources`)
        n_net = sum(int(v.numel()) for v in self.sds[0].values())
        p.set_arena("grads", 3 * (n_net + 256 * 1024) + 64 * 2048)
        self._g = build_unet_backward(p, W, self.sds, tb, self.dvs, self.film, self.dfilm)
        self._x = build_film_time_backward(p, W, self.sds, self.tf, B, self.film, self.dfilm, self._g)

    def refresh(self, sds_bvs: Sequence[SD]) -> None:
        """New parameter values (after an optimizer step): re-pack the forward operands and every transposed / sliced copy the
        backward GEMMs read, into the same device tensors."""
        self.sds = [dict(sd) for sd in sds_bvs]
        self.W.refresh(sds_bvs)
        ub.repack_all(self.W, sds_bvs)

    # ---- re-pack by gather maps ----
    def setup_gather(self, params_bvs: Sequence[Dict[str, torch.nn.Parameter]], sds_bvs: Sequence[SD]) -> bool:
        """Replace the ~2000 tiny tensor ops of `refresh` (6.5 ms of device time per training step even as a CUDA graph) by one
        gather per operand tensor.  Every packed operand (forward weights, the transposed / tap-sliced dgrad copies, stacked FiLM
        matrices, bias / gamma vectors) is a fixed rearrangement of parameter elements plus zero padding, so it is described by an
        index map -- obtained by running the very same packing functions on index-valued tensors -- into ONE contiguous fp32 arena
        that the parameters are re-pointed into (`p.data` becomes a view; values, Parameter objects and state_dict keys are
        unchanged).  A bf16 mirror of the arena (one cast per step) feeds the bf16 operands.  The maps are verified against the
        tensor-op re-pack before they are used; returns False (and changes nothing) if the verification fails."""
        if self.plan.device.type != "cuda" or self.W.mode.precise or getattr(self, "_gather", None) is not None:
            return getattr(self, "_gather", None) is not None
        dev = self.plan.device
        self.refresh(sds_bvs)                               # operands = the current parameter values: what the maps are verified against
        plist = [p for ps in params_bvs for p in ps.values()]
        if any(p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev for p in plist):
            return False
        offs, total = [], 0
        for p in plist:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        zero_slot = total
        arena = torch.zeros(total + 4, dtype=torch.float32, device=dev)
        # index-valued stand-ins: parameter id + 1 and element offset + 1 (both exact in fp32), 0 = padding
        fake_id, fake_off, i = [], [], 0
        for g, ps in enumerate(params_bvs):
            fid, fof = {}, {}
            for k, p in ps.items():
                fid[k] = torch.full(p.shape, float(i + 1), dtype=torch.float32, device=dev)
                fof[k] = (torch.arange(p.numel(), dtype=torch.float32, device=dev) + 1.0).view(p.shape)
                i += 1
            fake_id.append(fid)
            fake_off.append(fof)
        off_t = torch.tensor(offs, dtype=torch.int64, device=dev)
        numel_t = torch.tensor([p.numel() for p in plist], dtype=torch.int64, device=dev)
        dests = [(k, t) for k, t in self.W.t.items()] + [(f"tracked{j}", d) for j, (d, _) in enumerate(getattr(self.W, "repack", []))]

        def traced(fake):
            Mode.TRACE = True
            try:
                out = list(self.W._pack(fake).values()) + [fn(fake) for _, fn in getattr(self.W, "repack", [])]
            finally:
                Mode.TRACE = False
            return out
        ids, ofs = traced(fake_id), traced(fake_off)
        maps = []
        for (name, dest), mi, mo in zip(dests, ids, ofs):
            if tuple(mi.shape) != tuple(dest.shape):
                return False
            mi_f, mo_f = mi.reshape(-1), mo.reshape(-1)
            mi = mi_f.round().long()
            mo = mo_f.round().long()
            if int(mi.max()) == 0:
                continue                                    # independent of the parameters (zero vectors)
            # a pure rearrangement: integer ids in range, offsets inside the parameter, padding exactly where the id is 0
            sel = mi > 0
            lim = numel_t[(mi - 1).clamp(0, len(plist) - 1)]
            if (bool(((mi_f - mi).abs() > 1e-3).any()) or bool(((mo_f - mo).abs() > 1e-3).any()) or int(mi.min()) < 0 or int(mi.max()) > len(plist)
                    or bool((sel & ((mo < 1) | (mo > lim))).any()) or bool(((~sel) & (mo != 0)).any())):
                return False
            idx = torch.where(mi > 0, off_t[(mi - 1).clamp_min(0)] + mo - 1, torch.full_like(mi, zero_slot))
            maps.append((dest, idx.to(torch.int32)))
        # move the parameters into the arena, verify the maps against the tensor-op re-pack, then switch over
        with torch.no_grad():
            for p, o in zip(plist, offs):
                arena[o: o + p.numel()].copy_(p.detach().reshape(-1))
        arena_bf = arena.to(torch.bfloat16)
        for dest, idx in maps:
            src = arena_bf if dest.dtype == torch.bfloat16 else arena
            if dest.dtype not in (torch.bfloat16, torch.float32):
                return False
            got = torch.index_select(src, 0, idx).view(dest.shape)
            if not torch.equal(got, dest):
                return False
        with torch.no_grad():
            for p, o in zip(plist, offs):
                p.data = arena[o: o + p.numel()].view(p.shape)
        # one native launch for all operands: records {dst, map, n, bf16} + (record, offset) chunks of 8192 elements
        import struct
        recs = b"".join(struct.pack("<QQqii", d.data_ptr(), idx.data_ptr(), d.numel(), 1 if d.dtype == torch.bfloat16 else 0, 0) for d, idx in maps)
        chunks = [(i, off) for i, (d, _) in enumerate(maps) for off in range(0, d.numel(), 8192)]
        self._gather = (arena, maps, torch.frombuffer(bytearray(recs), dtype=torch.uint8).to(dev), torch.tensor(chunks, dtype=torch.int64, device=dev), len(chunks))
        self._gather_probe = [(p, arena.data_ptr() + 4 * o) for p, o in zip(plist, offs)]
        return True

    def gather_valid(self) -> bool:
        """False once a parameter no longer lives in the arena (`net.to(...)`, a replaced `p.data`): the caller falls back to the
        tensor-op re-pack and may set the gather up again."""
        return getattr(self, "_gather", None) is not None and all(p.data_ptr() == ptr_ for p, ptr_ in self._gather_probe)

    def refresh_gather(self) -> None:
        """dst[i] = arena[map[i]] for every packed operand, one kernel launch (csrc/vt_elem.cuh gather_repack_kernel)."""
        arena, _, recs, chunks, n_chunks = self._gather
        nv.check(nv.lib().vt_gather_repack(recs.data_ptr(), chunks.data_ptr(), n_chunks, arena.data_ptr(), nv.current_stream_ptr()))

    def refresh_graphed(self, sds_bvs: Sequence[SD]) -> None:
        """`refresh` replayed as one CUDA graph.  The re-pack is ~2000 tiny tensor ops (permute / pad / cast per parameter and
        operand copy): ~20 ms of host time per training step when issued one by one, with the GPU idle in between.  The
        parameters are updated in place by the optimizer, so the whole re-pack is a static sequence between fixed addresses: it
        is captured once (the intermediates live in the graph's private pool) and replayed per step; re-captured if a parameter
        tensor was replaced."""
        ptrs = tuple(v.data_ptr() for sd in sds_bvs for v in sd.values())
        if getattr(self, "_refresh_ptrs", None) != ptrs or self.plan.device.type != "cuda":
            self.refresh(sds_bvs)
            if self.plan.device.type == "cuda":
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.refresh(sds_bvs)
                self._refresh_graph, self._refresh_ptrs = g, ptrs
            return
        self._refresh_graph.replay()

    def set_inputs(self, x0, x1, cond, step, z_unit) -> None:
        self.x0.copy_(x0); self.x1.copy_(x1); self.cond.copy_(cond); self.step.copy_(step); self.z.copy_(z_unit)

    def run(self, graph: bool = False) -> torch.Tensor:
        """Launch the whole program on the current stream.  graph=True replays it as one CUDA graph (captured on first use;
        the program is static: fixed buffers, weights re-packed in place), which removes the ~300 launch latencies."""
        prog = self.plan.compile()
        self.runs = getattr(self, "runs", 0) + 1
        if graph:
            if not getattr(self, "_graph_ready", False):
                prog.graph_build()
                self._graph_ready = True
            prog.graph_launch()
        else:
            prog.run()
        return self.out

    @property
    def grads(self) -> Dict[str, torch.Tensor]:
        out = {}
        for key, v in self._g.items():
            full = grad_tensor(v, self.sds[0][key].shape)
            for n, pfx in enumerate(self.NETS):
                out[pfx + key] = full[n].reshape(self.sds[0][key].shape)
        return out

    @property
    def d_cond(self) -> torch.Tensor:
        return self._x["dcond"].sum(dim=0)

    def grad_arena(self):
        """(flat fp32 tensor holding every parameter gradient, [(element offset, numel, op index)] in production order)."""
        t, used, allocs = self.plan.arena("grads")
        return t[:used], allocs

    def grad_sources(self) -> Dict[str, Tuple[torch.Tensor, int, int, int]]:
        """{'b_net.' / 'v_net.' / 's_net.' + reference key: (contiguous gradient buffer of that net, taps, c, c_pad)} in the
        layout the kernels wrote (vt_opt_tensor: taps == 0 -> same order as the parameter, else [rows][taps][c_pad])."""
        out = {}
        for key, (buf, taps) in self._g.items():
            shape = tuple(self.sds[0][key].shape)
            for n, pfx in enumerate(self.NETS):
                b = buf[n]
                if taps:
                    out[pfx + key] = (b, taps, shape[1], b.shape[-1] // taps)
                elif b.dim() == 2 and b.shape[-1] != shape[-1]:      # linear weight whose K dimension is zero padded
                    out[pfx + key] = (b, 1, shape[1], b.shape[-1])
                else:
                    out[pfx + key] = (b, 0, 0, 0)
        return out
