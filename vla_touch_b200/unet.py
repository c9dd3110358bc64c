"""Plan builder for the conditional 1-D U-Net denoisers (reference:
bridge/networks/conditional_unet_1D.py:40-105,108-247 and conditional_unet_1D_si.py:4-50).

Every Conv1d / ConvTranspose1d / Linear of `G` structurally identical nets (v_net + s_net in the sampler,
b_net + v_net + s_net in the loss) is one grouped tcgen05 implicit-GEMM launch; GroupNorm(8) + Mish + FiLM +
the residual add run in that GEMM's epilogue.  Activations are channels-last [G][B][T][C]; torch.cat((x, skip))
is free because producers write straight into the two channel halves of the concat buffers.

FiLM:  cond_encoder = Sequential(Mish, Linear(512 -> 2C)) applied to gf = cat(temb, cond) is linear in Mish(gf),
so  W Mish(gf) + b = W[:, :256] Mish(temb) + (W[:, 256:] Mish(cond) + b).  The FiLM rows of all 12 residual blocks
are stacked into one [11264 x 512] matrix per net:  one GEMM per call gives the per-sample table `film_c`; in the
sampler the time half is batch-independent and pre-computed for all steps (`film_t`).
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import native as nv
from .plan import Plan, gemm_desc, linear_desc, ptr, round_up, tf32_round

SD = Dict[str, torch.Tensor]
DOWN_DIMS = (256, 512, 512)
KS = 5
N_GROUPS = 8
DSED = 256


class Mode:
    """bf16: kind::f16 MMA, bf16 activations.  f32: 3-pass split-tf32 MMA, fp32 activations stored as tf32 hi|lo planes."""

    def __init__(self, precise: bool):
        self.precise = precise
        self.dt = nv.VT_F32 if precise else nv.VT_BF16
        self.tdt = torch.float32 if precise else torch.bfloat16
        self.ke = 32 if precise else 64
        self.passes = 3 if precise else 1
        self.planes = 2 if precise else 1

    def ld(self, C: int) -> int:
        return C * self.planes

    def plane(self, C: int) -> int:
        return C if self.precise else 0

    TRACE = False     # unet_train.GatherRepack: run the packing functions on index-valued tensors (no cast to the operand dtype)

    def pack_w(self, w: torch.Tensor) -> torch.Tensor:
        """[..., K] fp32 (already zero padded) -> operand dtype; precise: [..., hi K | lo K]."""
        if Mode.TRACE:
            return w.float().contiguous()
        if not self.precise:
            return w.to(torch.bfloat16).contiguous()
        hi = tf32_round(w.float())
        return torch.cat((hi, w.float() - hi), dim=-1).contiguous()


def block_names() -> List[Tuple[str, int, int]]:
    """(prefix, C_in, C_out) of the 12 ConditionalResidualBlock1D in execution order (input dim filled by caller)."""
    d0, d1, d2 = DOWN_DIMS
    return [("down_modules.0.0.", -1, d0), ("down_modules.0.1.", d0, d0),
            ("down_modules.1.0.", d0, d1), ("down_modules.1.1.", d1, d1),
            ("down_modules.2.0.", d1, d2), ("down_modules.2.1.", d2, d2),
            ("mid_modules.0.", d2, d2), ("mid_modules.1.", d2, d2),
            ("up_modules.0.0.", 2 * d2, d1), ("up_modules.0.1.", d1, d1),
            ("up_modules.1.0.", 2 * d1, d0), ("up_modules.1.1.", d0, d0)]


FILM_ROWS = sum(2 * c for _, _, c in block_names())   # 11264


def _pack_conv(ws: Sequence[torch.Tensor], cin_pad: int, mode: Mode, n_pad: Optional[int] = None) -> torch.Tensor:
    """Conv1d weights [Cout, Cin, K] of G nets -> [G][n_pad][K * cin_pad] (tap-major, channels zero padded)."""
    out = []
    for w in ws:
        co, ci, k = w.shape
        npad = n_pad or round_up(co, 128)
        p = torch.zeros(npad, k, cin_pad, dtype=torch.float32, device=w.device)
        p[:co, :, :ci] = w.float().permute(0, 2, 1)
        out.append(p.reshape(npad, k * cin_pad))
    return mode.pack_w(torch.stack(out))


def _pack_vec(vs: Sequence[torch.Tensor], n_pad: int) -> torch.Tensor:
    out = torch.zeros(len(vs), n_pad, dtype=torch.float32, device=vs[0].device)
    for g, v in enumerate(vs):
        out[g, : v.numel()] = v.float().reshape(-1)
    return out.contiguous()


class UnetWeights:
    """Device-resident packed parameters of G nets with identical structure."""

    def __init__(self, sds: Sequence[SD], action_dim: int, device, precise: bool = False):
        self.G = len(sds)
        self.A = action_dim
        self.mode = m = Mode(precise)
        self.device = torch.device(device)
        self.cin0 = round_up(action_dim, m.ke)            # channel-padded network input
        self.film_off = {}
        self.t = self._pack(sds)

    def refresh(self, sds: Sequence[SD]) -> None:
        """Re-pack new parameter values into the existing device tensors (addresses unchanged)."""
        new = self._pack(sds)
        for k, t in self.t.items():
            t.copy_(new[k])

    def _pack(self, sds: Sequence[SD]) -> Dict[str, torch.Tensor]:
        m, action_dim = self.mode, self.A
        dev = lambda t: t.detach().to(self.device)
        g = lambda k: [dev(sd[k]) for sd in sds]
        T: Dict[str, torch.Tensor] = {}
        # time MLP: Linear(256,1024) Mish Linear(1024,256)
        T["time1.w"] = m.pack_w(torch.stack([w.float() for w in g("diffusion_step_encoder.1.weight")]))
        T["time1.b"] = _pack_vec(g("diffusion_step_encoder.1.bias"), 4 * DSED)
        T["time2.w"] = m.pack_w(torch.stack([w.float() for w in g("diffusion_step_encoder.3.weight")]))
        T["time2.b"] = _pack_vec(g("diffusion_step_encoder.3.bias"), DSED)
        # FiLM: rows of all blocks stacked in execution order
        names = block_names()
        off = 0
        wf = [[] for _ in sds]
        bf = [[] for _ in sds]
        for pfx, _, co in names:
            self.film_off[pfx] = off
            off += 2 * co
            for i, sd in enumerate(sds):
                wf[i].append(dev(sd[pfx + "cond_encoder.1.weight"]).float())
                bf[i].append(dev(sd[pfx + "cond_encoder.1.bias"]).float())
        wfull = torch.stack([torch.cat(x) for x in wf])                     # [G][11264][512]
        self.cond_dim = wfull.shape[-1] - DSED
        T["film.w_full"] = m.pack_w(wfull)
        T["film.w_time"] = m.pack_w(wfull[:, :, :DSED].contiguous())
        T["film.w_cond"] = m.pack_w(wfull[:, :, DSED:].contiguous())
        T["film.b"] = torch.stack([torch.cat(x) for x in bf]).contiguous()
        T["film.zero_b"] = torch.zeros_like(T["film.b"])
        # residual blocks
        for pfx, ci, co in names:
            ci = action_dim if ci < 0 else ci
            cpad = self.cin0 if ci == action_dim and pfx == "down_modules.0.0." else ci
            T[pfx + "c0.w"] = _pack_conv(g(pfx + "blocks.0.block.0.weight"), cpad, m)
            T[pfx + "c0.b"] = _pack_vec(g(pfx + "blocks.0.block.0.bias"), co)
            T[pfx + "g0.w"] = _pack_vec(g(pfx + "blocks.0.block.1.weight"), co)
            T[pfx + "g0.b"] = _pack_vec(g(pfx + "blocks.0.block.1.bias"), co)
            T[pfx + "c1.w"] = _pack_conv(g(pfx + "blocks.1.block.0.weight"), co, m)
            T[pfx + "c1.b"] = _pack_vec(g(pfx + "blocks.1.block.0.bias"), co)
            T[pfx + "g1.w"] = _pack_vec(g(pfx + "blocks.1.block.1.weight"), co)
            T[pfx + "g1.b"] = _pack_vec(g(pfx + "blocks.1.block.1.bias"), co)
            if pfx + "residual_conv.weight" in sds[0]:
                T[pfx + "r.w"] = _pack_conv(g(pfx + "residual_conv.weight"), cpad, m)
                T[pfx + "r.b"] = _pack_vec(g(pfx + "residual_conv.bias"), co)
        # down / up sampling
        for L in (0, 1):
            c = DOWN_DIMS[L]
            T[f"ds{L}.w"] = _pack_conv(g(f"down_modules.{L}.2.conv.weight"), c, m)       # k3 s2 p1, taps k=0,1,2
            T[f"ds{L}.b"] = _pack_vec(g(f"down_modules.{L}.2.conv.bias"), c)
        for U in (0, 1):
            ws = g(f"up_modules.{U}.2.conv.weight")                                       # ConvTranspose1d [Cin, Cout, 4]
            c = ws[0].shape[0]
            # out[2t]   = x[t] w[..,1] + x[t-1] w[..,3];   out[2t+1] = x[t+1] w[..,0] + x[t] w[..,2]
            even = [torch.stack((w[:, :, 1], w[:, :, 3]), dim=-1).permute(1, 0, 2) for w in ws]   # -> [Cout, Cin, 2]
            odd = [torch.stack((w[:, :, 0], w[:, :, 2]), dim=-1).permute(1, 0, 2) for w in ws]
            T[f"us{U}.w0"] = _pack_conv(even, c, m)
            T[f"us{U}.w1"] = _pack_conv(odd, c, m)
            T[f"us{U}.b"] = _pack_vec(g(f"up_modules.{U}.2.conv.bias"), c)
        c0 = DOWN_DIMS[0]
        T["final0.w"] = _pack_conv(g("final_conv.0.block.0.weight"), c0, m)
        T["final0.b"] = _pack_vec(g("final_conv.0.block.0.bias"), c0)
        T["final0.gw"] = _pack_vec(g("final_conv.0.block.1.weight"), c0)
        T["final0.gb"] = _pack_vec(g("final_conv.0.block.1.bias"), c0)
        T["final1.w"] = _pack_conv(g("final_conv.1.weight"), c0, m, n_pad=32)
        T["final1.b"] = _pack_vec(g("final_conv.1.bias"), 32)
        return T

    def register(self, plan: Plan) -> None:
        for t in self.t.values():
            plan.reg(t)


class UnetBuffers:
    """Activation buffers of one (G, B, T) U-Net evaluation (re-used by every step of the sampler)."""

    def __init__(self, plan: Plan, W: UnetWeights, B: int, T: int, tag: str = "u"):
        if T % 4 != 0 or T > 128 or T < 4:
            raise ValueError(f"horizon T={T} must be a multiple of 4 in [4, 128] (two stride-2 levels, one sample per tile)")
        m, G = W.mode, W.G
        self.B, self.T = B, T
        self.Ts = (T, T // 2, T // 4)
        d0, d1, d2 = DOWN_DIMS
        mk = lambda name, t, c: plan.buf(f"{tag}.{name}", (G, B, t, m.ld(c)), m.tdt)
        T0, T1, T2 = self.Ts
        self.xpad = plan.buf(f"{tag}.xpad", (B, T0, m.ld(W.cin0)), m.tdt)       # shared by all groups
        self.X0a, self.X0b, self.U0, self.R0 = (mk(n, T0, d0) for n in ("X0a", "X0b", "U0", "R0"))
        self.D1 = mk("D1", T1, d0)
        self.X1a, self.U1, self.R1 = (mk(n, T1, d1) for n in ("X1a", "U1", "R1"))
        self.cat1 = mk("cat1", T1, 2 * d1)
        self.D2 = mk("D2", T2, d1)
        self.X2a, self.X2b, self.U2, self.R2 = (mk(n, T2, d2) for n in ("X2a", "X2b", "U2", "R2"))
        self.cat0 = mk("cat0", T2, 2 * d2)
        self.V1, self.Y1a, self.Y1b, self.Q1 = (mk(n, T1, d0) for n in ("V1", "Y1a", "Y1b", "Q1"))
        self.out = plan.buf(f"{tag}.vs", (G, B, T0, W.A), torch.float32)        # net outputs [G][B][T][A]


class _View:
    """A channel window [c0, c0+C) of an activation buffer [G][B][T][ld]."""

    def __init__(self, t: torch.Tensor, T: int, C: int, c0: int = 0, shared: bool = False, ctot: Optional[int] = None):
        self.t, self.T, self.C, self.c0, self.shared = t, T, C, c0, shared
        self.ld = t.shape[-1]
        self.ctot = ctot if ctot is not None else C    # logical channels of the whole buffer (plane distance in precise mode)


def _conv(plan: Plan, W: UnetWeights, B: int, src: _View, dst: _View, w: torch.Tensor, bias: torch.Tensor, *, taps,
          cin_pad: int, n: int, t_out: int, phases: int = 1, gn=None, film=None, res: Optional[_View] = None,
          out_rows=None, bn: int = 128, out_f32: Optional[torch.Tensor] = None, raw: Optional[torch.Tensor] = None,
          tag: str = "") -> None:
    """One grouped implicit-GEMM convolution.  src/dst are channel windows of [G][B][T][ld] buffers.
    raw (GroupNorm convs of the training forward): fp32 [G][B][t_out][n] receiving conv + bias, the input of GroupNorm."""
    m, G = W.mode, W.G
    t_in_q = src.T // phases
    if bn == 128 and not m.precise and n % 256 == 0 and not (gn is not None and os.environ.get("VT_GN_BN") == "128"):
        bn = 256                                  # 256-wide tiles run as CTA pairs (M = 256 MMAs): half the operand bytes per MAC
    b_box = max(1, min(128 // t_out, B, 32 if bn == 128 else 16))
    a_sB = src.T * src.ld
    n_pad = w.shape[1]
    w_ld = w.shape[2]
    k_tot = len(taps) * cin_pad
    kw = dict(
        a=ptr(src.t, src.c0), in_dtype=m.dt, a_C=src.ld - src.c0, a_P=phases, a_T=t_in_q, a_B=B,
        a_G=1 if src.shared else G, a_ld=src.ld, a_sB=a_sB, a_sG=0 if src.shared else B * a_sB, kc=cin_pad, taps=taps,
        t_box=t_out, b_box=b_box, w=ptr(w), n_pad=n_pad, w_ld=w_ld, G=G, M=B * t_out, N=n, bn=bn, bias=ptr(bias),
        passes=m.passes, a_plane=m.plane(src.ctot), w_plane=k_tot if m.precise else 0)
    if out_f32 is not None:                                     # final 1x1 conv -> fp32 [G][B][T][A]
        kw.update(out=ptr(out_f32), out_dtype=nv.VT_F32, ldc=out_f32.shape[-1], out_g=B * t_out * out_f32.shape[-1],
                  row_div=1, out_q=1, out_r=0, out_off=0)
    else:
        oq, orr, ooff, t_dst = out_rows if out_rows else (t_out, 1, 0, t_out)
        f32_dst = dst.t.dtype == torch.float32 and not m.precise      # backward plans: fp32 gradient buffers in the bf16 mode
        kw.update(out=ptr(dst.t, dst.c0), out_dtype=nv.VT_F32 if f32_dst else m.dt, ldc=dst.ld, out_g=B * dst.T * dst.ld,
                  row_div=t_out, out_q=oq, out_r=orr, out_off=ooff, out_plane=0 if f32_dst else m.plane(dst.ctot))
        if gn is None and res is not None:                             # LINEAR epilogue: fp32 rows added last (same row mapping)
            assert res.t.dtype == torch.float32
            kw.update(res=ptr(res.t, res.c0), ldres=res.ld, res_g=B * res.T * res.ld, res_q=oq, res_r=orr, res_off=ooff)
    if gn is not None:
        gamma, beta = gn
        kw.update(epi=nv.EPI_GN, gn_gamma=ptr(gamma), gn_beta=ptr(beta), gn_group_ch=n // N_GROUPS, gn_eps=1e-5)
        if film is not None:
            film_c, film_t_ptr, film_tg, off = film
            kw.update(film_c=ptr(film_c), film_t=film_t_ptr, film_g=B * FILM_ROWS, film_tg=film_tg, film_ld=FILM_ROWS,
                      film_C=n, film_off=off)
        if res is not None:
            kw.update(res=ptr(res.t, res.c0), ldres=res.ld, res_g=B * res.T * res.ld, res_q=t_out, res_r=1, res_off=0,
                      res_plane=m.plane(res.ctot))
        if raw is not None:
            assert raw.dtype == torch.float32 and tuple(raw.shape) == (G, B, t_out, n)
            kw.update(raw_out=ptr(raw), raw_g=B * t_out * n, raw_ld=n)
    plan.add(gemm_desc(**kw), tag)


def build_unet_eval(plan: Plan, W: UnetWeights, bufs: UnetBuffers, film_c: torch.Tensor, film_t_ptr: Optional[int],
                    film_tg: int = 0, tag: str = "unet") -> None:
    """Append the 36 GEMMs of one evaluation of the G nets: bufs.xpad (+ FiLM tables) -> bufs.out.

    film_c: fp32 [G][B][11264]; film_t_ptr: address of group 0's row of an fp32 table added to it (groups are
    `film_tg` elements apart), or None."""
    T_, b, m, B = W.t, bufs, W.mode, bufs.B
    T0, T1, T2 = bufs.Ts
    d0, d1, d2 = DOWN_DIMS
    k5 = [(0, dt) for dt in (-2, -1, 0, 1, 2)]
    k1 = [(0, 0)]

    def crb(pfx, src: _View, cin_pad, cout, t, U: torch.Tensor, R: Optional[torch.Tensor], dst: _View):
        film = (film_c, film_t_ptr, film_tg, W.film_off[pfx])
        u = _View(U, t, cout)
        _conv(plan, W, B, src, u, T_[pfx + "c0.w"], T_[pfx + "c0.b"], taps=k5, cin_pad=cin_pad, n=cout, t_out=t,
              gn=(T_[pfx + "g0.w"], T_[pfx + "g0.b"]), film=film, tag=f"{tag}.{pfx}conv0+gn+mish+film")
        if pfx + "r.w" in T_:
            r = _View(R, t, cout)
            _conv(plan, W, B, src, r, T_[pfx + "r.w"], T_[pfx + "r.b"], taps=k1, cin_pad=cin_pad, n=cout, t_out=t,
                  tag=f"{tag}.{pfx}residual_conv")
            res = r
        else:
            res = src
        _conv(plan, W, B, u, dst, T_[pfx + "c1.w"], T_[pfx + "c1.b"], taps=k5, cin_pad=cout, n=cout, t_out=t,
              gn=(T_[pfx + "g1.w"], T_[pfx + "g1.b"]), res=res, tag=f"{tag}.{pfx}conv1+gn+mish+res")

    xin = _View(b.xpad, T0, W.cin0, shared=True)
    X0a, X0b = _View(b.X0a, T0, d0), _View(b.X0b, T0, d0)
    crb("down_modules.0.0.", xin, W.cin0, d0, T0, b.U0, b.R0, X0a)
    crb("down_modules.0.1.", X0a, d0, d0, T0, b.U0, None, X0b)
    ds_taps = [(1, -1), (0, 0), (1, 0)]                          # x[2t-1], x[2t], x[2t+1]  (k3, stride 2, pad 1)
    D1 = _View(b.D1, T1, d0)
    _conv(plan, W, B, X0b, D1, T_["ds0.w"], T_["ds0.b"], taps=ds_taps, cin_pad=d0, n=d0, t_out=T1, phases=2,
          tag=f"{tag}.down0.downsample")
    X1a = _View(b.X1a, T1, d1)
    h1 = _View(b.cat1, T1, d1, c0=d1, ctot=2 * d1)               # skip half of up_modules.1's input
    crb("down_modules.1.0.", D1, d0, d1, T1, b.U1, b.R1, X1a)
    crb("down_modules.1.1.", X1a, d1, d1, T1, b.U1, None, h1)
    D2 = _View(b.D2, T2, d1)
    _conv(plan, W, B, h1, D2, T_["ds1.w"], T_["ds1.b"], taps=ds_taps, cin_pad=d1, n=d1, t_out=T2, phases=2,
          tag=f"{tag}.down1.downsample")
    X2a, X2b = _View(b.X2a, T2, d2), _View(b.X2b, T2, d2)
    h2 = _View(b.cat0, T2, d2, c0=d2, ctot=2 * d2)               # skip half of up_modules.0's input
    crb("down_modules.2.0.", D2, d1, d2, T2, b.U2, None, X2a)
    crb("down_modules.2.1.", X2a, d2, d2, T2, b.U2, None, h2)
    crb("mid_modules.0.", h2, d2, d2, T2, b.U2, None, X2a)
    x_half0 = _View(b.cat0, T2, d2, c0=0, ctot=2 * d2)
    crb("mid_modules.1.", X2a, d2, d2, T2, b.U2, None, x_half0)
    cat0 = _View(b.cat0, T2, 2 * d2)
    crb("up_modules.0.0.", cat0, 2 * d2, d1, T2, b.U2, b.R2, X2a)
    crb("up_modules.0.1.", X2a, d1, d1, T2, b.U2, None, X2b)
    x_half1 = _View(b.cat1, T1, d1, c0=0, ctot=2 * d1)
    for ph, taps in ((0, [(0, 0), (0, -1)]), (1, [(0, 1), (0, 0)])):   # ConvTranspose1d(k4, s2, p1) by output phase
        _conv(plan, W, B, X2b, x_half1, T_[f"us0.w{ph}"], T_["us0.b"], taps=taps, cin_pad=d1, n=d1, t_out=T2,
              out_rows=(T1, 2, ph, T1), tag=f"{tag}.up0.upsample.phase{ph}")
    cat1 = _View(b.cat1, T1, 2 * d1)
    Y1a, Y1b = _View(b.Y1a, T1, d0), _View(b.Y1b, T1, d0)
    crb("up_modules.1.0.", cat1, 2 * d1, d0, T1, b.V1, b.Q1, Y1a)
    crb("up_modules.1.1.", Y1a, d0, d0, T1, b.V1, None, Y1b)
    for ph, taps in ((0, [(0, 0), (0, -1)]), (1, [(0, 1), (0, 0)])):
        _conv(plan, W, B, Y1b, X0a, T_[f"us1.w{ph}"], T_["us1.b"], taps=taps, cin_pad=d0, n=d0, t_out=T1,
              out_rows=(T0, 2, ph, T0), tag=f"{tag}.up1.upsample.phase{ph}")
    _conv(plan, W, B, X0a, X0b, T_["final0.w"], T_["final0.b"], taps=k5, cin_pad=d0, n=d0, t_out=T0,
          gn=(T_["final0.gw"], T_["final0.gb"]), tag=f"{tag}.final_conv.0+gn+mish")
    _conv(plan, W, B, X0b, None, T_["final1.w"], T_["final1.b"], taps=k1, cin_pad=d0, n=W.A, t_out=T0, bn=32,
          out_f32=b.out, tag=f"{tag}.final_conv.1")


def build_time_film(plan: Plan, W: UnetWeights, t_rows: torch.Tensor, rows: int, cond: Optional[torch.Tensor],
                    film_out: torch.Tensor, tag: str = "film") -> None:
    """FiLM table rows for `rows` time values.

    cond given  ([rows][cond_dim] fp32): film_out[G][rows][11264] = W_full Mish(cat(temb, cond)) + b  (per-sample t)
    cond None   : film_out[G][rows][11264] = W_time Mish(temb)                                          (sampler steps)
    """
    m, G, T_ = W.mode, W.G, W.t
    emb = plan.buf(f"{tag}.emb", (rows, m.ld(DSED)), m.tdt)
    plan.add(_tembed(t_rows, rows, emb, m), f"{tag}.sinusoid")
    hid = plan.buf(f"{tag}.hid", (G, rows, m.ld(4 * DSED)), m.tdt)
    plan.add(linear_desc(a=emb, rows=rows, k=DSED, a_ld=m.ld(DSED), w=T_["time1.w"], n=4 * DSED, n_pad=4 * DSED,
                         w_ld=T_["time1.w"].shape[-1], out=hid, ldc=m.ld(4 * DSED), bias=T_["time1.b"], act=nv.ACT_MISH,
                         G=G, a_G=1, out_g=rows * m.ld(4 * DSED), out_plane=m.plane(4 * DSED), passes=m.passes,
                         a_plane=m.plane(DSED), w_plane=DSED if m.precise else 0), f"{tag}.time_mlp.0+mish")
    kdim = DSED + (W.cond_dim if cond is not None else 0)
    mgf = plan.buf(f"{tag}.mgf", (G, rows, m.ld(kdim)), m.tdt)          # Mish(gf)
    plan.add(linear_desc(a=hid, rows=rows, k=4 * DSED, a_ld=m.ld(4 * DSED), w=T_["time2.w"], n=DSED, n_pad=DSED,
                         w_ld=T_["time2.w"].shape[-1], out=mgf, ldc=m.ld(kdim), bias=T_["time2.b"], act=nv.ACT_MISH, G=G,
                         a_G=G, a_sG=rows * m.ld(4 * DSED), out_g=rows * m.ld(kdim), out_plane=m.plane(kdim),
                         passes=m.passes, a_plane=m.plane(4 * DSED), w_plane=4 * DSED if m.precise else 0),
             f"{tag}.time_mlp.1+mish")
    if cond is not None:
        for g in range(G):
            d = nv.PackDesc()
            d.src, d.src_ld, d.rows, d.cols, d.act = ptr(cond), cond.shape[-1], rows, W.cond_dim, nv.ACT_MISH
            d.out, d.out_dtype, d.out_ld, d.dst_c0 = ptr(mgf, g * rows * m.ld(kdim)), m.dt, m.ld(kdim), DSED
            d.out_plane, d.zero_to = m.plane(kdim), 0
            plan.add(d, f"{tag}.mish(cond).g{g}")
        wname, bias = "film.w_full", T_["film.b"]
    else:
        wname, bias = "film.w_time", T_["film.zero_b"]
    plan.add(linear_desc(a=mgf, rows=rows, k=kdim, a_ld=m.ld(kdim), w=T_[wname], n=FILM_ROWS, n_pad=FILM_ROWS,
                         w_ld=T_[wname].shape[-1], out=film_out, ldc=FILM_ROWS, bias=bias, G=G, a_G=G,
                         a_sG=rows * m.ld(kdim), out_g=rows * FILM_ROWS, passes=m.passes, a_plane=m.plane(kdim),
                         w_plane=kdim if m.precise else 0), f"{tag}.film_gemm")


def build_cond_film(plan: Plan, W: UnetWeights, cond: torch.Tensor, B: int, film_c: torch.Tensor, tag: str = "filmc") -> None:
    """film_c[G][B][11264] = W_cond Mish(cond) + b : the per-call half of the sampler's FiLM (cond fp32 [B][cond_dim])."""
    m, G, T_ = W.mode, W.G, W.t
    mc = plan.buf(f"{tag}.mcond", (B, m.ld(W.cond_dim)), m.tdt)
    d = nv.PackDesc()
    d.src, d.src_ld, d.rows, d.cols, d.act = ptr(cond), cond.shape[-1], B, W.cond_dim, nv.ACT_MISH
    d.out, d.out_dtype, d.out_ld, d.dst_c0, d.out_plane, d.zero_to = ptr(mc), m.dt, m.ld(W.cond_dim), 0, m.plane(W.cond_dim), 0
    plan.add(d, f"{tag}.mish(cond)")
    plan.add(linear_desc(a=mc, rows=B, k=W.cond_dim, a_ld=m.ld(W.cond_dim), w=T_["film.w_cond"], n=FILM_ROWS,
                         n_pad=FILM_ROWS, w_ld=T_["film.w_cond"].shape[-1], out=film_c, ldc=FILM_ROWS, bias=T_["film.b"],
                         G=G, a_G=1, out_g=B * FILM_ROWS, passes=m.passes, a_plane=m.plane(W.cond_dim),
                         w_plane=W.cond_dim if m.precise else 0), f"{tag}.film_cond_gemm")


def _tembed(t_rows: torch.Tensor, rows: int, out: torch.Tensor, m: Mode) -> nv.TembedDesc:
    d = nv.TembedDesc()
    d.t, d.rows, d.dim, d.out, d.out_dtype, d.out_ld, d.out_plane = ptr(t_rows), rows, DSED, ptr(out), m.dt, m.ld(DSED), m.plane(DSED)
    return d


def xpad_desc(W: UnetWeights, x: torch.Tensor, rows: int, bufs: UnetBuffers) -> nv.PackDesc:
    """x fp32 [rows][A] -> the channel-padded operand copy bufs.xpad."""
    m = W.mode
    d = nv.PackDesc()
    d.src, d.src_ld, d.rows, d.cols, d.act = ptr(x), W.A, rows, W.A, nv.ACT_NONE
    d.out, d.out_dtype, d.out_ld, d.dst_c0, d.out_plane, d.zero_to = ptr(bufs.xpad), m.dt, m.ld(W.cin0), 0, m.plane(W.cin0), 0
    return d


class UnetProgram:
    """G nets evaluated on (sample [B,T,A], timestep [B], global_cond [B,cond]) -- the general forward used by tests
    and by the loss; mirrors DiffusionConditionalUnet1D.forward (conditional_unet_1D.py:194-247)."""

    def __init__(self, sds: Sequence[SD], action_dim: int, B: int, T: int, device, precise: bool = False):
        self.W = UnetWeights(sds, action_dim, device, precise)
        self.plan = Plan(device)
        self.W.register(self.plan)
        p = self.plan
        self.B, self.T, self.A = B, T, action_dim
        self.x = p.buf("in.x", (B, T, action_dim), torch.float32)
        self.t = p.buf("in.t", (B,), torch.float32)
        self.cond = p.buf("in.cond", (B, self.W.cond_dim), torch.float32)
        self.film = p.buf("film", (self.W.G, B, FILM_ROWS), torch.float32)
        self.bufs = UnetBuffers(p, self.W, B, T)
        p.add(xpad_desc(self.W, self.x, B * T, self.bufs), "xpad")
        build_time_film(p, self.W, self.t, B, self.cond, self.film)
        build_unet_eval(p, self.W, self.bufs, self.film, None)

    def __call__(self, sample: torch.Tensor, timestep: torch.Tensor, global_cond: torch.Tensor) -> torch.Tensor:
        """-> [G][B][T][A] fp32"""
        self.x.copy_(sample)
        self.t.copy_(timestep.expand(self.B))
        self.cond.copy_(global_cond)
        self.plan.compile().run()
        return self.bufs.out


class LossProgram:
    """StochasticInterpolants.get_loss forward value (bridge_model.py:220-257,183-218): q_sample, then b_net / v_net / s_net
    evaluated as one grouped program (G = 3) on the same (x_t, t, cond), then the three loss reductions.
    Inputs: x0 (vla_act), x1 (expert_act) [B,T,A], cond [B,cond], step [B] ~ U(0,1), z_unit [B,T,A] ~ N(0,1)."""

    def __init__(self, sds_bvs: Sequence[SD], action_dim: int, B: int, T: int, beta_max: float, device, precise: bool = False):
        assert len(sds_bvs) == 3, "expects [b_net, v_net, s_net]"
        self.W = UnetWeights(sds_bvs, action_dim, device, precise)
        self.plan = p = Plan(device)
        self.W.register(p)
        m, A = self.W.mode, action_dim
        self.B, self.T, self.A = B, T, A
        f32 = torch.float32
        self.x0 = p.buf("in.x0", (B, T, A), f32)
        self.x1 = p.buf("in.x1", (B, T, A), f32)
        self.cond = p.buf("in.cond", (B, self.W.cond_dim), f32)
        self.step = p.buf("in.step", (B,), f32)
        self.z = p.buf("in.z_unit", (B, T, A), f32)
        self.xt = p.buf("xt", (B, T, A), f32)
        self.tclip = p.buf("tclip", (B,), f32)
        self.film = p.buf("film", (3, B, FILM_ROWS), f32)
        self.bufs = UnetBuffers(p, self.W, B, T)
        self.per_sample = p.buf("loss.per_sample", (3, B), f32)
        self.out = p.buf("loss.out", (4,), f32)
        d = nv.QsampleDesc()
        d.x0, d.x1, d.step, d.z_unit, d.d = ptr(self.x0), ptr(self.x1), ptr(self.step), ptr(self.z), beta_max
        d.B, d.n, d.A, d.xt, d.tclip = B, T * A, A, ptr(self.xt), ptr(self.tclip)
        d.xpad, d.xpad_dtype, d.xpad_ld, d.xpad_plane = ptr(self.bufs.xpad), m.dt, m.ld(self.W.cin0), m.plane(self.W.cin0)
        p.add(d, "q_sample")
        build_time_film(p, self.W, self.tclip, B, self.cond, self.film)
        build_unet_eval(p, self.W, self.bufs, self.film, None)
        d = nv.SilossDesc()
        d.bvs, d.x0, d.x1, d.z_unit, d.tclip, d.d = ptr(self.bufs.out), ptr(self.x0), ptr(self.x1), ptr(self.z), ptr(self.tclip), beta_max
        d.B, d.n, d.per_sample, d.out = B, T * A, ptr(self.per_sample), ptr(self.out)
        p.add(d, "si_losses")

    def refresh(self, sds_bvs: Sequence[SD]) -> None:
        self.W.refresh(sds_bvs)

    def __call__(self, x0, x1, cond, step, z_unit) -> torch.Tensor:
        self.x0.copy_(x0); self.x1.copy_(x1); self.cond.copy_(cond); self.step.copy_(step); self.z.copy_(z_unit)
        self.plan.compile().run()
        return self.out
