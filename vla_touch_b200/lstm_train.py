"""Training plans of the residual-LSTM controller (row a12 of SURVEY 8, lstm_train.py:70-130: get_loss(...).backward()).

`LstmLayerTrain`: one nn.LSTM layer (lstm_step_controller.py:66-73,196-204) forward-for-training and back-propagation through
time; `LstmLossBackwardProgram`: the whole controller (force encoder, two layers, output head, MSE) forward + backward.

    forward   xw = W_ih x + b_ih + b_hh for all steps (one GEMM)  ->  lstm_seq_train_kernel (recurrence; keeps gates and c)
    backward  lstm_bwd_kernel: the sequential part (d gates of every step; dh_{t-1} = d gates_t W_hh)
              d W_ih = d gates^T x,  d W_hh = d gates^T h_{t-1}   two weight-gradient GEMMs over all B*T rows (unet_bwd.conv_wgrad:
                                                                  h_{t-1} is the hidden output read with tap offset -1)
              d b_ih = d b_hh = column sum of d gates;  d x = d gates W_ih   (one GEMM)

              head: ln_gelu_bwd_kernel (LayerNorm + GELU backward from the saved LayerNorm input), linears as dgrad / wgrad
              GEMMs, d obs_cond = sum over time of the head's input gradient (column sum per sample)

`lstm_loss_backward` of oracle/vt_oracle_bwd.py (pinned to the reference's gradient digests) is the checker of the deterministic
(eval-mode) network; training mode adds the reference's dropout (between the LSTM layers and in the head) as stored Philox masks
(dropmask_kernel), checked against torch autograd with the same masks.  Written after the round's
GPU budget ended: the new kernels (lstm_seq_train, lstm_bwd, ln_gelu_bwd, ewise ops 2-4) have been compiled for sm_100a and
checked on the CPU descriptor interpreter only.  TactileLSTMController.get_loss(batch, differentiable=True) uses it.
"""
from __future__ import annotations

import os
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
                 b_hh: torch.Tensor, B: int, T: int, tag: str = "lstm.l0", y: Optional[torch.Tensor] = None,
                 y_f32: bool = False):
        H, dev, f32, bf = H_LSTM, plan.device, torch.float32, torch.bfloat16
        assert w_hh.shape == (4 * H, H) and x.dtype == bf and x.shape[-1] % 64 == 0
        self.plan, self.x, self.k_in, self.k_pad, self.B, self.T, self.tag = plan, x, k_in, x.shape[-1], B, T, tag
        packed = self._pack_weights(w_ih, w_hh, b_ih, b_hh)
        # forward operand [4H][k_pad], d x operand [k_pad][4H], b_ih + b_hh, W_hh [4H][H] (backward) and W_hh^T [H][4H] (forward)
        self.w_ih, self.w_ih_t, self.b_sum, self.w_hh, self.w_hh_t, self.w_hh_tc, self.w_hh_t_tc = (plan.reg(t) for t in packed)
        self.h_tc = plan.buf(f"{tag}.h_tc", (B, T, H), bf)                 # tensor-core recurrence: its bf16 copy of h (vt_lstm_tc.cuh)
        R = B * T
        self.xw = plan.buf(f"{tag}.xw", (R, 4 * H), f32)
        # hidden outputs: bf16 (the next GEMM's operand; may be the first H columns of a wider buffer) or fp32 (when dropout follows)
        self.y = y if y is not None else plan.buf(f"{tag}.y", (B, T, H), f32 if y_f32 else bf)
        self.y_ld = self.y.shape[-1]
        self.gates = plan.buf(f"{tag}.gates", (B, T, 4 * H), f32)
        self.c = plan.buf(f"{tag}.c", (B, T, H), f32)
        self.grads: Dict[str, torch.Tensor] = {}
        self.dx: Optional[torch.Tensor] = None

    def _pack_weights(self, w_ih, w_hh, b_ih, b_hh):
        H, dev, f32, bf = H_LSTM, self.plan.device, torch.float32, torch.bfloat16
        wp = torch.zeros(4 * H, self.k_pad, device=dev)
        wp[:, : self.k_in] = w_ih.detach().to(dev, f32)
        whh = w_hh.detach().to(dev, f32)
        return (wp.to(bf).contiguous(), wp.t().contiguous().to(bf), (b_ih + b_hh).detach().to(dev, f32).contiguous(),
                whh.contiguous(), whh.t().contiguous(),
                whh.view(4, H, H).permute(1, 0, 2).reshape(4 * H, H).to(bf).contiguous(),      # rows regrouped per unit: unit * 4 + gate
                whh.t().contiguous().to(bf))                                                    # W_hh^T [H][4H]

    def refresh(self, w_ih, w_hh, b_ih, b_hh) -> None:
        """New parameter values -> the same device tensors (after an optimizer step)."""
        for dst, src in zip((self.w_ih, self.w_ih_t, self.b_sum, self.w_hh, self.w_hh_t, self.w_hh_tc, self.w_hh_t_tc),
                            self._pack_weights(w_ih, w_hh, b_ih, b_hh)):
            dst.copy_(src)

    def forward(self) -> torch.Tensor:
        p, H, R = self.plan, H_LSTM, self.B * self.T
        p.add(linear_desc(a=self.x, rows=R, k=self.k_pad, a_ld=self.k_pad, w=self.w_ih, n=4 * H, n_pad=4 * H, w_ld=self.k_pad,
                          out=self.xw, ldc=4 * H, bias=self.b_sum), f"{self.tag}.input_proj")
        d = nv.LstmTrainDesc()
        d.xw, d.w_hh, d.y, d.y_ld = ptr(self.xw), ptr(self.w_hh_t), ptr(self.y), self.y_ld
        d.y_dtype = nv.VT_BF16 if self.y.dtype == torch.bfloat16 else nv.VT_F32
        d.gates, d.c, d.B, d.T, d.H = ptr(self.gates), ptr(self.c), self.B, self.T, H
        if os.environ.get("VT_LSTM_TC", "1") != "0":
            d.w_hh_tc, d.h_tc = ptr(self.w_hh_tc), ptr(self.h_tc)
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
        dg_bf = None
        if os.environ.get("VT_LSTM_TC", "1") != "0":
            dg_bf = p.buf(f"{tag}.dgates_bf16", (B, T, 4 * H), torch.bfloat16)      # the tensor-core BPTT kernel's bf16 copy of d gates
            d.w_hh_t_tc, d.dg_tc = ptr(self.w_hh_t_tc), ptr(dg_bf)
        p.add(d, f"{tag}.bptt")
        ctx = ub.DgradCtx(1, precise=False)
        V = _View
        dgv = V(dg.view(1, B, T, 4 * H), T, 4 * H)                                     # fp32 source: tcol converts to bf16
        memo: dict = {}                                                                # both weight gradients read the same d gates^T
        dw_ih = ub.conv_wgrad(p, ctx, B, dgv, V(self.x.view(1, B, T, self.k_pad), T, self.k_pad), tap_off=[0], t_out=T,
                              tag=f"{tag}.weight_ih.wgrad", rows_t_memo=memo)
        dw_hh = ub.conv_wgrad(p, ctx, B, dgv, V(self.y.view(1, B, T, self.y_ld), T, H), tap_off=[-1], t_out=T,
                              tag=f"{tag}.weight_hh.wgrad", rows_t_memo=memo)          # h_{t-1}: zero at t = 0
        db = ub.colsum(p, 1, B, dgv, T, f"{tag}.bias.colsum")
        self.grads = {"weight_ih": dw_ih[0, :, : self.k_in], "weight_hh": dw_hh[0], "bias_ih": db[0], "bias_hh": db[0]}
        if need_dx:
            # the tensor-core BPTT kernel (host: B >= 16, 16-byte aligned dy rows) already wrote d gates as bf16
            tc_bptt = dg_bf is not None and p.device.type == "cuda" and B >= 16 and dy.shape[-1] % 4 == 0
            dgb = V(dg_bf.view(1, B, T, 4 * H), T, 4 * H) if tc_bptt else ub.cast_bf16(p, 1, B, dgv, T, f"{tag}.dgates.bf16")
            self.dx = p.buf(f"{tag}.dx", (B, T, self.k_pad), torch.float32)
            p.add(linear_desc(a=dgb.t, rows=R, k=4 * H, a_ld=4 * H, w=self.w_ih_t, n=self.k_pad, n_pad=self.k_pad,
                              w_ld=4 * H, out=self.dx, ldc=self.k_pad), f"{tag}.dx")
        return self.dx


def lstm_layers_train(plan: Plan, x: torch.Tensor, k_in: int, lstm_sd: Dict[str, torch.Tensor], B: int, T: int,
                      num_layers: int = 2, last_y: Optional[torch.Tensor] = None) -> Sequence[LstmLayerTrain]:
    """The stacked layers of nn.LSTM(num_layers) in eval-equivalent training (no inter-layer dropout yet): forward ops of all
    layers are appended to `plan`; call `.backward(dy)` on them in reverse order."""
    layers = []
    for l in range(num_layers):
        lay = LstmLayerTrain(plan, x, k_in, lstm_sd[f"weight_ih_l{l}"], lstm_sd[f"weight_hh_l{l}"], lstm_sd[f"bias_ih_l{l}"],
                             lstm_sd[f"bias_hh_l{l}"], B, T, tag=f"lstm.l{l}", y=last_y if l == num_layers - 1 else None)
        x = lay.forward()
        k_in = H_LSTM
        layers.append(lay)
    return layers


# ------------------------------------------------------------------------------------------------
# the whole controller: force encoder -> LSTM -> output head -> MSE loss, forward + backward as one program
# ------------------------------------------------------------------------------------------------
class _Packer:
    """Packs parameters into plan-owned operand tensors and remembers how, so that `refresh(mods)` can re-pack new values into
    the same device tensors.  `get(mods)` selects the raw parameter from {'force_encoder': sd, 'lstm': sd, 'output_head': sd}."""

    def __init__(self, plan: Plan, mods):
        self.plan, self.mods, self.items = plan, mods, []

    def _add(self, fn) -> torch.Tensor:
        t = self.plan.reg(fn(self.mods))
        self.items.append((t, fn))
        return t

    def lin(self, get, k_pad: int, transpose: bool = False) -> torch.Tensor:
        """nn.Linear weight [N, K] -> bf16 operand [N][k_pad] (zero padded); transpose: [K][k_pad >= N] for the d-input GEMM."""
        dev = self.plan.device

        def fn(mods):
            w = get(mods).detach().to(dev, torch.float32)
            if transpose:
                w = w.t()
            out = torch.zeros((w.shape[0] + 31) // 32 * 32, k_pad, device=dev)     # rows padded to the narrowest tile width
            out[: w.shape[0], : w.shape[1]] = w
            return out.to(torch.bfloat16).contiguous()
        return self._add(fn)

    def vec(self, get) -> torch.Tensor:
        dev = self.plan.device

        def fn(mods):
            v = get(mods).detach().to(dev, torch.float32).reshape(-1)
            out = torch.zeros((v.numel() + 255) // 256 * 256, device=dev)          # padded to the widest tile: no masked loads needed
            out[: v.numel()] = v
            return out
        return self._add(fn)

    def refresh(self, mods) -> None:
        for t, fn in self.items:
            t.copy_(fn(mods))


def _pack(plan: Plan, src: torch.Tensor, src_ld: int, rows: int, cols: int, out: torch.Tensor, out_ld: int, dst_c0: int, act: int,
          tag: str, zero_to: int = 0, src_row_div: int = 0) -> None:
    d = nv.PackDesc()
    d.src, d.src_ld, d.rows, d.cols, d.act = ptr(src), src_ld, rows, cols, act
    d.out, d.out_dtype, d.out_ld, d.dst_c0 = ptr(out), nv.VT_BF16 if out.dtype == torch.bfloat16 else nv.VT_F32, out_ld, dst_c0
    d.out_plane, d.zero_to, d.src_row_div = 0, zero_to, src_row_div
    plan.add(d, tag)


def _ew(plan: Plan, a, a_ld: int, b, b_ld: int, out: torch.Tensor, rows: int, cols: int, op: int, tag: str, alpha: float = 0.0):
    e = nv.EwiseDesc()
    e.a, e.a_ld, e.b, e.b_ld, e.out, e.out_ld, e.rows, e.cols, e.op, e.alpha = a, a_ld, b, b_ld, ptr(out), out.shape[-1], rows, cols, op, alpha
    plan.add(e, tag)


class LstmLossBackwardProgram:
    """TactileLSTMController.get_loss(...) and its backward (lstm_step_controller.py:170-204, 321-337; lstm_train.py:120-130) as
    one program, in eval-equivalent training (dropout off, like `lstm_loss_backward` of oracle/vt_oracle_bwd.py):

        forces -> Linear(F,128) GELU Linear(128,128) -> cat(., vla_n) -> LSTM x 2 -> cat(., obs_cond) -> Linear(512,256)
               -> LayerNorm -> GELU -> Linear(256,A) = delta;  out = vla_n + delta;  loss = mean((out - expert)^2)

    Inputs: vla (normalised) [B,T,A], forces [B,T,F], cond [B,256], expert [B,T,A].  Outputs: .loss() (python float),
    .grads {'force_encoder.0.weight', ..., 'lstm.weight_hh_l1', ..., 'output_head.4.bias'} (18 tensors), .d_cond [B,256]."""

    def __init__(self, mods: Dict[str, Dict[str, torch.Tensor]], A: int, Fd: int, B: int, T: int, device, dropout: float = 0.0,
                 inject_uniforms: bool = False):
        """dropout = p > 0 (or a pair (p_lstm, p_head)): training-mode semantics of the reference (nn.LSTM(dropout=p) between the two layers, nn.Dropout(p) in
        the head): two inverted-dropout masks per step from Philox (seed = self.seed[0], set it per step) or, for parity tests,
        from the injected uniforms self.u[0] / self.u[1]."""
        self.plan = p = Plan(device)
        self.p_drop = tuple(float(x) for x in dropout) if isinstance(dropout, (tuple, list)) else (float(dropout), float(dropout))
        H, R, f32, bf = H_LSTM, B * T, torch.float32, torch.bfloat16
        self.A, self.B, self.T = A, B, T
        self.pk = pk = _Packer(p, mods)
        FE, HD = (lambda k: (lambda m: m["force_encoder"][k])), (lambda k: (lambda m: m["output_head"][k]))
        V = _View
        ctx = ub.DgradCtx(1, precise=False)
        r64 = lambda n: (n + 63) // 64 * 64
        fpad, apad, kin = r64(Fd), r64(A), H // 2 + A
        kin_pad = r64(kin)
        self.vla, self.expert = p.buf("in.vla", (B, T, A), f32), p.buf("in.expert", (B, T, A), f32)
        self.forces, self.cond = p.buf("in.forces", (B, T, Fd), f32), p.buf("in.cond", (B, H), f32)
        # ---------------- forward ----------------
        f_op = p.buf("f_op", (R, fpad), bf)
        _pack(p, self.forces, Fd, R, Fd, f_op, fpad, 0, nv.ACT_NONE, "lstm.force->operand", zero_to=fpad)
        a1 = p.buf("fe.a1", (R, H // 2), f32)
        p.add(linear_desc(a=f_op, rows=R, k=fpad, a_ld=fpad, w=pk.lin(FE("0.weight"), fpad), n=H // 2, n_pad=H // 2, w_ld=fpad,
                          out=a1, ldc=H // 2, bias=pk.vec(FE("0.bias"))), "force_encoder.0")
        g1 = p.buf("fe.g1", (R, H // 2), bf)
        _pack(p, a1, H // 2, R, H // 2, g1, H // 2, 0, nv.ACT_GELU, "force_encoder.gelu")
        lin = p.buf("lstm_in", (B, T, kin_pad), bf)
        p.add(linear_desc(a=g1, rows=R, k=H // 2, a_ld=H // 2, w=pk.lin(FE("2.weight"), H // 2), n=H // 2, n_pad=H // 2,
                          w_ld=H // 2, out=lin, ldc=kin_pad, bias=pk.vec(FE("2.bias"))), "force_encoder.2 -> lstm_in[:, :128]")
        _pack(p, self.vla, A, R, A, lin, kin_pad, H // 2, nv.ACT_NONE, "lstm.cat.vla")
        head_in = p.buf("head_in", (B, T, 2 * H), bf)                                # cat(lstm_out, obs_cond broadcast over T)
        sd = mods["lstm"]
        lw = lambda l: (sd[f"weight_ih_l{l}"], sd[f"weight_hh_l{l}"], sd[f"bias_ih_l{l}"], sd[f"bias_hh_l{l}"])
        drop = max(self.p_drop) > 0                                                  # (between the LSTM layers, in the head)
        self.masks, self.u = [], []
        if drop:
            self.seed = p.buf("dropout.seed", (1,), torch.int64)
            for i, nm in enumerate(("lstm", "head")):
                m_ = p.buf(f"dropout.mask.{nm}", (R, H), f32)
                u_ = p.buf(f"dropout.u.{nm}", (R, H), f32) if inject_uniforms else None
                dm = nv.DropmaskDesc()
                dm.inject, dm.p, dm.seed, dm.seed_dev, dm.stream, dm.mask, dm.n = ptr(u_), self.p_drop[i], 0, ptr(self.seed), i, ptr(m_), R * H
                p.add(dm, f"dropout.mask.{nm}")
                self.masks.append(m_)
                self.u.append(u_)
        l0 = LstmLayerTrain(p, lin, kin, *lw(0), B, T, tag="lstm.l0", y_f32=drop)
        x1 = l0.forward()
        if drop:                                                                      # nn.LSTM(dropout=p): on layer 0's outputs
            y0d = p.buf("lstm.l0.y_dropped", (R, H), f32)
            _ew(p, ptr(l0.y), H, ptr(self.masks[0]), H, y0d, R, H, nv.EW_MUL, "lstm.l0.dropout")
            x1 = p.buf("lstm.l1.x", (B, T, H), bf)
            _pack(p, y0d, H, R, H, x1, H, 0, nv.ACT_NONE, "lstm.l0.dropout -> bf16")
        l1 = LstmLayerTrain(p, x1, H, *lw(1), B, T, tag="lstm.l1", y=head_in)
        l1.forward()
        self.layers = [l0, l1]
        _pack(p, self.cond, H, R, H, head_in, 2 * H, H, nv.ACT_NONE, "lstm.cat.obs_cond", src_row_div=T)
        z0 = p.buf("head.z0", (R, H), f32)
        p.add(linear_desc(a=head_in, rows=R, k=2 * H, a_ld=2 * H, w=pk.lin(HD("0.weight"), 2 * H), n=H, n_pad=H, w_ld=2 * H, out=z0,
                          ldc=H, bias=pk.vec(HD("0.bias"))), "output_head.0")
        zn = p.buf("head.zn", (R, H), bf)                                            # Linear(256, A)'s operand (after dropout)
        zn_f = p.buf("head.zn_f32", (R, H), f32) if drop else None
        ln_w, ln_b = pk.vec(HD("1.weight")), pk.vec(HD("1.bias"))
        d = nv.LnDesc()
        d.x, d.in_ld, d.in_row_stride, d.rows, d.D, d.gamma, d.beta, d.eps = ptr(z0), H, 1, R, H, ptr(ln_w), ptr(ln_b), 1e-5
        d.out, d.out_dtype, d.out_ld, d.out_plane, d.act = ptr(zn_f if drop else zn), nv.VT_F32 if drop else nv.VT_BF16, H, 0, nv.ACT_GELU
        p.add(d, "output_head.layernorm+gelu")
        if drop:                                                                      # nn.Dropout(p) of the head
            znd = p.buf("head.zn_dropped", (R, H), f32)
            _ew(p, ptr(zn_f), H, ptr(self.masks[1]), H, znd, R, H, nv.EW_MUL, "output_head.dropout")
            _pack(p, znd, H, R, H, zn, H, 0, nv.ACT_NONE, "output_head.dropout -> bf16")
        self.out = p.buf("out", (R, A), f32)
        w4 = pk.lin(HD("4.weight"), H)
        delta = p.buf("head.delta", (R, A), f32)
        p.add(linear_desc(a=zn, rows=R, k=H, a_ld=H, w=w4, n=A, n_pad=w4.shape[0], w_ld=H, out=delta, ldc=A,
                          bias=pk.vec(HD("4.bias"))), "output_head.4")
        _ew(p, ptr(delta), A, ptr(self.vla), A, self.out, R, A, nv.EW_ADD, "out = vla + delta (residual)")
        # ---------------- loss and its derivative ----------------
        dout = p.buf("dout", (R, A), f32)
        _ew(p, ptr(self.out), A, ptr(self.expert), A, dout, R, A, nv.EW_SCALED_DIFF, "mse.bwd", alpha=2.0 / (R * A))
        sq = p.buf("loss.sq", (R, A), f32)
        _ew(p, ptr(dout), A, ptr(dout), A, sq, R, A, nv.EW_MUL, "mse.square")
        self._sq_sum = ub.colsum(p, 1, 1, sq.view(1, 1, R, A), R, "mse.colsum")
        self.n_forward_ops = len(p)
        # ---------------- backward ----------------
        g: Dict[str, torch.Tensor] = {}
        dv = V(dout.view(1, B, T, A), T, A)
        g["output_head.4.weight"] = ub.conv_wgrad(p, ctx, B, dv, V(zn.view(1, B, T, H), T, H), tap_off=[0], t_out=T, tag="head.4.wgrad")[0]
        g["output_head.4.bias"] = ub.colsum(p, 1, B, dv, T, "head.4.dbias")[0]
        dob = ub.cast_bf16(p, 1, B, dv, T, "dout.bf16", c_pad=apad)
        dzn = p.buf("head.dzn", (R, H), f32)
        p.add(linear_desc(a=dob.t, rows=R, k=apad, a_ld=apad, w=pk.lin(HD("4.weight"), apad, transpose=True), n=H, n_pad=H,
                          w_ld=apad, out=dzn, ldc=H), "head.4.dgrad")
        if drop:
            dzn_m = p.buf("head.dzn_masked", (R, H), f32)
            _ew(p, ptr(dzn), H, ptr(self.masks[1]), H, dzn_m, R, H, nv.EW_MUL, "output_head.dropout.bwd")
            dzn = dzn_m
        dz0, d1, d1zh = (p.buf(f"head.{n}", (R, H), f32) for n in ("dz0", "d1", "d1zh"))
        d = nv.LnGeluBwdDesc()
        d.z0, d.dzn, d.gamma, d.beta, d.eps, d.dz0, d.d1, d.d1zh, d.rows, d.D = ptr(z0), ptr(dzn), ptr(ln_w), ptr(ln_b), 1e-5, ptr(dz0), ptr(d1), ptr(d1zh), R, H
        p.add(d, "head.layernorm+gelu.bwd")
        g["output_head.1.weight"] = ub.colsum(p, 1, B, d1zh.view(1, B, T, H), T, "head.ln.dgamma")[0]
        g["output_head.1.bias"] = ub.colsum(p, 1, B, d1.view(1, B, T, H), T, "head.ln.dbeta")[0]
        dz0v = V(dz0.view(1, B, T, H), T, H)
        g["output_head.0.weight"] = ub.conv_wgrad(p, ctx, B, dz0v, V(head_in.view(1, B, T, 2 * H), T, 2 * H), tap_off=[0], t_out=T,
                                                  tag="head.0.wgrad")[0]
        g["output_head.0.bias"] = ub.colsum(p, 1, B, dz0v, T, "head.0.dbias")[0]
        dz0b = ub.cast_bf16(p, 1, B, dz0v, T, "dz0.bf16")
        dcomb = p.buf("head.dcomb", (R, 2 * H), f32)
        p.add(linear_desc(a=dz0b.t, rows=R, k=H, a_ld=H, w=pk.lin(HD("0.weight"), H, transpose=True), n=2 * H, n_pad=2 * H, w_ld=H, out=dcomb,
                          ldc=2 * H), "head.0.dgrad")
        self.d_cond = p.buf("d_cond", (B, H), f32)
        cs = nv.ColsumDesc()
        cs.x, cs.ld, cs.x_g, cs.G, cs.rows, cs.C, cs.out, cs.out_ld = ptr(dcomb, H), 2 * H, T * 2 * H, B, T, H, ptr(self.d_cond), H
        p.add(cs, "d_cond = sum_t dcomb[:, t, H:]")
        dx1 = self.layers[1].backward(dcomb.view(B, T, 2 * H))                       # reads columns [0, H) with row stride 2H
        if drop:
            dy0 = p.buf("lstm.l0.dy", (B, T, H), f32)
            _ew(p, ptr(dx1), H, ptr(self.masks[0]), H, dy0, R, H, nv.EW_MUL, "lstm.l0.dropout.bwd")
            dx1 = dy0
        dx0 = self.layers[0].backward(dx1)
        for l, lay in enumerate(self.layers):
            for k, v in lay.grads.items():
                g[f"lstm.{k}_l{l}"] = v
        dfv = V(dx0.view(1, B, T, kin_pad), T, H // 2)                               # d fenc = first 128 columns of d lstm_in
        g["force_encoder.2.weight"] = ub.conv_wgrad(p, ctx, B, dfv, V(g1.view(1, B, T, H // 2), T, H // 2), tap_off=[0], t_out=T,
                                                    tag="fe.2.wgrad")[0]
        g["force_encoder.2.bias"] = ub.colsum(p, 1, B, dfv, T, "fe.2.dbias")[0]
        dfb = ub.cast_bf16(p, 1, B, dfv, T, "dfenc.bf16")
        dg1 = p.buf("fe.dg1", (R, H // 2), f32)
        p.add(linear_desc(a=dfb.t, rows=R, k=H // 2, a_ld=H // 2, w=pk.lin(FE("2.weight"), H // 2, transpose=True), n=H // 2,
                          n_pad=H // 2, w_ld=H // 2, out=dg1, ldc=H // 2), "fe.2.dgrad")
        da1 = p.buf("fe.da1", (R, H // 2), f32)
        _ew(p, ptr(dg1), H // 2, ptr(a1), H // 2, da1, R, H // 2, nv.EW_GELU_BWD, "fe.gelu.bwd")
        da1v = V(da1.view(1, B, T, H // 2), T, H // 2)
        g["force_encoder.0.weight"] = ub.conv_wgrad(p, ctx, B, da1v, V(f_op.view(1, B, T, fpad), T, fpad), tap_off=[0], t_out=T,
                                                    tag="fe.0.wgrad")[0][:, :Fd]
        g["force_encoder.0.bias"] = ub.colsum(p, 1, B, da1v, T, "fe.0.dbias")[0]
        self.grads = g

    def refresh(self, mods) -> None:
        """New parameter values (after an optimizer step) into the same operand tensors."""
        self.pk.refresh(mods)
        for l, lay in enumerate(self.layers):
            sd = mods["lstm"]
            lay.refresh(sd[f"weight_ih_l{l}"], sd[f"weight_hh_l{l}"], sd[f"bias_ih_l{l}"], sd[f"bias_hh_l{l}"])

    def set_inputs(self, vla_n, forces, cond, expert) -> None:
        self.vla.copy_(vla_n); self.forces.copy_(forces); self.cond.copy_(cond); self.expert.copy_(expert)

    def run(self) -> None:
        self.runs = getattr(self, "runs", 0) + 1
        self.plan.compile().run()

    def loss_tensor(self) -> torch.Tensor:
        """The loss as a 0-d device tensor (no host synchronisation: the training loop reads it when it logs)."""
        return self._sq_sum.sum() * (self.B * self.T * self.A / 4.0)

    def loss(self) -> float:
        """mean((out - expert)^2) from the squared loss derivative: sum(dout^2) * numel / 4"""
        n = self.B * self.T * self.A
        return float(self._sq_sum.sum()) * n / 4.0
