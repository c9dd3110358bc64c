"""Plan builder for the frozen DinoV2 image encoder (reference: visual_encoder.py:9-106 calling HF `Dinov2Model`;
HF = transformers/models/dinov2/modeling_dinov2.py: embeddings HF:57-149, attention HF:199-235, layer HF:367-386,
model tail HF:473-485).

One program encodes `n_calls` image tensors of `B` images each (the two cameras of DiffusionController.encode_images,
bridge_controller.py:86-110) in ONE pass over 2B images, while keeping the reference's per-call batch-global
predicates (`max > 1`, `mean < 0.5`) separate per camera and evaluating them on the device (no host sync).
"""
from __future__ import annotations

import os

from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import native as nv
from .plan import Plan, gemm_desc, linear_desc, pick_bn, ptr, round_up
from .unet import Mode

SD = Dict[str, torch.Tensor]
PATCH = 14


def native_pos_resize(pos_patch: torch.Tensor, s: int, nh: int, nw: int) -> torch.Tensor:
    """[s*s, D] fp32 on the GPU -> [nh*nw, D] with the bicubic kernel of libvt_b200 (HF:57-95)."""
    out = torch.empty(nh * nw, pos_patch.shape[1], dtype=torch.float32, device=pos_patch.device)
    src = pos_patch.contiguous()
    nv.check(nv.lib().vt_pos_embed_resize(src.data_ptr(), s, out.data_ptr(), nh, nw, src.shape[1], nv.current_stream_ptr()))
    return out


class DinoWeights:
    def __init__(self, sd: SD, heads: int, device, precise: bool = False):
        self.mode = m = Mode(precise)
        self.device = torch.device(device)
        self.heads = heads
        d = lambda k: sd[k].detach().to(self.device, torch.float32)
        self.D = D = d("embeddings.cls_token").shape[-1]
        if D != heads * 64:
            raise NotImplementedError(f"head_dim must be 64 (hidden {D}, heads {heads})")
        self.layers = 0
        while f"encoder.layer.{self.layers}.norm1.weight" in sd:
            self.layers += 1
        self.t: Dict[str, torch.Tensor] = {}
        T = self.t
        wp = d("embeddings.patch_embeddings.projection.weight")
        if wp.shape[1] != 3 or wp.shape[2] != PATCH or wp.shape[3] != PATCH:
            raise NotImplementedError("patch projection must be Conv2d(3, D, 14, 14)")
        self.kp = 3 * PATCH * PATCH                           # 588
        self.kp_pad = round_up(self.kp, 64)                   # 640
        w = torch.zeros(D, self.kp_pad, device=self.device)
        w[:, : self.kp] = wp.reshape(D, -1)
        T["patch.w"] = m.pack_w(w)
        T["patch.b"] = d("embeddings.patch_embeddings.projection.bias").contiguous()
        T["cls"] = d("embeddings.cls_token").reshape(D).contiguous()
        self.pos_full = d("embeddings.position_embeddings").reshape(-1, D).contiguous()   # [1 + s*s, D]
        for i in range(self.layers):
            p = f"encoder.layer.{i}."
            a = p + "attention.attention."
            T[f"{i}.ln1.w"], T[f"{i}.ln1.b"] = d(p + "norm1.weight").contiguous(), d(p + "norm1.bias").contiguous()
            T[f"{i}.qkv.w"] = m.pack_w(torch.cat((d(a + "query.weight"), d(a + "key.weight"), d(a + "value.weight"))))
            T[f"{i}.qkv.b"] = torch.cat((d(a + "query.bias"), d(a + "key.bias"), d(a + "value.bias"))).contiguous()
            T[f"{i}.o.w"] = m.pack_w(d(p + "attention.output.dense.weight"))
            T[f"{i}.o.b"] = d(p + "attention.output.dense.bias").contiguous()
            T[f"{i}.ls1"] = d(p + "layer_scale1.lambda1").contiguous()
            T[f"{i}.ln2.w"], T[f"{i}.ln2.b"] = d(p + "norm2.weight").contiguous(), d(p + "norm2.bias").contiguous()
            if p + "mlp.fc1.weight" not in sd:
                raise NotImplementedError("SwiGLU MLP (dinov2-giant, HF:331-345) is not implemented")
            T[f"{i}.fc1.w"] = m.pack_w(d(p + "mlp.fc1.weight"))
            T[f"{i}.fc1.b"] = d(p + "mlp.fc1.bias").contiguous()
            T[f"{i}.fc2.w"] = m.pack_w(d(p + "mlp.fc2.weight"))
            T[f"{i}.fc2.b"] = d(p + "mlp.fc2.bias").contiguous()
            T[f"{i}.ls2"] = d(p + "layer_scale2.lambda1").contiguous()
        T["ln.w"], T["ln.b"] = d("layernorm.weight").contiguous(), d("layernorm.bias").contiguous()
        self._pos_cache: Dict[tuple, torch.Tensor] = {}

    def pos_embed(self, H: int, W: int, resize: Callable) -> torch.Tensor:
        """[1 + (H/14)*(W/14), D] position table for an H x W input (HF:57-95), cached per resolution (weights are frozen)."""
        key = (H, W)
        if key not in self._pos_cache:
            nh, nw = H // PATCH, W // PATCH
            n_pos = self.pos_full.shape[0] - 1
            s = int(n_pos ** 0.5)
            if nh * nw == n_pos and H == W:
                pos = self.pos_full
            else:
                pos = torch.cat((self.pos_full[:1], resize(self.pos_full[1:], s, nh, nw)), dim=0).contiguous()
            self._pos_cache[key] = pos
        return self._pos_cache[key]

    def register(self, plan: Plan) -> None:
        for t in self.t.values():
            plan.reg(t)


class DinoProgram:
    """images (n_calls tensors of [B,H,W,3] uint8/float or [B,3,H,W] float) -> features fp32 [n_calls][B][D]."""

    def __init__(self, plan: Plan, W: DinoWeights, n_calls: int, B: int, H: int, Wd: int, img_dtype: torch.dtype,
                 layout: int, resize: Callable = native_pos_resize, tag: str = "dino", host_flags: bool = False):
        """host_flags: no IMGSTATS ops -- the caller writes `self.flags` ([divide by 255, ImageNet-normalise] per call) itself
        (episode_store.DeviceEpisodeStore caches features for both outcomes of the batch-global predicate)."""
        if H < PATCH or Wd < PATCH:
            raise ValueError(f"image size {H}x{Wd} is smaller than one {PATCH}x{PATCH} patch")
        # like Conv2d(k14, s14), trailing rows/columns that do not fill a patch are ignored (384 -> 27 patches)
        if img_dtype not in (torch.uint8, torch.float32):
            raise ValueError(f"images must be uint8 or float32, got {img_dtype}")
        self.plan, self.W = plan, W
        m, D, T_ = W.mode, W.D, W.t
        self.n_calls, self.B = n_calls, B
        images = n_calls * B
        npatch = (H // PATCH) * (Wd // PATCH)
        N = npatch + 1
        M = images * N
        self.tokens = N
        W.register(plan)
        pos = plan.reg(W.pos_embed(H, Wd, resize))
        shape = (B, H, Wd, 3) if layout == nv.LAYOUT_BHWC else (B, 3, H, Wd)
        self.img = [plan.buf(f"{tag}.img{c}", shape, img_dtype, zero=False) for c in range(n_calls)]
        self.flags = plan.buf(f"{tag}.flags", (n_calls, 4), torch.int32)
        scratch = plan.buf(f"{tag}.stats", (n_calls, 3 * 1024), torch.float32)
        col = plan.buf(f"{tag}.im2col", (images * npatch, m.ld(W.kp_pad)), m.tdt)
        self.h = h = plan.buf(f"{tag}.h", (M, D), torch.float32)
        xn = plan.buf(f"{tag}.xn", (M, m.ld(D)), m.tdt)
        qkv = plan.buf(f"{tag}.qkv", (M, 3 * D), m.tdt)
        ctx = plan.buf(f"{tag}.ctx", (M, m.ld(D)), m.tdt)
        hid = plan.buf(f"{tag}.hid", (M, m.ld(4 * D)), m.tdt)
        self.feat = plan.buf(f"{tag}.feat", (n_calls, B, D), torch.float32)
        self.first_op = len(plan)

        for c in range(n_calls):
            d = nv.ImgStatsDesc()
            d.img, d.dtype, d.count = ptr(self.img[c]), nv.VT_U8 if img_dtype == torch.uint8 else nv.VT_F32, self.img[c].numel()
            d.partial, d.flags = ptr(scratch, c * 3 * 1024), ptr(self.flags, c * 4)
            if not host_flags:
                plan.add(d, f"{tag}.imgstats{c}")
            d = nv.PatchifyDesc()
            d.img, d.dtype, d.layout = ptr(self.img[c]), nv.VT_U8 if img_dtype == torch.uint8 else nv.VT_F32, layout
            d.images, d.H, d.W, d.patch, d.flags = B, H, Wd, PATCH, ptr(self.flags, c * 4)
            d.out, d.out_dtype, d.out_ld = ptr(col, c * B * npatch * m.ld(W.kp_pad)), m.dt, m.ld(W.kp_pad)
            d.out_cols = W.kp_pad
            d.out_plane = m.plane(W.kp_pad)
            plan.add(d, f"{tag}.patchify{c}")
        # patch projection + bias + position embedding, written to token rows 1.. of every image (HF:97-116,139-149)
        rows = images * npatch
        plan.add(gemm_desc(
            a=ptr(col), in_dtype=m.dt, a_C=m.ld(W.kp_pad), a_T=rows, a_B=1, a_ld=m.ld(W.kp_pad), kc=W.kp_pad,
            t_box=min(128, rows), b_box=1, w=ptr(T_["patch.w"]), n_pad=D, w_ld=T_["patch.w"].shape[-1], M=rows, N=D, bn=pick_bn(D, D, m.dt),
            out=ptr(h), out_dtype=nv.VT_F32, ldc=D, row_div=npatch, out_q=N, out_r=1, out_off=1, bias=ptr(T_["patch.b"]),
            res=ptr(pos), ldres=D, res_q=0, res_r=1, res_off=1, passes=m.passes, a_plane=m.plane(W.kp_pad),
            w_plane=W.kp_pad if m.precise else 0), f"{tag}.patch_embed+pos")
        d = nv.ClsDesc()
        d.cls, d.pos, d.h, d.images, d.tokens, d.D = ptr(T_["cls"]), ptr(pos), ptr(h), images, N, D
        plan.add(d, f"{tag}.cls")

        def ln(x_off_rows, rows_, stride, gamma, beta, out, out_dt, out_ld, out_plane, out_off, name):
            d = nv.LnDesc()
            d.x, d.in_ld, d.in_row_stride, d.rows, d.D = ptr(h, x_off_rows * D), D, stride, rows_, D
            d.gamma, d.beta, d.eps = ptr(gamma), ptr(beta), 1e-6
            d.out, d.out_dtype, d.out_ld, d.out_plane, d.act = ptr(out, out_off), out_dt, out_ld, out_plane, nv.ACT_NONE
            plan.add(d, name)

        lin = dict(rows=M, passes=m.passes)
        # DinoV2-S in bf16 on the GPU: fc1 + GELU + fc2 + LayerScale + residual as ONE kernel (hidden activation stays on chip)
        fused_mlp = (not m.precise and D == 384 and plan.device.type == "cuda" and os.environ.get("VT_FUSED_MLP", "1") != "0")
        # ... and the NEXT block's norm1 comes out of the same kernel (the CTA re-reads the rows it has just written from L2), so only
        # the first block runs a standalone norm1
        fused_ln1 = fused_mlp and os.environ.get("VT_FUSED_LN1", "1") != "0"
        # ... and norm2 out of the attention output projection: rowproj_kernel owns whole 384-wide rows (csrc/vt_rowproj.cuh)
        fused_ln2 = fused_mlp and os.environ.get("VT_FUSED_LN2", "1") != "0"
        for i in range(W.layers):
            L = f"{tag}.l{i}."
            if i == 0 or not fused_ln1:
                ln(0, M, 1, T_[f"{i}.ln1.w"], T_[f"{i}.ln1.b"], xn, m.dt, m.ld(D), m.plane(D), 0, L + "norm1")
            plan.add(linear_desc(a=xn, k=D, a_ld=m.ld(D), w=T_[f"{i}.qkv.w"], n=3 * D, n_pad=3 * D,
                                 w_ld=T_[f"{i}.qkv.w"].shape[-1], out=qkv, ldc=3 * D, bias=T_[f"{i}.qkv.b"],
                                 a_plane=m.plane(D), w_plane=D if m.precise else 0, **lin), L + "qkv")
            d = nv.AttnDesc()
            d.qkv, d.ctx, d.in_dtype, d.images, d.tokens, d.heads = ptr(qkv), ptr(ctx), m.dt, images, N, W.heads
            d.ctx_ld, d.ctx_plane = m.ld(D), m.plane(D)
            plan.add(d, L + "attention")
            if fused_ln2:
                d = nv.RowprojDesc()
                d.x, d.ld_x, d.w, d.w_ld = ptr(ctx), m.ld(D), ptr(T_[f"{i}.o.w"]), T_[f"{i}.o.w"].shape[-1]
                d.bias, d.colscale, d.h, d.ld_h, d.rows, d.D = ptr(T_[f"{i}.o.b"]), ptr(T_[f"{i}.ls1"]), ptr(h), D, M, D
                d.ln_gamma, d.ln_beta = ptr(T_[f"{i}.ln2.w"]), ptr(T_[f"{i}.ln2.b"])
                d.ln_out, d.ln_ld, d.ln_eps = ptr(xn), m.ld(D), 1e-6
                plan.add(d, L + "attn_out+ls1+res+norm2")
            else:
                plan.add(linear_desc(a=ctx, k=D, a_ld=m.ld(D), w=T_[f"{i}.o.w"], n=D, n_pad=D, w_ld=T_[f"{i}.o.w"].shape[-1],
                                     out=h, ldc=D, bias=T_[f"{i}.o.b"], colscale=T_[f"{i}.ls1"], res=h, ldres=D,
                                     a_plane=m.plane(D), w_plane=D if m.precise else 0, **lin), L + "attn_out+ls1+res")
                ln(0, M, 1, T_[f"{i}.ln2.w"], T_[f"{i}.ln2.b"], xn, m.dt, m.ld(D), m.plane(D), 0, L + "norm2")
            if fused_mlp:
                d = nv.MlpDesc()
                d.xn, d.ld_x = ptr(xn), m.ld(D)
                d.w1, d.w1_ld, d.b1 = ptr(T_[f"{i}.fc1.w"]), T_[f"{i}.fc1.w"].shape[-1], ptr(T_[f"{i}.fc1.b"])
                d.w2, d.w2_ld, d.b2 = ptr(T_[f"{i}.fc2.w"]), T_[f"{i}.fc2.w"].shape[-1], ptr(T_[f"{i}.fc2.b"])
                d.ls2, d.h, d.ld_h, d.rows, d.D = ptr(T_[f"{i}.ls2"]), ptr(h), D, M, D
                if fused_ln1 and i + 1 < W.layers:
                    d.ln_gamma, d.ln_beta = ptr(T_[f"{i + 1}.ln1.w"]), ptr(T_[f"{i + 1}.ln1.b"])
                    d.ln_out, d.ln_ld, d.ln_eps = ptr(xn), m.ld(D), 1e-6
                plan.add(d, L + ("mlp(fc1+gelu+fc2+ls2+res)+next.norm1" if d.ln_out else "mlp(fc1+gelu+fc2+ls2+res)"))
                continue
            plan.add(linear_desc(a=xn, k=D, a_ld=m.ld(D), w=T_[f"{i}.fc1.w"], n=4 * D, n_pad=4 * D,
                                 w_ld=T_[f"{i}.fc1.w"].shape[-1], out=hid, ldc=m.ld(4 * D), bias=T_[f"{i}.fc1.b"],
                                 act=nv.ACT_GELU, out_plane=m.plane(4 * D), a_plane=m.plane(D), w_plane=D if m.precise else 0,
                                 **lin), L + "fc1+gelu")
            plan.add(linear_desc(a=hid, k=4 * D, a_ld=m.ld(4 * D), w=T_[f"{i}.fc2.w"], n=D, n_pad=D,
                                 w_ld=T_[f"{i}.fc2.w"].shape[-1], out=h, ldc=D, bias=T_[f"{i}.fc2.b"], colscale=T_[f"{i}.ls2"],
                                 res=h, ldres=D, a_plane=m.plane(4 * D), w_plane=4 * D if m.precise else 0, **lin),
                     L + "fc2+ls2+res")
        # final LayerNorm on the CLS rows only (HF:475-478 -> pooler_output = sequence_output[:, 0])
        for c in range(n_calls):
            ln(c * B * N, B, N, T_["ln.w"], T_["ln.b"], self.feat, nv.VT_F32, D, 0, c * B * D, f"{tag}.final_norm.cls{c}")
        self.last_op = len(plan)
