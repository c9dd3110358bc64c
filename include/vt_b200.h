/*
 * vt_b200.h -- C ABI of libvt_b200.so: the sm_100a (B200) kernels of the VLA-Touch action-refinement hot path.
 *
 * Plain C, no torch types.  Every pointer named `*_dev`/inside a descriptor is a DEVICE pointer owned by the
 * caller; `stream` is a cudaStream_t passed as void*.  Every function returns 0 on success or a negative
 * VT_E_* code; vt_last_error() returns a human-readable message for the calling thread.  Nothing throws.
 *
 * The library executes PROGRAMS: ordered lists of pre-encoded kernel launches (TMA tensor maps are encoded once
 * when an op is added).  The Python host (vla_touch_b200/) builds one program per reference function it
 * replaces -- file:line below are into /root/reference/VLA/residual_controller unless prefixed HF: (transformers
 * modeling_dinov2.py, the un-vendored dependency holding the DinoV2 arithmetic):
 *
 *   DINOv2Encoder.forward                      visual_encoder.py:56-106, HF:97-149,199-235,367-386,473-485
 *       -> IMGSTATS, PATCHIFY, GEMM(patch-embed), CLS, then per layer LAYERNORM, GEMM(qkv), ATTENTION,
 *          GEMM(out-proj + LayerScale + residual), LAYERNORM, GEMM(fc1 + GELU), GEMM(fc2 + LayerScale + residual),
 *          final LAYERNORM on the CLS rows
 *   DiffusionController.encode_observation     bridge_controller.py:112-134 (state_encoder :42-48)
 *       -> PACK, 3x GEMM(+GELU)
 *   normalize_actions / denormalize_actions    controller_dataset.py:303-384             -> AFFINE
 *   DiffusionConditionalUnet1D.forward         bridge/networks/conditional_unet_1D.py:194-247 (blocks :40-105)
 *       -> TEMBED, GEMM(time MLP), GEMM(FiLM), 36x GEMM(implicit conv + GroupNorm + Mish + FiLM + residual)
 *   StochasticInterpolants.sample -> sde_vs    bridge/bridge_model.py:259-279,334-387     -> SDE_STEP per step
 *   TactileLSTMController.forward / predict    lstm_step_controller.py:170-286            -> GEMM, LSTM_SEQ, HEAD
 */
#ifndef VT_B200_H_
#define VT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VT_ABI_VERSION 6

enum { VT_OK = 0, VT_E_INVALID = -1, VT_E_CUDA = -2, VT_E_UNSUPPORTED = -3, VT_E_NODEVICE = -4 };
enum { VT_BF16 = 0, VT_F32 = 1, VT_U8 = 2 };
enum { VT_ACT_NONE = 0, VT_ACT_GELU = 1, VT_ACT_MISH = 2 };
enum { VT_EPI_LINEAR = 0, VT_EPI_GN = 1 };
enum { VT_LAYOUT_BHWC = 0, VT_LAYOUT_BCHW = 1 };
#define VT_MAX_TAPS 8

/* ------------------------------------------------------------------------------------------------------------
 * GEMM / implicit-GEMM convolution on the tcgen05 tensor cores (replaces nn.Linear, nn.Conv1d, nn.ConvTranspose1d,
 * nn.Conv2d(k14,s14) + the op that follows them: bias, GELU/Mish, LayerScale+residual, GroupNorm+Mish+FiLM+residual).
 *
 *   out[g][row(m)][n] = epilogue( sum_{tap,c} A[g][b][tq + tap_t[tap]][tap_p[tap]][a_c0 + c] * W[g][n][tap*kc + c] )
 *
 * A is a channels-last activation viewed as 5-D (channel, phase, tq, sample, group); rows outside [0, a_T) read as
 * zero (= the convolution's zero padding).  A logical output row m = b * t_out + t  (b sample, t position).
 * Tiles hold 128 rows: t_box positions x b_box samples (t_box * b_box <= 128); either b_box == 1 and t_box == 128
 * (plain row-major GEMM over a_T rows of one "sample"), or t_box == t_out (whole samples per tile).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct vt_gemm_desc {
  /* A operand */
  const void* a;      /* base of the activation tensor */
  int32_t in_dtype;   /* VT_BF16 (kind::f16) or VT_F32 (kind::tf32); W has the same dtype */
  int32_t a_C;        /* channels visible to TMA from `a` (reads beyond are zero) */
  int32_t a_P;        /* phases: positions are stored as (tq, phase), position = tq * a_P + phase */
  int32_t a_T;        /* tq extent per sample */
  int32_t a_B;        /* samples */
  int32_t a_G;        /* 1: all groups read the same A; else == G */
  int64_t a_ld;       /* elements between consecutive positions */
  int64_t a_sB;       /* elements between samples */
  int64_t a_sG;       /* elements between groups */
  int32_t a_c0;       /* first channel */
  int32_t kc;         /* channels per tap, multiple of 64 (bf16) / 32 (f32) */
  int32_t taps;
  int32_t tap_p[VT_MAX_TAPS];
  int32_t tap_t[VT_MAX_TAPS];
  int32_t t_box, b_box;
  int32_t passes;     /* 1, or 3 = split-tf32 (A_hi*W_hi + A_lo*W_hi + A_hi*W_lo), in_dtype == VT_F32 only */
  int32_t a_plane;    /* channel distance from the hi to the lo plane of A (passes == 3) */
  int32_t w_plane;    /* K distance from the hi to the lo plane of W (passes == 3) */
  /* W operand: [G][n_pad][w_ld] with K contiguous */
  const void* w;
  int32_t n_pad;      /* rows of W per group (a multiple of bn when G > 1) */
  int32_t w_ld;       /* elements per W row (>= taps * kc, multiple of 8) */
  /* problem */
  int32_t G;
  int32_t M;          /* logical rows per group */
  int32_t N;          /* valid output columns per group */
  int32_t bn;         /* tile width: 32, 128, 192 or 256 (VT_EPI_GN: 128 or 256; f32 operands: 32 or 128) */
  /* output mapping: q = m / row_div, rem = m % row_div, out row = q*out_q + rem*out_r + out_off */
  void* out;
  int32_t out_dtype;
  int32_t ldc;
  int64_t out_g;      /* elements between groups */
  int32_t row_div;
  int64_t out_q, out_r, out_off;
  int64_t out_plane;  /* >0 (f32 out): also store the tf32 lo plane at +out_plane elements */
  /* epilogue */
  int32_t epi;        /* VT_EPI_* */
  int32_t act;        /* VT_ACT_* (LINEAR only) */
  const float* bias;      /* [G][n_pad] or null */
  const float* colscale;  /* [N] (LayerScale) or null, LINEAR only */
  const void* res;        /* residual added last: f32 for LINEAR, out_dtype for GN; null = none */
  int32_t ldres;
  int64_t res_g, res_q, res_r, res_off;
  int64_t res_plane;      /* >0 (GN, f32): residual stored as tf32 hi|lo planes, both are added */
  const float* gn_gamma;  /* [G][n_pad] */
  const float* gn_beta;   /* [G][n_pad] */
  int32_t gn_group_ch;    /* channels per GroupNorm group: 32 or 64 */
  float gn_eps;
  const float* film_c;    /* [G][B][film_ld]: per-sample FiLM (scale at column n, shift at film_C + n) or null */
  const float* film_t;    /* [G][film_ld]: batch-independent FiLM part added to film_c, or null */
  int64_t film_g;         /* elements between groups of film_c */
  int64_t film_tg;        /* elements between groups of film_t */
  int32_t film_ld, film_C, film_off;
  float* raw_out;         /* VT_EPI_GN with bf16 in / out, bn 256 (training forward): ALSO store conv + bias, the value GroupNorm
                             sees, as fp32 at raw_out[g * raw_g + m * raw_ld + n] (m = logical row).  It is what the backward of
                             GroupNorm + Mish reads (autograd keeps this tensor for conditional_unet_1D.py:48-52), so the backward
                             program needs no recomputation of the convolution.  NULL = not stored.  (ABI v6) */
  int64_t raw_g;
  int32_t raw_ld;
} vt_gemm_desc;

/* LayerNorm over the last dim (HF:354,359,449; lstm_step_controller.py:76-82).  Row r of the output is computed
 * from input row r * in_row_stride.  D in {256, 384, 768, 1024}. */
typedef struct vt_ln_desc {
  const float* x;
  int64_t in_ld;          /* elements between input rows that are `in_row_stride` = 1 apart */
  int64_t in_row_stride;  /* 1 = every row; N_tokens = CLS rows only */
  int32_t rows, D;
  const float* gamma;
  const float* beta;
  float eps;
  void* out;
  int32_t out_dtype;
  int64_t out_ld;
  int64_t out_plane;      /* >0 (f32): tf32 hi/lo split */
  int32_t act;            /* VT_ACT_GELU applies GELU after the affine (LSTM head) */
} vt_ln_desc;

/* Multi-head softmax(QK^T/sqrt(64))V over packed qkv rows [images*tokens][3*D] (HF:199-235). head_dim == 64. */
typedef struct vt_attn_desc {
  const void* qkv;   /* bf16 [images*tokens][3*D]; for in_dtype F32: f32 */
  void* ctx;         /* [images*tokens][ctx_ld] */
  int32_t in_dtype;  /* VT_BF16: tcgen05 path; VT_F32: fp32 CUDA-core path (parity mode) */
  int32_t images, tokens, heads;
  int64_t ctx_ld;    /* elements between ctx rows (>= D) */
  int64_t ctx_plane; /* >0 (f32): tf32 hi/lo split of ctx */
} vt_attn_desc;

/* Fused ViT MLP block of DinoV2-S (HF Dinov2MLP + layer_scale2 + residual, HF:312-328,380-386):
 *   h += ls2 * (GELU(xn W1^T + b1) W2^T + b2)   with the [rows, 4D] hidden activation kept on chip.  D must be 384. */
typedef struct vt_mlp_desc {
  const void* xn;   /* bf16 [rows][ld_x]: LayerNorm2(h) */
  int64_t ld_x;
  const void* w1;   /* bf16 [4D][w1_ld], K contiguous */
  int64_t w1_ld;
  const float* b1;  /* [4D] */
  const void* w2;   /* bf16 [D][w2_ld], K contiguous */
  int64_t w2_ld;
  const float* b2;  /* [D] */
  const float* ls2; /* [D] LayerScale, or NULL */
  float* h;         /* fp32 [rows][ld_h] residual stream, updated in place */
  int64_t ld_h;
  int32_t rows, D;
  /* optional (ln_out != NULL): LayerNorm(h_new) * ln_gamma + ln_beta as bf16 -- the NEXT block's norm1 (HF:367-372), written by the
     CTA that produced the rows; ln_out may alias xn (a tile's xn rows are dead once its last fc1 MMA has been issued) */
  const float* ln_gamma;
  const float* ln_beta;
  void* ln_out;     /* bf16 [rows][ln_ld] */
  int64_t ln_ld;
  float ln_eps;
} vt_mlp_desc;

/* Attention output projection of a DinoV2-S block with whole 384-wide rows per CTA (HF Dinov2SelfOutput + layer_scale1 + residual,
 * HF:238-251,374-378):  h += colscale * (x W^T + bias);  optionally the block's norm2 (HF:380-381) of the updated rows as bf16 from
 * the same kernel.  D must be 384.  Means exactly what a GEMM descriptor with bias, colscale, res = out = h followed by a LayerNorm descriptor means. */
typedef struct vt_rowproj_desc {
  const void* x;         /* bf16 [rows][ld_x]: attention context */
  int64_t ld_x;
  const void* w;         /* bf16 [D][w_ld], K contiguous (nn.Linear weight) */
  int64_t w_ld;
  const float* bias;     /* [D] */
  const float* colscale; /* [D] LayerScale, or NULL */
  float* h;              /* fp32 [rows][ld_h] residual stream, updated in place */
  int64_t ld_h;
  int32_t rows, D;
  const float* ln_gamma; /* optional (ln_out != NULL): LayerNorm(h_new) * ln_gamma + ln_beta as bf16 */
  const float* ln_beta;
  void* ln_out;          /* bf16 [rows][ln_ld] */
  int64_t ln_ld;
  float ln_eps;
} vt_rowproj_desc;

/* visual_encoder.py:66-81,95-106: batch-global predicates max>1 and mean<0.5, evaluated on the device. */
typedef struct vt_imgstats_desc {
  const void* img;
  int32_t dtype;     /* VT_U8 or VT_F32 */
  int64_t count;     /* elements in the whole call tensor */
  float* partial;    /* scratch, 8-byte aligned: 1024 floats followed by 1024 doubles (12 KiB) */
  int32_t* flags;    /* out: flags[0] = max > 1 (divide by 255), flags[1] = mean >= 0.5 (ImageNet normalise) */
} vt_imgstats_desc;

/* layout fix + /255 + ImageNet normalise + 14x14 im2col (visual_encoder.py:70-106, HF:139-149) */
typedef struct vt_patchify_desc {
  const void* img;
  int32_t dtype, layout;
  int32_t images, H, W, patch;
  const int32_t* flags;
  void* out;          /* [images * (H/patch)*(W/patch)][out_ld], k = c*patch*patch + i*patch + j */
  int32_t out_dtype;
  int32_t out_cols;   /* logical columns written per row (>= 3*patch*patch, zero padded) */
  int32_t out_ld;     /* row pitch in elements */
  int64_t out_plane;
} vt_patchify_desc;

/* h[b][0][:] = cls + pos[0]  (HF:104-112) */
typedef struct vt_cls_desc {
  const float* cls;
  const float* pos;
  float* h;
  int32_t images, tokens, D;
} vt_cls_desc;

/* Generic row packer: out[r][dst_c0 + c] = cast(act(src[r * src_ld + c])), c < cols; used for torch.cat
 * (bridge_controller.py:129-132) and Mish(gf) (conditional_unet_1D.py:76-80). */
typedef struct vt_pack_desc {
  const float* src;
  int64_t src_ld;
  int32_t rows, cols;
  int32_t act;
  void* out;
  int32_t out_dtype;
  int64_t out_ld;
  int32_t dst_c0;
  int64_t out_plane;
  int32_t zero_to;    /* >cols: also zero-fill columns [dst_c0+cols, dst_c0+zero_to) */
  int32_t src_row_div; /* >1: output row r reads source row r / src_row_div (broadcast over time steps) */
} vt_pack_desc;

/* normalize_actions / denormalize_actions, controller_dataset.py:303-384 (padding factor 1.4) */
typedef struct vt_affine_desc {
  const float* x;     /* [rows][A] */
  float* out;         /* [rows][A] */
  const float* mins;  /* [A] */
  const float* maxs;  /* [A] */
  int32_t rows, A;
  int32_t denorm;     /* 0 normalise (with the safe_range guard), 1 de-normalise (without), 2 identity (x + add) */
  float pad;          /* padding factor (reference default 1.4) */
  void* xpad;         /* optional: also write channel-padded copy [rows][xpad_ld] in xpad_dtype */
  int32_t xpad_dtype, xpad_ld;
  int64_t xpad_plane;
  const float* add;   /* optional [rows][A] added after the map (LSTM: vla + delta) */
} vt_affine_desc;

/* SinusoidalPosEmb, conditional_unet_1D.py:7-19: out[r] = cat(sin(t_r f_i), cos(t_r f_i)) */
typedef struct vt_tembed_desc {
  const float* t;
  int32_t rows, dim;
  void* out;
  int32_t out_dtype;
  int64_t out_ld;
  int64_t out_plane;
} vt_tembed_desc;

/* One Euler-Maruyama step of sde_vs, bridge_model.py:352-385:
 *   s' = s*ginv;  b = v - dgg*s'*eps;  x += (b + eps*s')*dt + nscale*(d*z)
 * z = noise[row][a] when noise != null, else Philox4x32-10(seed, step, element). */
typedef struct vt_sde_desc {
  float* x;           /* [rows][A] in/out */
  const float* v;     /* [rows][A] */
  const float* s;     /* [rows][A] */
  const float* noise; /* [rows][A] or null */
  int32_t rows, A;
  float ginv, dgg, eps, dt, nscale, d;
  uint64_t seed;
  const uint64_t* seed_dev; /* optional: seed read from device memory at run time (added to `seed`) */
  int32_t step;
  void* xpad;         /* channel-padded copy of the new x for the next U-Net evaluation */
  int32_t xpad_dtype, xpad_ld;
  int64_t xpad_plane;
} vt_sde_desc;

/* q_sample of the stochastic interpolant, bridge_model.py:103-107,248-257:
 *   t = clip(step, 1e-3, 1-1e-3);  gamma = 1.4142 t (1-t);  z = d * z_unit;  xt = (1-t) x0 + t x1 + gamma z */
typedef struct vt_qsample_desc {
  const float* x0;      /* [B][n] prior (vla) actions, n = T*A */
  const float* x1;      /* [B][n] target (expert) actions */
  const float* step;    /* [B] raw U(0,1) draws */
  const float* z_unit;  /* [B][n] N(0,1) draws */
  float d;              /* beta_max */
  int32_t B, n, A;
  float* xt;            /* [B][n] */
  float* tclip;         /* [B] */
  void* xpad;           /* channel-padded operand copy of xt: [B*T][xpad_ld] */
  int32_t xpad_dtype, xpad_ld;
  int64_t xpad_plane;
} vt_qsample_desc;

/* velocity / score / b losses, bridge_model.py:183-218,240-246, for nets stacked as [b_net, v_net, s_net]:
 *   L_v = mean_b(0.5|v|^2 - <x1-x0, v>),  L_s = mean_b(0.5|s|^2 + <z, s>),  L_b = mean_b(0.5|b|^2 - <x1-x0 + gdot z, b>)
 * with z = d*z_unit, gdot = 1.4142 (1 - 2 t).  out[4] = (L_v + L_s + L_b, L_v, L_s, L_b). */
typedef struct vt_siloss_desc {
  const float* bvs;     /* [3][B][n] outputs of b_net, v_net, s_net */
  const float* x0;
  const float* x1;
  const float* z_unit;
  const float* tclip;   /* [B] */
  float d;
  int32_t B, n;
  float* per_sample;    /* scratch [3][B] */
  float* out;           /* [4] */
} vt_siloss_desc;

/* ------------------------------------------------------------------------------------------------------------
 * Backward of the U-Net's Conv1dBlock = Conv1d -> GroupNorm(8) -> Mish (conditional_unet_1D.py:40-55) and of the FiLM
 * modulation that follows blocks[0] (:97-102).  The data gradients of the convolutions run on vt_gemm_desc unchanged
 * (transposed weight slices, shifted taps); the WEIGHT gradients are plain row-major GEMMs over transposed operand copies
 * (K = B*T rows) made by vt_tcol_desc; everything elementwise is vt_gnbwd_desc.
 * ------------------------------------------------------------------------------------------------------------ */

/* Transposed im2col copy (bf16):  out[g][tap * c_pad + c][b * t_out + t] = src[g][b][t * stride + tap_off[tap]][c],
 * zero where the position falls outside [0, T_src).  Rows c in [C, c_pad) and columns >= B * t_out are never written
 * (the caller zero-initialises the buffer once).
 *   wgrad Conv1d(k, s, p):        dW[co][k][ci] = dY^T (taps = 1, off 0)  x  X^T (tap_off[k] = k - p, stride s)
 *   wgrad ConvTranspose1d(4,2,1): dW[ci][k][co] = X^T                     x  dY^T (tap_off[k] = k - 1, stride 2) */
typedef struct vt_tcol_desc {
  const void* src;      /* channels-last activation, first channel of the window */
  int32_t src_dtype;    /* VT_BF16 or VT_F32 */
  int64_t ld;           /* elements between positions */
  int64_t sB, sG;       /* elements between samples / groups */
  int32_t G, B, T_src, C;
  int32_t taps;
  int32_t tap_off[VT_MAX_TAPS];
  int32_t stride;
  int32_t t_out;        /* positions per sample in the K index */
  void* out;            /* bf16 [G][taps * c_pad][k_ld] */
  int32_t c_pad;
  int64_t k_ld;         /* >= B * t_out, multiple of 64 */
  int64_t out_g;        /* elements between groups of out */
} vt_tcol_desc;

/* Weight gradient of a convolution / linear layer straight from the channels-last activations (no transposed operand copy):
 *   out[g][r][tap * c_pad + c] = sum_{b < B, t < t_out} rows[g][b][t + rows_t (phase rows_p)][r] * cols[g][b][t + tap_t[tap] (phase tap_p[tap])][c]
 * Both operands are bf16 [G][B][positions][ld] tensors whose positions are stored as (tq, phase) pairs like vt_gemm_desc's A operand
 * (position = tq * P + phase; a stride-2 convolution reads one phase); positions outside [0, T) read as zero.  The contraction runs
 * over positions with the operands in the tensor cores' MN-major form (csrc/vt_wgrad.cuh).  t_out must divide 64.
 *   Conv1d(k, s, p):         rows = dY, cols = X  (tap k at position offset k - p, stride s)  ->  [C_out][k][C_in]
 *   ConvTranspose1d(4,2,1):  rows = X,  cols = dY (tap k at offset k - 1, stride 2)           ->  [C_in][k][C_out] */
typedef struct vt_wgrad_desc {
  const void* rows;        /* first channel of the window */
  int32_t rows_C;          /* channels visible from `rows` (reads beyond are zero) */
  int32_t rows_P, rows_T;  /* phases, tq extent per sample */
  int64_t rows_ld, rows_sB, rows_sG;
  int32_t rows_p, rows_t;
  const void* cols;
  int32_t cols_C, cols_P, cols_T;
  int64_t cols_ld, cols_sB, cols_sG;
  int32_t taps;
  int32_t tap_p[VT_MAX_TAPS];
  int32_t tap_t[VT_MAX_TAPS];
  int32_t c_pad;           /* columns per tap, multiple of 64 */
  int32_t G, B, t_out;
  int32_t R;               /* output rows (channels of `rows`) */
  float* out;              /* fp32 [G][R][ldc] */
  int32_t ldc;
  int64_t out_g;
} vt_wgrad_desc;

/* GroupNorm + Mish (+ FiLM) backward from the raw conv output (bias included), per net g and sample b:
 *   xh = (raw - mean) * rstd over each (sample, group);  y = xh * gamma + beta;  m = mish(y)
 *   film != null (forward out = scale * m + shift):  dm = dout * scale,  dfilm[b][film_off + c] = sum_t dout * m,
 *                                                    dfilm[b][film_off + C + c] = sum_t dout;   else dm = dout
 *   da = dm * mish'(y);  dgamma[c] = sum_{b,t} da * xh;  dbeta[c] = sum_{b,t} da;  dxh = da * gamma
 *   draw = rstd * (dxh - mean_g(dxh) - xh * mean_g(dxh * xh));  dbias[c] = sum_{b,t} draw       (two launches) */
typedef struct vt_gnbwd_desc {
  const float* raw;     /* [G][B][T][C] */
  const float* dout;    /* [G][B][T][dout_ld] */
  int64_t dout_ld, dout_g;
  const float* gamma;   /* [G][p_ld] */
  const float* beta;    /* [G][p_ld] */
  int32_t p_ld;
  const float* film;    /* [G][B][film_ld] or null */
  float* dfilm;         /* [G][B][film_ld] or null (same geometry as film) */
  int64_t film_g;
  int32_t film_ld, film_off;
  void* draw;           /* bf16 [G][B][T][C]: the A operand of the conv's dgrad / wgrad */
  float* part;          /* scratch [G][B][3][C] */
  float* dgamma;        /* [G][p_ld] */
  float* dbeta;         /* [G][p_ld] */
  float* dbias;         /* [G][p_ld] */
  int32_t G, B, T, C, groups;
  float eps;
} vt_gnbwd_desc;

/* out[g][c] = sum_r x[g][r][c]: bias gradient of a convolution that has no GroupNorm (down / up-sampling, 1x1 convs). */
typedef struct vt_colsum_desc {
  const float* x;
  int64_t ld, x_g;
  int32_t G, rows, C;
  float* out;           /* [G][out_ld] */
  int32_t out_ld;
} vt_colsum_desc;

/* Strided fp32 elementwise combine over [rows][cols] windows: VT_EW_ADD out = a + b (gradient fan-in where a skip connection
 * and the main path meet, conditional_unet_1D.py:226-240); VT_EW_MISH_BWD out = a * mish'(b) (backward of the Mish in front of
 * the cond_encoder / diffusion_step_encoder linears, :76-80,186-191); VT_EW_GELU_BWD out = a * gelu'(b) (force encoder of the LSTM
 * controller, lstm_step_controller.py:52-58); VT_EW_MUL out = a * b; VT_EW_SCALED_DIFF out = alpha * (a - b) (derivative of
 * F.mse_loss, lstm_step_controller.py:335). */
enum { VT_EW_ADD = 0, VT_EW_MISH_BWD = 1, VT_EW_GELU_BWD = 2, VT_EW_MUL = 3, VT_EW_SCALED_DIFF = 4 };
typedef struct vt_ewise_desc {
  const float* a;
  int64_t a_ld;
  const float* b;
  int64_t b_ld;
  float* out;
  int64_t out_ld;
  int64_t rows;
  int32_t cols;
  int32_t op;
  float alpha;
} vt_ewise_desc;

/* Inverted-dropout mask of nn.Dropout(p) / nn.LSTM(dropout=p) in training mode (lstm_step_controller.py:66-82):
 * mask[i] = u_i >= p ? 1/(1-p) : 0,  u_i ~ U[0,1) from Philox4x32-10(seed (+ *seed_dev), element i, stream) or from `inject`
 * (parity tests).  Forward and backward multiply by the stored mask (VT_EW_MUL). */
typedef struct vt_dropmask_desc {
  const float* inject;      /* [n] uniforms or null */
  float p;
  uint64_t seed;
  const uint64_t* seed_dev; /* optional: added to seed at run time (a new mask per step without re-encoding the program) */
  int32_t stream;
  float* mask;              /* [n] */
  int64_t n;
} vt_dropmask_desc;

/* Backward of LayerNorm(256) -> GELU of the LSTM controller's output head (lstm_step_controller.py:76-82) from the saved
 * LayerNorm input z0:  dz0 = d loss / d z0;  d1 = dzn * gelu'(z1) and d1zh = d1 * zh are stored for the column sums that give
 * d beta and d gamma. */
typedef struct vt_lngelubwd_desc {
  const float* z0;      /* [rows][256] */
  const float* dzn;     /* [rows][256] gradient of the GELU output */
  const float* gamma;
  const float* beta;
  float eps;
  float* dz0;           /* [rows][256] */
  float* d1;            /* [rows][256] */
  float* d1zh;          /* [rows][256] */
  int32_t rows, D;      /* D == 256 */
} vt_lngelubwd_desc;

/* Derivative of the summed loss of vt_siloss_desc with respect to the stacked net outputs (the seed of the backward pass):
 *   dvs[0] = (b - (x1 - x0 + gdot z)) / B,  dvs[1] = (v - (x1 - x0)) / B,  dvs[2] = (s + z) / B     (bridge_model.py:183-246) */
typedef struct vt_silossbwd_desc {
  const float* bvs;     /* [3][B][n] */
  const float* x0;
  const float* x1;
  const float* z_unit;
  const float* tclip;   /* [B] */
  float d;
  int32_t B, n;
  float* dvs;           /* [3][B][n] */
} vt_silossbwd_desc;

/* nn.LSTM (gate order i,f,g,o), lstm_step_controller.py:66-73,196-204: the input projections xw = W_ih x + b_ih
 * + b_hh are precomputed by a GEMM; this op runs the recurrence over T steps for one layer. */
typedef struct vt_lstm_desc {
  const float* xw;    /* [B][T][4H] */
  const float* w_hh;  /* W_hh TRANSPOSED: [H][4H] */
  float* h;           /* [B][H] in/out state */
  float* c;           /* [B][H] in/out state */
  void* y;            /* [B][T][y_ld] hidden outputs */
  int32_t y_dtype;
  int64_t y_ld;
  int64_t y_plane;    /* >0 (f32): tf32 hi/lo split of y */
  int32_t B, T, H;
  /* Tensor-core recurrence (csrc/vt_lstm_tc.cuh: a cluster of 8 CTAs per 128 batch rows keeps W_hh in shared memory for all T steps,
   * tcgen05 MMA per step).  Used when all three are set and B >= 16; the fp32 parity mode and single control ticks leave them null. */
  const void* w_hh_tc; /* bf16 [4H][H], row = unit * 4 + gate (W_hh rows regrouped per hidden unit) */
  void* h_tc;          /* bf16 [B][T][H] scratch: h of every step */
  int32_t zero_init;   /* 1: the state h / c is zero on entry (whole-sequence passes); required by the tensor-core path */
} vt_lstm_desc;

/* Training forward of one nn.LSTM layer from a zero initial state (lstm_step_controller.py:196-204): like vt_lstm_desc, and keeps
 * the activated gates and the cell state of every step for the backward pass. */
typedef struct vt_lstm_train_desc {
  const float* xw;    /* [B][T][4H] input projections W_ih x + b_ih + b_hh */
  const float* w_hh;  /* W_hh TRANSPOSED: [H][4H] */
  void* y;            /* [B][T][y_ld] hidden outputs */
  int32_t y_dtype;
  int64_t y_ld;
  float* gates;       /* [B][T][4H] sigmoid(i), sigmoid(f), tanh(g), sigmoid(o) */
  float* c;           /* [B][T][H] */
  int32_t B, T, H;
  const void* w_hh_tc; /* see vt_lstm_desc: both set and B >= 16 -> tensor-core recurrence */
  void* h_tc;
} vt_lstm_train_desc;

/* Back-propagation through time of one LSTM layer: the sequential part of what torch autograd runs for nn.LSTM in
 * lstm_train.py:130 (loss.backward()).  Writes the gradients of the pre-activation gates of all steps; W_ih / W_hh / bias
 * gradients and d x are GEMMs / column sums over them (vt_tcol_desc + vt_gemm_desc, vt_colsum_desc). */
typedef struct vt_lstm_bwd_desc {
  const float* gates; /* [B][T][4H] from vt_lstm_train_desc */
  const float* c;     /* [B][T][H] */
  const float* dy;    /* [B][T][dy_ld] gradient of the hidden outputs */
  int64_t dy_ld;
  const float* w_hh;  /* W_hh as stored by nn.LSTM: [4H][H] */
  float* dgates;      /* [B][T][4H] */
  int32_t B, T, H;
  const void* w_hh_t_tc; /* bf16 W_hh TRANSPOSED [H][4H]; with dg_tc set and B >= 16 -> tensor-core BPTT (csrc/vt_lstm_tc.cuh) */
  void* dg_tc;           /* bf16 [B][T][4H] scratch: d gates, the recurrence's A operand */
} vt_lstm_bwd_desc;

/* Fused multi-tensor optimizer step of the bridge trainer (bridge_train.py:330-337 + torch_ema update):
 *   g *= grad_scale (1/world after the gradient all-reduce);  AdamW (torch.optim.AdamW semantics, decoupled weight decay,
 *   bias correction with step t);  EMA shadow s -= (1 - ema_decay) (s - p)  when ema != null.
 * `tensors` is a DEVICE array of n records, `chunks` a DEVICE array of n_chunks (tensor index, element offset) pairs, one
 * per thread block (chunk_elems elements each). */
typedef struct vt_opt_tensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* ema;       /* may be null */
  int64_t numel;
  /* Gradient layout.  taps == 0: g is contiguous like p.  taps > 0: p is a convolution weight [rows][c][taps] (nn.Conv1d /
   * nn.ConvTranspose1d order) and g is the weight-gradient GEMM's output [rows][taps][c_pad] (vt_tcol_desc), i.e.
   * g index of p[(r * c + ci) * taps + k] = (r * taps + k) * c_pad + ci: the optimizer reads the training program's buffers
   * in place (no unpack pass, no autograd copy). */
  int32_t taps, c, c_pad;
  int32_t reserved;
  /* optional: the updated weight is also written as the bf16 GEMM operand of the next forward pass, at the same index as the
   * gradient (the forward packing [rows][taps][c_pad] of unet._pack_conv; contiguous when taps == 0) */
  void* w_op;
} vt_opt_tensor;
typedef struct vt_adamw_desc {
  const vt_opt_tensor* tensors;
  const int64_t* chunks;       /* [n_chunks][2] */
  int32_t n_chunks, chunk_elems;
  float lr, beta1, beta2, eps, weight_decay;
  float bias_corr1, bias_corr2; /* 1 - beta1^t, 1 - beta2^t */
  float ema_decay;              /* min(decay, (1+n)/(10+n)) */
  float grad_scale;
} vt_adamw_desc;

/* ------------------------------------------------------------------------------------------------------------
 * Persistent multi-layer launch: a whole U-Net evaluation (DiffusionConditionalUnet1D.forward, conditional_unet_1D.py:194-247)
 * or the whole sampling loop of sde_vs (bridge/bridge_model.py:343-385: n_steps x [evaluation of v_net and s_net,
 * Euler-Maruyama update]) as ONE kernel launch instead of 36 (+1) launches per evaluation.  The layers are the same
 * vt_gemm_desc records the multi-launch form uses; `deps` tells which earlier layers produce a layer's A operand / residual.
 * One CTA pair per TPC walks a static tile schedule; tiles are ordered by per-(layer, net, sample block) counters in global
 * memory (nothing in the U-Net crosses samples), see csrc/vt_persist.cuh.  bf16 operands; GroupNorm layers bn = 256, linear
 * layers bn = 256 or 32; every tile holds whole samples (t_box == positions per sample).
 * ------------------------------------------------------------------------------------------------------------ */
#define VT_PERSIST_MAX_DEPS 4
#define VT_PERSIST_MAX_LAYERS 37
typedef struct vt_persist_desc {
  const vt_gemm_desc* gemms;  /* HOST array: the layers of ONE evaluation, execution order */
  int32_t n_gemms;
  const vt_sde_desc* sde;     /* HOST, optional: the update closing every step (x, v, s, noise, d, seed, xpad; its per-step scalars
                                 come from sde_coef); referred to in `deps` as layer index n_gemms */
  const int32_t* deps;        /* HOST [n_gemms + (sde ? 1 : 0)][VT_PERSIST_MAX_DEPS]: producing layer indices, -1 = unused */
  const int32_t* dep_lag;     /* same shape: 1 = the producer's output of the PREVIOUS step (0 otherwise) */
  int32_t n_steps;            /* >= 1; 1 = a single evaluation */
  int64_t film_t_step;        /* elements between consecutive steps' rows of the layers' film_t tables */
  const float* sde_coef;      /* HOST [n_steps][5]: ginv, dgg, eps, dt, nscale of each step */
  int64_t noise_step;         /* elements between consecutive steps of sde->noise */
  int32_t sde_T;              /* positions per sample of x */
} vt_persist_desc;

typedef struct vt_program vt_program;

const char* vt_last_error(void);
int vt_abi_version(void);
/* sm count / compute capability of the current device; VT_E_NODEVICE without a GPU */
int vt_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

int vt_program_create(vt_program** out);
int vt_program_destroy(vt_program* p);
int vt_program_num_ops(const vt_program* p);
/* number of kernel launches vt_program_run(p, first, count) performs */
int vt_program_num_launches(const vt_program* p, int first, int count);

int vt_program_add_gemm(vt_program* p, const vt_gemm_desc* d);
int vt_program_add_layernorm(vt_program* p, const vt_ln_desc* d);
int vt_program_add_attention(vt_program* p, const vt_attn_desc* d);
int vt_program_add_mlp(vt_program* p, const vt_mlp_desc* d);
int vt_program_add_rowproj(vt_program* p, const vt_rowproj_desc* d);
/* developer instrumentation (VT_GEMM_DEBUG bit 128): (tag, clock64) pairs of one epilogue warp; returns the entry count */
int vt_debug_timestamps(long long* out, int max_entries);
/* developer instrumentation (debug-knobs builds, VT_GEMM_DEBUG bit 512): timeline of the persistent multi-layer kernel.
   out: [workers][tiles][slots] clock64 stamps, cal: [workers][4] = (clock64, globaltimer) at kernel entry and exit;
   dims receives {workers, tiles, slots}.  Returns 0, or -1 when the library was built without the instrumentation. */
int vt_debug_persist_trace(long long* out, long long* cal, int32_t* dims);
int vt_program_add_imgstats(vt_program* p, const vt_imgstats_desc* d);
int vt_program_add_patchify(vt_program* p, const vt_patchify_desc* d);
int vt_program_add_cls(vt_program* p, const vt_cls_desc* d);
int vt_program_add_pack(vt_program* p, const vt_pack_desc* d);
int vt_program_add_affine(vt_program* p, const vt_affine_desc* d);
int vt_program_add_tembed(vt_program* p, const vt_tembed_desc* d);
int vt_program_add_sde(vt_program* p, const vt_sde_desc* d);
int vt_program_add_lstm(vt_program* p, const vt_lstm_desc* d);
int vt_program_add_qsample(vt_program* p, const vt_qsample_desc* d);
int vt_program_add_siloss(vt_program* p, const vt_siloss_desc* d);
int vt_program_add_tcol(vt_program* p, const vt_tcol_desc* d);
int vt_program_add_gnbwd(vt_program* p, const vt_gnbwd_desc* d);
int vt_program_add_colsum(vt_program* p, const vt_colsum_desc* d);
int vt_program_add_ewise(vt_program* p, const vt_ewise_desc* d);
int vt_program_add_silossbwd(vt_program* p, const vt_silossbwd_desc* d);
int vt_program_add_lstm_train(vt_program* p, const vt_lstm_train_desc* d);
int vt_program_add_lstm_bwd(vt_program* p, const vt_lstm_bwd_desc* d);
int vt_program_add_lngelubwd(vt_program* p, const vt_lngelubwd_desc* d);
int vt_program_add_dropmask(vt_program* p, const vt_dropmask_desc* d);
int vt_program_add_persist(vt_program* p, const vt_persist_desc* d);
int vt_program_add_wgrad(vt_program* p, const vt_wgrad_desc* d);

/* Launch ops [first, first+count) in order on `stream` (count < 0: to the end). */
int vt_program_run(vt_program* p, int first, int count, void* stream);
/* Capture ops [first, first+count) into a CUDA graph (replacing any earlier one); launch it. */
int vt_program_graph_build(vt_program* p, int first, int count);
int vt_program_graph_launch(vt_program* p, void* stream);

/* immediate launch of the fused AdamW + EMA step */
int vt_adamw_ema_step(const vt_adamw_desc* d, void* stream);

/* Operand re-pack after an optimizer step (no reference counterpart: torch re-reads nn.Parameters, here every GEMM reads a packed
   bf16 copy): dst[i] = arena[map[i]] for all records in one launch.  recs_dev: device array of {void* dst; const int32_t* map;
   int64_t n; int32_t bf16; int32_t pad} (n a multiple of 8 per 16-byte aligned dst, or any n with a scalar tail); chunks_dev:
   (record, first element) pairs, 8192 elements per chunk. */
int vt_gather_repack(const void* recs_dev, const int64_t* chunks_dev, int32_t n_chunks, const float* arena_dev, void* stream);

/* pad_and_resize_for_siglip (scripts/utils_eef.py:44-77; called at scripts/franka_inference_eef.py:329-330): n uint8 frames
   [h][w][c] on the device -> [target][target][c], zero-padded to a centred square and down-scaled like
   cv2.resize(..., interpolation=cv2.INTER_AREA) -- bit-identical to OpenCV (integer-factor and fractional-factor paths).
   Up-scaling (max(h, w) < target) returns VT_E_INVALID. */
int vt_pad_resize_area(const uint8_t* src_dev, int32_t n, int32_t h, int32_t w, int32_t c, uint8_t* dst_dev, int32_t target, void* stream);

/* Minibatch assembly from the HBM-resident episode store: replaces ControllerDataset.__getitem__
   (controller_dataset.py:101-170), the DataLoader's collate + host->device copy (controller_dataset.py:467-476,
   bridge_train.py:304-307) and the action normalisation of _prepare_batch_for_diffusion (bridge_train.py:120-145) for B samples
   in one launch.  All episode streams are concatenated over frames in device memory; `start[b]` is the global frame index of
   sample b's first context frame.  Output values are bit-identical to the reference's fp32 tensors. */
typedef struct vt_batch_gather_desc {
  const float* qpos;            /* [frames][A]  fp32(converted_ee_pose_with_gripper), scripts/utils_eef.py:80-90 */
  const float* grip_scaled;     /* [frames]     fp32(gripper / 255), divided in the source precision (controller_dataset.py:124) */
  const float* vla;             /* [frames][vla_T][A] vla_action as stored */
  const float* vla_last_scaled; /* [frames][vla_T]    fp32(vla_action[..., -1] / 255) (controller_dataset.py:130) */
  const float* forces;          /* [frames][Fd] gelsight_force/forces */
  const float* disps;           /* [frames][Dd] gelsight_force/displacement flattened (63*2), or NULL */
  const float* feats;           /* [frames][2 cameras][2 branches][D] cached DINOv2Encoder.forward features (branch 0: images
                                   passed un-normalised, 1: ImageNet-normalised, visual_encoder.py:95-106), or NULL */
  const double* frame_mean;     /* [frames][2] mean pixel value in [0, 1] per frame and camera (required with feats) */
  const int64_t* start;         /* [B] */
  int32_t B, A, vla_T, Fd, Dd, D;
  int32_t context_frames, horizon;
  float* states;                /* out [B][context_frames + horizon][A], may be NULL (like every output below) */
  float* expert_actions;        /* out [B][horizon][A] */
  float* vla_actions;           /* out [B][horizon][A] */
  float* forces_out;            /* out [B][context_frames + horizon][Fd] */
  float* disps_out;             /* out [B][context_frames + horizon][Dd] */
  float* feat_cam1;             /* out [B][D]: features of the last context frame (bridge_train.py:142-145); required with feats */
  float* feat_cam2;
  int32_t* branch;              /* out [2]: branch chosen per camera from the batch mean (visual_encoder.py:100), may be NULL */
  const float* action_mins;     /* [A] fp32 stats; NULL = no normalised outputs */
  const float* action_maxs;
  const float* vla_mins;
  const float* vla_maxs;
  float pad;                    /* padding_factor of normalize_actions (controller_dataset.py:303), 1.4 */
  float* expert_n;              /* out [B][horizon][A] = normalize_actions(expert_actions, stats, 'expert') */
  float* vla_n;                 /* out [B][horizon][A] = normalize_actions(vla_actions, stats, 'vla') */
} vt_batch_gather_desc;
int vt_batch_gather(const vt_batch_gather_desc* d, void* stream);

/* RDT policy.step -> controller.predict hand-off: replaces RoboticDiffusionTransformerModel._unformat_action_to_joint
   + `.to(torch.float32)` (scripts/franka_model_eef.py:199-222,312) and the deployment loop's `vla_tensor[:, :, -1] /= 255` + slice
   (scripts/franka_inference_eef.py:546,552-554).  action_dev: [B][N][S] unified action vectors (VT_BF16 or VT_F32) as the policy
   wrote them; idx_dev [A] positions of the robot's dims in the unified vector; scale_dev [A] (1,...,1,255).
   raw_dev   out [B][N][A] fp32 = what policy.step returns (may be NULL);
   chunk_dev out [B][T_exec][A] fp32 = the tensor predict() receives: last dim / last_div, first T_exec rows (may be NULL). */
int vt_chunk_handoff(const void* action_dev, int32_t dtype, int32_t B, int32_t N, int32_t S, const int32_t* idx_dev, const float* scale_dev,
                     int32_t A, float last_div, float* raw_dev, float* chunk_dev, int32_t T_exec, void* stream);

/* bicubic (A=-0.75, align_corners=False) resize of the patch position embeddings, HF:57-95: src [s*s][D] -> dst [nh*nw][D] */
int vt_pos_embed_resize(const float* src_dev, int32_t s, float* dst_dev, int32_t nh, int32_t nw, int32_t D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VT_B200_H_ */
