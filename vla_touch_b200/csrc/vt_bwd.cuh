// Backward-pass kernels of the interpolant U-Net's Conv1dBlock (conditional_unet_1D.py:40-55, 86-105):
//   tcol_kernel        transposed (im2col-T) operand copies for the weight-gradient GEMMs (K = B*T rows)
//   gn_mish_bwd_kernel GroupNorm(8) + Mish (+ FiLM) backward from the raw conv output, one CTA per (net, sample)
//   colsum_kernel      per-channel parameter gradients = sum over samples of the per-sample partials
// CUDA-core kernels: HBM / L2 bound elementwise + reduction work (the contractions run on gemm_tc_kernel).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "vt_ptx.cuh"
#include "vt_elem.cuh"

namespace vt {

struct TcolArgs {
  const void* src;
  int src_dtype;                 // 0 bf16, 1 f32
  long long ld, sB, sG;
  int G, B, T_src, C;
  int taps;
  int tap_off[8];
  int stride, t_out;
  __nv_bfloat16* out;
  int c_pad;
  long long k_ld, out_g;
};

// out[g][tap * c_pad + c][b * t_out + t] = src[g][b][t * stride + tap_off[tap]][c]  (zero outside [0, T_src)).
// 32 x 32 tiles through shared memory: reads coalesced along channels, writes coalesced along the K index.
// grid = (ceil(B * t_out / 32), ceil(C / 32), G * taps), block = (32, 8).
__global__ void __launch_bounds__(256) tcol_kernel(const TcolArgs a) {
  __shared__ float tile[32][33];
  const int g = blockIdx.z / a.taps, tap = blockIdx.z % a.taps;
  const int k0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int K = a.B * a.t_out;
  const int off = a.tap_off[tap];
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int k = k0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (k < K && c < a.C) {
      const int b = k / a.t_out, t = k % a.t_out;
      const int pos = t * a.stride + off;
      if (pos >= 0 && pos < a.T_src) {
        const long long idx = (long long)g * a.sG + (long long)b * a.sB + (long long)pos * a.ld + c;
        v = a.src_dtype == 0 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.src)[idx])
                             : reinterpret_cast<const float*>(a.src)[idx];
      }
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, k = k0 + threadIdx.x;
    if (c < a.C && k < K)
      a.out[(long long)g * a.out_g + ((long long)tap * a.c_pad + c) * a.k_ld + k] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

// d mish(x) / dx and mish(x) from one exponential: tanh(softplus(x)) = n / (n + 2), n = e (e + 2), e = exp(x)
__device__ __forceinline__ void mish_and_grad(float x, float& m, float& dm) {
  const float e = expf(fminf(x, 20.f));
  const float n = e * (e + 2.f);
  const float tsp = n / (n + 2.f);
  const float sig = e / (1.f + e);
  m = x * tsp;
  dm = tsp + x * (1.f - tsp * tsp) * sig;
}

struct GnBwdArgs {
  const float* raw;      // [G][B][T][C]
  const float* dout;     // [G][B][T][dout_ld]
  long long dout_ld, dout_g;
  const float* gamma;    // [G][p_ld]
  const float* beta;
  int p_ld;
  const float* film;     // [G][B][film_ld] or null: scale at film_off + c
  long long film_g;
  int film_ld, film_off;
  float* dfilm;          // [G][B][film_ld]: d scale at film_off + c, d shift at film_off + C + c
  __nv_bfloat16* draw;   // [G][B][T][C]
  float* part;           // [G][B][3][C]: per-sample (d gamma, d beta, d bias)
  int G, B, T, C, groups;
  float eps;
};

constexpr int GNB_MAX_CPT = 2;   // channels per thread: C <= 512 with 256 threads

// One CTA (256 threads) per (net g, sample b).  Thread owns channels tid, tid + 256; loops over the T positions, so global
// accesses are coalesced along channels.  Four sweeps over the sample's raw tile (<= 128 KB: L1 / L2 resident):
//   1 mean  2 variance (two-pass, like torch)  3 per-channel gradient sums  4 d raw
__global__ void __launch_bounds__(256) gn_mish_bwd_kernel(const GnBwdArgs a) {
  __shared__ float s_ch[2][512];
  __shared__ float s_mean[64], s_rstd[64], s_m1[64], s_m2[64];
  const int g = blockIdx.x / a.B, b = blockIdx.x % a.B;
  const int tid = threadIdx.x;
  const int C = a.C, T = a.T, Cg = C / a.groups;
  const long long sample = ((long long)g * a.B + b) * T;
  const float* raw = a.raw + sample * C;
  const float* dout = a.dout + (long long)g * a.dout_g + (long long)b * T * a.dout_ld;
  const float inv_n = 1.f / (float)(Cg * T);

  // sweep 1: per-channel sums -> group means
  for (int j = 0; j < GNB_MAX_CPT; ++j) {
    const int c = tid + j * 256;
    if (c < C) {
      float s = 0.f;
      for (int t = 0; t < T; ++t) s += raw[(long long)t * C + c];
      s_ch[0][c] = s;
    }
  }
  __syncthreads();
  if (tid < a.groups) {
    float s = 0.f;
    for (int i = 0; i < Cg; ++i) s += s_ch[0][tid * Cg + i];
    s_mean[tid] = s * inv_n;
  }
  __syncthreads();
  // sweep 2: centred second moment -> rstd
  for (int j = 0; j < GNB_MAX_CPT; ++j) {
    const int c = tid + j * 256;
    if (c < C) {
      const float mu = s_mean[c / Cg];
      float s = 0.f;
      for (int t = 0; t < T; ++t) {
        const float d = raw[(long long)t * C + c] - mu;
        s += d * d;
      }
      s_ch[1][c] = s;
    }
  }
  __syncthreads();
  if (tid < a.groups) {
    float s = 0.f;
    for (int i = 0; i < Cg; ++i) s += s_ch[1][tid * Cg + i];
    s_rstd[tid] = rsqrtf(s * inv_n + a.eps);
  }
  __syncthreads();
  // sweep 3: da = dm * mish'(y);  per-channel sums of da, da * xh (and the FiLM gradients)
  float gam[GNB_MAX_CPT], bet[GNB_MAX_CPT], scl[GNB_MAX_CPT], s_da[GNB_MAX_CPT], s_dax[GNB_MAX_CPT];
  for (int j = 0; j < GNB_MAX_CPT; ++j) {
    const int c = tid + j * 256;
    gam[j] = bet[j] = 0.f;
    scl[j] = 1.f;
    s_da[j] = s_dax[j] = 0.f;
    if (c < C) {
      gam[j] = a.gamma[(long long)g * a.p_ld + c];
      bet[j] = a.beta[(long long)g * a.p_ld + c];
      const float mu = s_mean[c / Cg], rs = s_rstd[c / Cg];
      if (a.film) scl[j] = a.film[(long long)g * a.film_g + (long long)b * a.film_ld + a.film_off + c];
      float f_sc = 0.f, f_sh = 0.f;
      for (int t = 0; t < T; ++t) {
        const float xh = (raw[(long long)t * C + c] - mu) * rs;
        const float go = dout[(long long)t * a.dout_ld + c];
        float m, dm;
        mish_and_grad(xh * gam[j] + bet[j], m, dm);
        const float da = go * scl[j] * dm;
        s_da[j] += da;
        s_dax[j] += da * xh;
        f_sc += go * m;
        f_sh += go;
      }
      if (a.dfilm) {
        float* df = a.dfilm + (long long)g * a.film_g + (long long)b * a.film_ld + a.film_off;
        df[c] = f_sc;
        df[C + c] = f_sh;
      }
      s_ch[0][c] = gam[j] * s_da[j];
      s_ch[1][c] = gam[j] * s_dax[j];
    }
  }
  __syncthreads();
  if (tid < a.groups) {
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < Cg; ++i) {
      s1 += s_ch[0][tid * Cg + i];
      s2 += s_ch[1][tid * Cg + i];
    }
    s_m1[tid] = s1 * inv_n;
    s_m2[tid] = s2 * inv_n;
  }
  __syncthreads();
  // sweep 4: d raw = rstd * (da * gamma - mean_g(dxh) - xh * mean_g(dxh * xh));  d bias = sum_t d raw
  for (int j = 0; j < GNB_MAX_CPT; ++j) {
    const int c = tid + j * 256;
    if (c < C) {
      const int gi = c / Cg;
      const float mu = s_mean[gi], rs = s_rstd[gi], m1 = s_m1[gi], m2 = s_m2[gi];
      float s_dr = 0.f;
      for (int t = 0; t < T; ++t) {
        const float xh = (raw[(long long)t * C + c] - mu) * rs;
        const float go = dout[(long long)t * a.dout_ld + c];
        float m, dm;
        mish_and_grad(xh * gam[j] + bet[j], m, dm);
        const float dr = rs * (go * scl[j] * dm * gam[j] - m1 - xh * m2);
        s_dr += dr;
        a.draw[(sample + t) * C + c] = __float2bfloat16(dr);
      }
      float* p = a.part + ((long long)g * a.B + b) * 3 * C;
      p[c] = s_dax[j];
      p[C + c] = s_da[j];
      p[2 * C + c] = s_dr;
    }
  }
}

// Shared-memory-resident version (C in {128, 256, 512}, T * C * 8 bytes <= 128 KB: every shape of the U-Net at T <= 64).
// gn_mish_bwd_kernel re-reads the sample's tile from L1 / L2 in four sweeps with one dependent load per thread and position and
// evaluates Mish twice with expf: 0.85 TB/s of its 10 bytes per element, 28 % of the training program.  Here the CTA (512 threads)
// loads the raw and d out tiles ONCE with all its 128-bit loads in flight, keeps them in shared memory, overwrites them in place
// with x_hat and d a in sweep 3, so Mish (ex2 + rcp on MUFU) is evaluated once per element, and splits the T positions of a
// channel over 512 / C threads.  Same outputs as gn_mish_bwd_kernel (different summation order over T).
constexpr int GNBS_THREADS = 512;

// d mish / dx and mish on two lanes: packed f32x2 arithmetic, one ex2 and ONE rcp per lane (1 / ((1 + e)(n + 2)) gives both
// 1 / (n + 2) and 1 / (1 + e) by a multiplication)
__device__ __forceinline__ void mish_and_grad2(float2 x, float2& m, float2& dm) {
  const float2 one = make_float2(1.f, 1.f), two = make_float2(2.f, 2.f);
  const float2 xe = fmul2(make_float2(fminf(x.x, 20.f), fminf(x.y, 20.f)), make_float2(1.4426950408889634f, 1.4426950408889634f));
  const float2 e = make_float2(ex2_approx(xe.x), ex2_approx(xe.y));
  const float2 e1 = fadd2(e, one);                 // 1 + e
  const float2 n = fmul2(e, fadd2(e, two));        // e (e + 2)
  const float2 n2 = fadd2(n, two);                 // n + 2
  const float2 den = fmul2(e1, n2);
  float2 r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(den.y));
  const float2 tsp = fmul2(n, fmul2(r, e1));       // n / (n + 2)
  const float2 sig = fmul2(e, fmul2(r, n2));       // e / (1 + e)
  m = fmul2(x, tsp);
  // dm = tsp + x (1 - tsp^2) sig
  const float2 omt = ffma2(make_float2(-tsp.x, -tsp.y), tsp, one);
  dm = ffma2(fmul2(x, omt), sig, tsp);
}

// Thread = one channel PAIR (packed f32x2 arithmetic, 8-byte shared-memory accesses) and every PARTS-th position; C is a template
// parameter so that the tile addressing is strength-reduced (the first version spent 128 instructions per element, most of them
// integer address arithmetic and loop control: 57 % issue utilisation at 0.8 TB/s).
// MINB = 2: two CTAs per SM (<= 64 registers per thread) for the shapes whose tiles take <= 100 KB of shared memory, so that one CTA's
// loads overlap the other's arithmetic; the 128 KB shapes stay at one CTA per SM.
// THREADS = 1024 (64 registers): the 128 KB shapes, which cannot share an SM, run with twice the warps per sample instead.
template <int C, int MINB = 1, int THREADS = GNBS_THREADS>
__global__ void __launch_bounds__(THREADS, MINB) gn_mish_bwd_smem_kernel(const GnBwdArgs a) {
  extern __shared__ float gnb_smem[];
  constexpr int CP = C / 2;                     // channel pairs
  constexpr int PARTS = THREADS / CP;           // threads per channel pair at 512 threads: 2 (C = 512), 4 (256), 8 (128)
  const int T = a.T, Cg = C / a.groups;
  float* s_raw = gnb_smem;              // [T][C]  raw conv output, later x_hat
  float* s_do = s_raw + T * C;          // [T][C]  d out, later d a = d out * scale * mish'
  float* s_red = s_do + T * C;          // [4][PARTS][C] partial sums of the threads that share a channel pair
  float* s_ch = s_red + 4 * PARTS * C;  // [2][C] per-channel totals
  float* s_grp = s_ch + 2 * C;          // [4][64] mean, E[x^2] / rstd, m1, m2 per group
  const int g = blockIdx.x / a.B, b = blockIdx.x % a.B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long sample = ((long long)g * a.B + b) * T;
  const float* raw = a.raw + sample * C;
  const float* dout = a.dout + (long long)g * a.dout_g + (long long)b * T * a.dout_ld;
  const float inv_n = 1.f / (float)(Cg * T);
  const int c = 2 * (tid % CP), part = tid / CP;
  const int rows = T / PARTS;                   // positions of this thread: part, part + PARTS, ...

  // ---- the two tiles -> shared memory (one tile at a time: 16 x 128-bit loads in flight per thread) ----
  {
    const int n4 = T * C / 4;
    constexpr int c4 = C / 4;
    constexpr int MAXV = MINB == 2 ? 8 : 64 * 512 / 4 / THREADS;   // MINB = 2: tiles of <= 12 800 elements (<= 100 KB for both)
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const float* src = which ? dout : raw;
      const long long ld = which ? a.dout_ld : (long long)C;
      float4* dst = reinterpret_cast<float4*>(which ? s_do : s_raw);
      float4 v[MAXV];
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        const int i = tid + k * THREADS;
        if (i < n4) {
          const int t = i / c4, cc = (i - t * c4) * 4;
          v[k] = *reinterpret_cast<const float4*>(src + (long long)t * ld + cc);
        }
      }
#pragma unroll
      for (int k = 0; k < MAXV; ++k) {
        const int i = tid + k * THREADS;
        if (i < n4) dst[i] = v[k];
      }
    }
  }
  __syncthreads();
  float2* raw2 = reinterpret_cast<float2*>(s_raw + part * C + c);   // this thread's column pair, first position; stride PARTS * C
  float2* do2 = reinterpret_cast<float2*>(s_do + part * C + c);
  constexpr int ST2 = PARTS * C / 2;                                // ... in float2 units
  auto put = [&](int slot, float2 v) { *reinterpret_cast<float2*>(s_red + (slot * PARTS + part) * C + c) = v; };
  // totals over the channel's threads -> s_ch[slot][.] (n_slots <= 2 at a time), then per-group totals by one warp per group
  auto channel_totals = [&](int n_slots, int first_slot) {
    __syncthreads();
    if (tid < C) {
      for (int sl = 0; sl < n_slots; ++sl) {
        float v = 0.f;
#pragma unroll
        for (int p = 0; p < PARTS; ++p) v += s_red[((first_slot + sl) * PARTS + p) * C + tid];
        s_ch[sl * C + tid] = v;
      }
    }
    __syncthreads();
  };
  auto group_totals = [&](float* out0, float* out1, float scale) {
    if (warp < a.groups) {
      float v0 = 0.f, v1 = 0.f;
      for (int i = lane; i < Cg; i += 32) {
        v0 += s_ch[warp * Cg + i];
        v1 += s_ch[C + warp * Cg + i];
      }
      v0 = warp_sum(v0);
      v1 = warp_sum(v1);
      if (lane == 0) {
        out0[warp] = v0 * scale;
        out1[warp] = v1 * scale;
      }
    }
    __syncthreads();
  };
  // sweeps 1 + 2 in one pass: group mean and variance from sum / sum of squares in fp32, exactly what the forward GroupNorm
  // epilogue normalises with (vt_gemm.cuh epilogue_gn_fast), so x_hat here equals the forward's
  {
    float2 s = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int i = 0; i < rows; ++i) {
      const float2 x = raw2[i * ST2];
      s = fadd2(s, x);
      q = ffma2(x, x, q);
    }
    put(0, s);
    put(1, q);
    channel_totals(2, 0);
    group_totals(s_grp, s_grp + 64, inv_n);
  }
  const int gi = c / Cg;                         // both channels of the pair are in the same group (Cg is even)
  const float mu = s_grp[gi];
  const float rs = rsqrtf(fmaxf(s_grp[64 + gi] - mu * mu, 0.f) + a.eps);
  const float2 mu2 = make_float2(mu, mu), rs2 = make_float2(rs, rs);
  const float2 gam = *reinterpret_cast<const float2*>(a.gamma + (long long)g * a.p_ld + c);
  const float2 bet = *reinterpret_cast<const float2*>(a.beta + (long long)g * a.p_ld + c);
  const float2 scl = a.film ? *reinterpret_cast<const float2*>(a.film + (long long)g * a.film_g + (long long)b * a.film_ld + a.film_off + c)
                            : make_float2(1.f, 1.f);
  // sweep 3: x_hat and d a in place; per-channel sums of d a, d a * x_hat and the FiLM gradients
  {
    float2 s_da = make_float2(0.f, 0.f), s_dax = s_da, f_sc = s_da, f_sh = s_da;
    const float2 nmr = make_float2(-mu * rs, -mu * rs);
#pragma unroll 2
    for (int i = 0; i < rows; ++i) {
      const float2 xh = ffma2(raw2[i * ST2], rs2, nmr);
      const float2 go = do2[i * ST2];
      float2 m, dm;
      mish_and_grad2(ffma2(xh, gam, bet), m, dm);
      const float2 da = fmul2(fmul2(go, scl), dm);
      raw2[i * ST2] = xh;
      do2[i * ST2] = da;
      s_da = fadd2(s_da, da);
      s_dax = ffma2(da, xh, s_dax);
      f_sc = ffma2(go, m, f_sc);
      f_sh = fadd2(f_sh, go);
    }
    put(0, s_da);
    put(1, s_dax);
    put(2, f_sc);
    put(3, f_sh);
    __syncthreads();
    if (tid < C) {
      float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
      for (int p = 0; p < PARTS; ++p) {
        t0 += s_red[(0 * PARTS + p) * C + tid];
        t1 += s_red[(1 * PARTS + p) * C + tid];
        t2 += s_red[(2 * PARTS + p) * C + tid];
        t3 += s_red[(3 * PARTS + p) * C + tid];
      }
      if (a.dfilm) {
        float* df = a.dfilm + (long long)g * a.film_g + (long long)b * a.film_ld + a.film_off;
        df[tid] = t2;
        df[C + tid] = t3;
      }
      float* p = a.part + ((long long)g * a.B + b) * 3 * C;
      p[tid] = t1;           // d gamma contribution: sum d a * x_hat
      p[C + tid] = t0;       // d beta contribution: sum d a
      const float gm = a.gamma[(long long)g * a.p_ld + tid];
      s_ch[tid] = gm * t0;
      s_ch[C + tid] = gm * t1;
    }
    __syncthreads();
    group_totals(s_grp + 128, s_grp + 192, inv_n);
  }
  // sweep 4: d raw = rstd * (d a * gamma - mean_g(d xh) - x_hat * mean_g(d xh * x_hat));  d bias = sum_t d raw
  {
    const float m1 = s_grp[128 + gi], m2 = s_grp[192 + gi];
    const float2 nm1 = make_float2(-m1, -m1), nm2 = make_float2(-m2, -m2);
    float2 s_dr = make_float2(0.f, 0.f);
    __nv_bfloat16* dr_out = a.draw + (sample + part) * C + c;
#pragma unroll 4
    for (int i = 0; i < rows; ++i) {
      const float2 t0 = ffma2(do2[i * ST2], gam, nm1);
      const float2 dr = fmul2(rs2, ffma2(raw2[i * ST2], nm2, t0));
      s_dr = fadd2(s_dr, dr);
      *reinterpret_cast<uint32_t*>(dr_out + (long long)i * PARTS * C) = pack_bf16x2(dr.x, dr.y);
    }
    put(0, s_dr);          // slot 0 was last read before the group_totals barrier above
    __syncthreads();
    if (tid < C) {
      float v = 0.f;
#pragma unroll
      for (int p = 0; p < PARTS; ++p) v += s_red[p * C + tid];
      a.part[((long long)g * a.B + b) * 3 * C + 2 * C + tid] = v;
    }
  }
}
__host__ __device__ constexpr size_t gnbs_smem_bytes(int T, int C, int threads = GNBS_THREADS) {
  return (size_t)(2 * T * C + 4 * (threads / (C / 2)) * C + 2 * C + 256) * sizeof(float);
}

// out_k[g][c] = sum_b part[g][b][k][c], k = 0..2 -> (d gamma, d beta, d bias), each [G][p_ld]
// One CTA per 32 columns of a (g, k) row block: 32 column lanes x 8 sample lanes, four independent loads in flight per thread, the
// eight partials combined through shared memory in a fixed order (deterministic).  (The first version gave every thread one column
// and the whole serial chain of B dependent L2 loads, on a grid of G * 3 * C / 256 = 18 CTAs: ~75 us per call, more than the
// GroupNorm backward kernel it follows.)
__global__ void __launch_bounds__(256) gn_colsum_kernel(const float* __restrict__ part, int G, int B, int C, float* dgamma,
                                                        float* dbeta, float* dbias, int p_ld) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int strips = (C + 31) / 32;
  const int strip = blockIdx.x % strips, k = (blockIdx.x / strips) % 3, g = blockIdx.x / (3 * strips);
  const int c = strip * 32 + tx;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < C) {
    const float* p = part + (long long)g * B * 3 * C + (long long)k * C + c;
    const long long st = 3LL * C;
    int b = ty;
    for (; b + 24 < B; b += 32) {
      s0 += p[(long long)b * st];
      s1 += p[(long long)(b + 8) * st];
      s2 += p[(long long)(b + 16) * st];
      s3 += p[(long long)(b + 24) * st];
    }
    for (; b < B; b += 8) s0 += p[(long long)b * st];
  }
  red[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][tx];
    float* o = k == 0 ? dgamma : (k == 1 ? dbeta : dbias);
    if (o) o[(long long)g * p_ld + c] = s;
  }
}

// out[g][c] = sum_r x[g][r][c] : bias gradient of a convolution without GroupNorm (conv1d_bwd's `dy.sum(dim=(0, 2))`).
// HBM-bound column reduction: a CLUSTER of COLSUM_SPLIT CTAs shares one (32-column, group) strip, each CTA sums its slice of
// the rows (32 row lanes x 8 float4 column lanes, 4 independent loads in flight per thread), the partials are combined
// through distributed shared memory in a fixed order (deterministic, no atomics, no scratch buffer).
// grid = (ceil(C / 32), G, COLSUM_SPLIT), cluster (1, 1, COLSUM_SPLIT), block = 256.
constexpr int COLSUM_SPLIT = 8;
__global__ void __cluster_dims__(1, 1, COLSUM_SPLIT) __launch_bounds__(256)
    colsum_kernel(const float* __restrict__ x, long long ld, long long x_g, int rows, int C, float* __restrict__ out, int out_ld) {
  __shared__ float red[32][33];
  __shared__ float part[32];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;       // 8 column lanes (4 columns each) x 32 row lanes
  const int c0 = blockIdx.x * 32 + tx * 4, g = blockIdx.y;
  const unsigned rank = cluster_ctarank();
  const int per = (rows + COLSUM_SPLIT - 1) / COLSUM_SPLIT;
  const int r0 = (int)rank * per, r1 = min(rows, r0 + per);
  const float* xg = x + (long long)g * x_g;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(xg) & 15) == 0) && (c0 + 4 <= C);
  if (vec) {
    int r = r0 + ty;
    for (; r + 96 < r1; r += 128) {
      const float4 a = *reinterpret_cast<const float4*>(xg + (long long)r * ld + c0);
      const float4 b = *reinterpret_cast<const float4*>(xg + (long long)(r + 32) * ld + c0);
      const float4 c = *reinterpret_cast<const float4*>(xg + (long long)(r + 64) * ld + c0);
      const float4 d = *reinterpret_cast<const float4*>(xg + (long long)(r + 96) * ld + c0);
      s.x += (a.x + b.x) + (c.x + d.x);
      s.y += (a.y + b.y) + (c.y + d.y);
      s.z += (a.z + b.z) + (c.z + d.z);
      s.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; r < r1; r += 32) {
      const float4 a = *reinterpret_cast<const float4*>(xg + (long long)r * ld + c0);
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
  } else {
    for (int r = r0 + ty; r < r1; r += 32) {
      const float* row = xg + (long long)r * ld;
      if (c0 < C) s.x += row[c0];
      if (c0 + 1 < C) s.y += row[c0 + 1];
      if (c0 + 2 < C) s.z += row[c0 + 2];
      if (c0 + 3 < C) s.w += row[c0 + 3];
    }
  }
  red[ty][tx * 4] = s.x;
  red[ty][tx * 4 + 1] = s.y;
  red[ty][tx * 4 + 2] = s.z;
  red[ty][tx * 4 + 3] = s.w;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][threadIdx.x];
    part[threadIdx.x] = t;
  }
  cluster_sync_all();
  if (rank == 0 && threadIdx.x < 32) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    float t = 0.f;
#pragma unroll
    for (unsigned k = 0; k < COLSUM_SPLIT; ++k) t += ld_shared_cluster_f32(mapa_shared(smem_u32(&part[threadIdx.x]), k));
    if (c < C) out[(long long)g * out_ld + c] = t;
  }
  cluster_sync_all();   // the partials must stay readable until rank 0 is done
}

// The same reduction for FEW rows and many columns (the split-K partial sums of a weight gradient: rows = slices, C = the whole
// gradient): one thread per 4 columns, rows summed in order (deterministic), 128-bit accesses.
__global__ void __launch_bounds__(256) colsum_fewrows_kernel(const float* __restrict__ x, long long ld, long long x_g, int rows, long long C,
                                                             float* __restrict__ out, long long out_ld) {
  const int g = blockIdx.y;
  const float* xg = x + (long long)g * x_g;
  float* og = out + (long long)g * out_ld;
  const bool vec = ((ld & 3) == 0) && ((C & 3) == 0) && ((reinterpret_cast<uintptr_t>(xg) & 15) == 0) && ((reinterpret_cast<uintptr_t>(og) & 15) == 0);
  if (vec) {
    const long long n4 = C >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      float4 s = *reinterpret_cast<const float4*>(xg + 4 * i);
      for (int r = 1; r < rows; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(xg + (long long)r * ld + 4 * i);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      *reinterpret_cast<float4*>(og + 4 * i) = s;
    }
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < C; i += (long long)gridDim.x * blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < rows; ++r) s += xg[(long long)r * ld + i];
      og[i] = s;
    }
  }
}

__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * expf(-0.5f * x * x) * 0.3989422804014327f;
}

// Strided fp32 elementwise combine over [rows][cols] windows:  op 0: out = a + b (gradient fan-in at a skip connection),
// op 1: out = a * mish'(b) (backward of the Mish in front of the FiLM / time-embedding linears), op 2: out = a * gelu'(b),
// op 3: out = a * b, op 4: out = alpha * (a - b) (derivative of the MSE loss).
__global__ void __launch_bounds__(256) ewise_kernel(const float* __restrict__ a, long long a_ld, const float* __restrict__ b,
                                                    long long b_ld, float* __restrict__ out, long long out_ld, long long rows,
                                                    int cols, int op, float alpha) {
  const long long total = rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cols);
    const long long r = idx / cols;
    const float x = a[r * a_ld + c], y = b[r * b_ld + c];
    float v;
    if (op == 0) {
      v = x + y;
    } else if (op == 1) {
      float m, dm;
      mish_and_grad(y, m, dm);
      v = x * dm;
    } else if (op == 2) {
      v = x * gelu_erf_grad(y);
    } else if (op == 3) {
      v = x * y;
    } else {
      v = alpha * (x - y);
    }
    out[r * out_ld + c] = v;
  }
}

// Backward of LayerNorm(256) -> GELU (the output head of the LSTM controller, lstm_step_controller.py:76-82) from the saved
// LayerNorm input: one warp per row, 8 columns per lane.
//   zh = (z0 - mean) rstd;  z1 = zh gamma + beta;  d1 = dzn gelu'(z1);  dzh = d1 gamma
//   dz0 = rstd (dzh - mean(dzh) - zh mean(dzh zh));  also stores d1 and d1 zh (their column sums are d beta and d gamma)
__global__ void __launch_bounds__(256) ln_gelu_bwd_kernel(const float* __restrict__ z0, const float* __restrict__ dzn,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                          float* __restrict__ dz0, float* __restrict__ d1_out,
                                                          float* __restrict__ d1zh_out, int rows) {
  constexpr int D = 256, PER = D / 32;
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float x[PER], s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    x[i] = z0[row * D + lane + 32 * i];
    s += x[i];
  }
  const float mean = warp_sum(s) * (1.f / D);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) v += (x[i] - mean) * (x[i] - mean);
  const float rstd = rsqrtf(warp_sum(v) * (1.f / D) + eps);
  float zh[PER], dzh[PER], s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    zh[i] = (x[i] - mean) * rstd;
    const float g = gamma[c];
    const float d1 = dzn[row * D + c] * gelu_erf_grad(zh[i] * g + beta[c]);
    d1_out[row * D + c] = d1;
    d1zh_out[row * D + c] = d1 * zh[i];
    dzh[i] = d1 * g;
    s1 += dzh[i];
    s2 += dzh[i] * zh[i];
  }
  s1 = warp_sum(s1) * (1.f / D);
  s2 = warp_sum(s2) * (1.f / D);
#pragma unroll
  for (int i = 0; i < PER; ++i) dz0[row * D + lane + 32 * i] = rstd * (dzh[i] - s1 - zh[i] * s2);
}

// Inverted-dropout mask (nn.Dropout / the inter-layer dropout of nn.LSTM in training mode, lstm_step_controller.py:66-82):
// mask[i] = u_i >= p ? 1 / (1 - p) : 0 with u_i ~ U[0, 1) from Philox4x32-10 (seed, element, stream) or, for parity tests,
// from an injected buffer.  The mask is stored: forward and backward both multiply by it (ewise MUL).
__global__ void __launch_bounds__(256) dropmask_kernel(const float* __restrict__ inject, float p, unsigned long long seed,
                                                       const unsigned long long* __restrict__ seed_dev, int stream,
                                                       float* __restrict__ mask, long long n) {
  if (seed_dev) seed += *seed_dev;
  const float keep = 1.f / (1.f - p);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    float u;
    if (inject) {
      u = inject[idx];
    } else {
      curandStatePhilox4_32_10_t st;
      curand_init(seed, (unsigned long long)idx, (unsigned long long)stream, &st);
      u = 1.f - curand_uniform(&st);          // curand_uniform is (0, 1]
    }
    mask[idx] = u >= p ? keep : 0.f;
  }
}

// d (L_v + L_s + L_b) / d (net outputs), bridge_model.py:183-218 with the batch mean of get_loss (:240-246), nets stacked as
// [b_net, v_net, s_net]:  d/db = (b - (x1 - x0 + gdot z)) / B,  d/dv = (v - (x1 - x0)) / B,  d/ds = (s + z) / B,  z = d z_unit
__global__ void __launch_bounds__(256) siloss_bwd_kernel(const float* __restrict__ bvs, const float* __restrict__ x0,
                                                         const float* __restrict__ x1, const float* __restrict__ z_unit,
                                                         const float* __restrict__ tclip, float d, int B, int n,
                                                         float* __restrict__ dvs) {
  const long long per = (long long)B * n, total = 3 * per;
  const float inv_b = 1.f / (float)B;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx / per);
    const long long i = idx - (long long)g * per;
    const int b = (int)(i / n);
    const float pt = x1[i] - x0[i];
    const float z = d * z_unit[i];
    const float gd = 1.4142f * (1.0f - 2.0f * tclip[b]);
    const float tgt = g == 0 ? pt + gd * z : (g == 1 ? pt : -z);
    dvs[idx] = (bvs[idx] - tgt) * inv_b;
  }
}

}  // namespace vt
