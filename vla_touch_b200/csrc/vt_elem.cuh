// HBM-bound glue kernels of the refinement path: image statistics + patchify (visual_encoder.py:66-106),
// CLS row, LayerNorm (HF:354,359,449), row packing / Mish(gf), normalise / de-normalise
// (controller_dataset.py:303-384), sinusoidal step embedding (conditional_unet_1D.py:7-19) and the
// Euler-Maruyama update of sde_vs (bridge_model.py:352-385).  128-bit accesses and warp-shuffle reductions.
#pragma once
#include <curand_kernel.h>

#include "vt_ptx.cuh"

namespace vt {

// Store one value as bf16, as plain f32, or as a tf32 (hi, lo) pair `plane` elements apart.
__device__ __forceinline__ void store_val(void* out, int dtype, long long idx, long long plane, float v) {
  if (dtype == 0) {
    reinterpret_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16(v);
  } else if (plane > 0) {
    const float hi = tf32_hi(v);
    reinterpret_cast<float*>(out)[idx] = hi;
    reinterpret_cast<float*>(out)[idx + plane] = v - hi;
  } else {
    reinterpret_cast<float*>(out)[idx] = v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------
// image statistics: max and mean of the WHOLE call tensor (batch-global predicates)
// ------------------------------------------------------------------------------------------
constexpr int STATS_MAX_BLOCKS = 1024;

template <typename T>
__global__ void __launch_bounds__(256) imgstats_partial_kernel(const T* __restrict__ img, long long count,
                                                               float* __restrict__ pmax, double* __restrict__ psum) {
  constexpr int V = 16 / sizeof(T);
  const long long nvec = count / V;
  const uint4* p4 = reinterpret_cast<const uint4*>(img);
  float mx = -INFINITY;
  double sum = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 q = __ldg(p4 + i);
    if constexpr (sizeof(T) == 1) {
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
      uint32_t s = 0, m = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t v = (w[k] >> (8 * b)) & 0xFFu;
          s += v;
          m = max(m, v);
        }
      }
      sum += (double)s;
      mx = fmaxf(mx, (float)m);
    } else {
      const float f[4] = {__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w)};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        sum += (double)f[k];
        mx = fmaxf(mx, f[k]);
      }
    }
  }
  if (blockIdx.x == 0) {  // ragged tail
    for (long long i = nvec * V + threadIdx.x; i < count; i += blockDim.x) {
      const float v = (float)img[i];
      sum += (double)v;
      mx = fmaxf(mx, v);
    }
  }
  __shared__ float smx[8];
  __shared__ double ssum[8];
  mx = warp_max(mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    smx[w] = mx;
    ssum[w] = sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      mx = fmaxf(mx, smx[k]);
      sum += ssum[k];
    }
    pmax[blockIdx.x] = mx;
    psum[blockIdx.x] = sum;
  }
}

__global__ void __launch_bounds__(256) imgstats_final_kernel(const float* __restrict__ pmax,
                                                             const double* __restrict__ psum, int nblk, long long count,
                                                             int* __restrict__ flags) {
  float mx = -INFINITY;
  double sum = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
    mx = fmaxf(mx, pmax[i]);
    sum += psum[i];
  }
  __shared__ float smx[8];
  __shared__ double ssum[8];
  mx = warp_max(mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    smx[w] = mx;
    ssum[w] = sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      mx = fmaxf(mx, smx[k]);
      sum += ssum[k];
    }
    const bool scale255 = mx > 1.0f;                  // visual_encoder.py:78
    double mean = sum / (double)count;
    if (scale255) mean /= 255.0;
    flags[0] = scale255 ? 1 : 0;
    flags[1] = ((float)mean < 0.5f) ? 0 : 1;          // visual_encoder.py:100 (skip normalisation when mean < 0.5)
  }
}

// ------------------------------------------------------------------------------------------
// patchify: layout fix + /255 + ImageNet normalise + im2col for the k14 s14 patch projection
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) patchify_kernel(const T* __restrict__ img, int layout, int images, int H, int W,
                                                       int patch, const int* __restrict__ flags, void* __restrict__ out,
                                                       int out_dtype, int out_cols, int out_ld, long long out_plane) {
  // one thread = 8 consecutive im2col columns of one patch row (one 16-byte store for bf16 outputs)
  const int gw = W / patch, gh = H / patch;
  const int pp = patch * patch;
  const int kk = 3 * pp;
  const int k8n = out_cols >> 3;
  const long long total = (long long)images * gh * gw * k8n;
  const bool scale255 = flags[0] != 0, donorm = flags[1] != 0;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  // uint8 pixels: the normalised value of every (channel, pixel value) is computed ONCE per block with exactly the
  // reference's operations (x / 255, (x - mean) / std in IEEE fp32, visual_encoder.py:78,95-106) and looked up afterwards.
  __shared__ float lut[sizeof(T) == 1 ? 3 * 256 : 1];
  if constexpr (sizeof(T) == 1) {
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
      const int c = i >> 8;
      float x = (float)(i & 255);
      if (scale255) x = __fdiv_rn(x, 255.0f);
      if (donorm) x = __fdiv_rn(__fsub_rn(x, mean[c]), stdv[c]);
      lut[i] = x;
    }
    __syncthreads();
  }
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k0 = (int)(idx % k8n) << 3;
    const long long m = idx / k8n;
    const int p = (int)(m % (gh * gw));
    const long long b = m / (gh * gw);
    const int py = p / gw, px = p - py * gw;
    // (channel, row, column) of column k0 inside the patch, advanced incrementally: two divisions per 8 elements
    int c = k0 / pp;
    int r = k0 - c * pp;
    int i = r / patch, j = r - i * patch;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float x = 0.f;
      if (k0 + e < kk) {
        const int yy = py * patch + i, xx = px * patch + j;
        const long long src = (layout == 0) ? (((b * H + yy) * W + xx) * 3 + c) : (((b * 3 + c) * H + yy) * (long long)W + xx);
        if constexpr (sizeof(T) == 1) {
          x = lut[(c << 8) | (int)img[src]];
        } else {
          x = (float)img[src];
          if (scale255) x = __fdiv_rn(x, 255.0f);
          if (donorm) x = __fdiv_rn(__fsub_rn(x, mean[c]), stdv[c]);
        }
      }
      v[e] = x;
      if (++j == patch) {
        j = 0;
        if (++i == patch) {
          i = 0;
          ++c;
        }
      }
    }
    const long long o = m * out_ld + k0;
    if (out_dtype == 0) {
      uint4 w;
      w.x = pack_bf16x2(v[0], v[1]);
      w.y = pack_bf16x2(v[2], v[3]);
      w.z = pack_bf16x2(v[4], v[5]);
      w.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + o) = w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) store_val(out, out_dtype, o + e, out_plane, v[e]);
    }
  }
}

__global__ void cls_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ h,
                           int images, int tokens, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= images * D) return;
  const int b = i / D, d = i - b * D;
  h[(long long)b * tokens * D + d] = cls[d] + pos[d];
}

// ------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, two-pass statistics held in registers
// ------------------------------------------------------------------------------------------
template <int VEC>  // D = VEC * 128
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long in_ld,
                                                        long long in_row_stride, int rows, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, void* __restrict__ out,
                                                        int out_dtype, long long out_ld, long long out_plane, int act) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int D = VEC * 128;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * in_row_stride * in_ld);
  float4 v[VEC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 g = g4[lane + 32 * i], b = b4[lane + 32 * i];
    float y[4] = {(v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                  (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w};
    if (act == 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) y[k] = gelu_erf(y[k]);
    }
    const long long o = (long long)row * out_ld + 4 * (lane + 32 * i);
    if (out_dtype == 0) {
      uint2 pk;
      pk.x = pack_bf16x2(y[0], y[1]);
      pk.y = pack_bf16x2(y[2], y[3]);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + o) = pk;
    } else if (out_plane > 0) {
      float hi[4], lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        hi[k] = tf32_hi(y[k]);
        lo[k] = y[k] - hi[k];
      }
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o + out_plane) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = make_float4(y[0], y[1], y[2], y[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// pack: out[r][dst_c0 + c] = cast(act(src[r][c])), optional zero tail
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, long long src_ld, int rows, int cols,
                                                   int act, void* __restrict__ out, int out_dtype, long long out_ld,
                                                   int dst_c0, long long out_plane, int zero_to, int src_row_div) {
  const int width = zero_to > cols ? zero_to : cols;
  const long long total = (long long)rows * width;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % width);
    const long long r = idx / width;
    float v = 0.f;
    if (c < cols) {
      v = src[(src_row_div > 1 ? r / src_row_div : r) * src_ld + c];
      if (act == 1) v = gelu_erf(v);
      else if (act == 2) v = mish_precise(v);
    }
    store_val(out, out_dtype, r * out_ld + dst_c0 + c, out_plane, v);
  }
}

// fp32 -> bf16 copy of whole 8-column groups (the gradient casts of the training program: 12.6 M elements per launch): two 128-bit
// loads and one 128-bit store per thread instead of eight scalar accesses with a 64-bit division each
__global__ void __launch_bounds__(256) pack_cast8_kernel(const float* __restrict__ src, long long src_ld, long long rows, int cols8,
                                                         __nv_bfloat16* __restrict__ out, long long out_ld) {
  const long long total = rows * cols8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols8;
    const int c = (int)(idx - r * cols8) * 8;
    const float4 a = *reinterpret_cast<const float4*>(src + r * src_ld + c);
    const float4 b = *reinterpret_cast<const float4*>(src + r * src_ld + c + 4);
    uint4 w;
    w.x = pack_bf16x2(a.x, a.y);
    w.y = pack_bf16x2(a.z, a.w);
    w.z = pack_bf16x2(b.x, b.y);
    w.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(out + r * out_ld + c) = w;
  }
}

// ------------------------------------------------------------------------------------------
// normalize_actions / denormalize_actions (padding factor 1.4), same operation order as the reference, no FMA
// contraction, so the fp32 result is bit-identical to the PyTorch CPU path.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) affine_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                     const float* __restrict__ mins, const float* __restrict__ maxs,
                                                     int rows, int A, int denorm, float pad, void* __restrict__ xpad, int xpad_dtype,
                                                     int xpad_ld, long long xpad_plane, const float* __restrict__ add) {
  const long long total = (long long)rows * A;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(idx % A);
    const long long r = idx / A;
    const float lo = mins[a], hi = maxs[a];
    const float orig = __fsub_rn(hi, lo);
    const float padded = __fmul_rn(orig, pad);
    const float center = __fdiv_rn(__fadd_rn(lo, hi), 2.0f);
    const float half = __fdiv_rn(padded, 2.0f);
    const float pmin = __fsub_rn(center, half);
    const float pmax = __fadd_rn(center, half);
    float range = __fsub_rn(pmax, pmin);
    float v = x[idx];
    if (add) v = __fadd_rn(v, add[idx]);
    float y;
    if (denorm == 2) {
      y = v;
    } else if (!denorm) {
      if (range < 1e-6f) range = 1.0f;
      y = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(v, pmin)), range), 1.0f);
    } else {
      y = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(v, 1.0f), 2.0f), range), pmin);
    }
    if (out) out[idx] = y;
    if (xpad) store_val(xpad, xpad_dtype, r * xpad_ld + a, xpad_plane, y);
  }
}

// SinusoidalPosEmb: out[r] = cat(sin(t f_i), cos(t f_i)),  f_i = exp(i * -(ln 1e4 / (half - 1)))
__global__ void __launch_bounds__(256) tembed_kernel(const float* __restrict__ t, int rows, int dim, void* __restrict__ out,
                                                     int out_dtype, long long out_ld, long long out_plane) {
  const int half = dim / 2;
  const float e = (float)(9.210340371976184 / (double)(half - 1));  // math.log(10000) / (half - 1)
  const int total = rows * half;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int i = idx % half, r = idx / half;
    const float f = expf(__fmul_rn((float)i, -e));
    const float arg = __fmul_rn(t[r], f);
    store_val(out, out_dtype, (long long)r * out_ld + i, out_plane, sinf(arg));
    store_val(out, out_dtype, (long long)r * out_ld + half + i, out_plane, cosf(arg));
  }
}

// One Euler-Maruyama step (bridge_model.py:363-385), reference operation order.
__global__ void __launch_bounds__(256) sde_step_kernel(float* __restrict__ x, const float* __restrict__ v,
                                                       const float* __restrict__ s, const float* __restrict__ noise,
                                                       int rows, int A, float ginv, float dgg, float eps, float dt,
                                                       float nscale, float d, unsigned long long seed,
                                                       const unsigned long long* __restrict__ seed_dev, int step,
                                                       void* __restrict__ xpad, int xpad_dtype, int xpad_ld,
                                                       long long xpad_plane) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)rows * A;
  if (seed_dev) seed += *seed_dev;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    float z;
    if (noise) {
      z = noise[idx];
    } else {
      curandStatePhilox4_32_10_t st;
      // one whole Philox counter block (4 outputs) per step: curand_normal consumes two outputs (Box-Muller), so an offset
      // of `step` alone would let consecutive steps of the same element share a uniform
      curand_init(seed, (unsigned long long)idx, 4ull * (unsigned long long)step, &st);
      z = curand_normal(&st);
    }
    const float sv = __fmul_rn(s[idx], ginv);                                  // s = s * gamma_inv
    const float b = __fsub_rn(v[idx], __fmul_rn(__fmul_rn(dgg, sv), eps));     // b = v - (dot_gamma*gamma) * s * eps
    const float dW = __fmul_rn(d, z);
    float nx = __fadd_rn(x[idx], __fmul_rn(__fadd_rn(b, __fmul_rn(eps, sv)), dt));   // x + (b + 1.0*eps*s) * dt
    nx = __fadd_rn(nx, __fmul_rn(nscale, dW));
    x[idx] = nx;
    if (xpad) {
      const int a = (int)(idx % A);
      const long long r = idx / A;
      store_val(xpad, xpad_dtype, r * xpad_ld + a, xpad_plane, nx);
    }
  }
}

// q_sample (bridge_model.py:103-107,248-257), reference operation order
__global__ void __launch_bounds__(256) qsample_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                                      const float* __restrict__ step, const float* __restrict__ z_unit, float d,
                                                      int B, int n, int A, float* __restrict__ xt, float* __restrict__ tclip,
                                                      void* __restrict__ xpad, int xpad_dtype, int xpad_ld, long long xpad_plane) {
  const long long total = (long long)B * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / n);
    const int e = (int)(idx - (long long)b * n);
    const float t = fminf(fmaxf(step[b], 0.001f), 1.0f - 0.001f);
    if (e == 0) tclip[b] = t;
    const float gamma = __fmul_rn(__fmul_rn(1.4142f, t), __fsub_rn(1.0f, t));
    const float z = __fmul_rn(d, z_unit[idx]);
    float v = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, t), x0[idx]), __fmul_rn(t, x1[idx]));
    v = __fadd_rn(v, __fmul_rn(gamma, z));
    xt[idx] = v;
    if (xpad) store_val(xpad, xpad_dtype, (idx / A) * xpad_ld + (idx % A), xpad_plane, v);
  }
}

// per-sample loss terms (one block per sample), then the batch mean
__global__ void __launch_bounds__(128) siloss_sample_kernel(const float* __restrict__ bvs, const float* __restrict__ x0,
                                                            const float* __restrict__ x1, const float* __restrict__ z_unit,
                                                            const float* __restrict__ tclip, float d, int B, int n,
                                                            float* __restrict__ per_sample) {
  const int b = blockIdx.x;
  const float t = tclip[b];
  const float gd = 1.4142f * (1.0f - 2.0f * t);
  const float* pb = bvs + (long long)b * n;
  const float* pv = pb + (long long)B * n;
  const float* ps = pv + (long long)B * n;
  float lv = 0.f, ls = 0.f, lb = 0.f;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const long long i = (long long)b * n + e;
    const float pt = x1[i] - x0[i];
    const float z = d * z_unit[i];
    const float v = pv[e], s = ps[e], bb = pb[e];
    lv += 0.5f * v * v - pt * v;
    ls += 0.5f * s * s + z * s;
    lb += 0.5f * bb * bb - (pt + gd * z) * bb;
  }
  __shared__ float red[3][4];
  lv = warp_sum(lv); ls = warp_sum(ls); lb = warp_sum(lb);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][w] = lv; red[1][w] = ls; red[2][w] = lb; }
  __syncthreads();
  if (threadIdx.x < 3) per_sample[threadIdx.x * B + b] = red[threadIdx.x][0] + red[threadIdx.x][1] + red[threadIdx.x][2] + red[threadIdx.x][3];
}
__global__ void __launch_bounds__(96) siloss_mean_kernel(const float* __restrict__ per_sample, int B, float* __restrict__ out) {
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;   // k: 0 = v, 1 = s, 2 = b (rows of per_sample)
  float acc = 0.f;
  for (int i = lane; i < B; i += 32) acc += per_sample[k * B + i];
  acc = warp_sum(acc) / (float)B;
  __shared__ float m[3];
  if (lane == 0) m[k] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    out[1] = m[0];
    out[2] = m[1];
    out[3] = m[2];
    out[0] = m[0] + m[1] + m[2];
  }
}

// fused multi-tensor AdamW + EMA (one thread block per chunk of one tensor).  The gradient may still be in the layout the
// weight-gradient GEMM wrote it in ([rows][taps][c_pad], see vt_opt_tensor), and the updated weight can be emitted as the
// bf16 operand of the next forward pass at the same index, so no unpack / re-pack pass over the parameters exists.
struct OptTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* ema;
  long long numel;
  int taps, c, c_pad, reserved;
  __nv_bfloat16* w_op;
};
struct AdamScalars {
  float lr_wd, beta1, beta2, omb1, omb2, step_size, inv_sqrt_bc2, eps, one_minus_decay, grad_scale;
};
// one element of torch.optim.AdamW (decoupled weight decay first) + the torch_ema shadow update; same operation order in the scalar
// and the 128-bit paths below
__device__ __forceinline__ void adamw_one(const AdamScalars& c, float g, float& p, float& m, float& v, float* s) {
  g *= c.grad_scale;
  p *= c.lr_wd;
  m = c.beta1 * m + c.omb1 * g;
  v = c.beta2 * v + c.omb2 * g * g;
  const float denom = sqrtf(v) * c.inv_sqrt_bc2 + c.eps;
  p -= c.step_size * (m / denom);
  if (s) *s = *s - c.one_minus_decay * (*s - p);
}
__global__ void __launch_bounds__(256) adamw_ema_kernel(const OptTensor* __restrict__ tensors, const long long* __restrict__ chunks,
                                                        int chunk_elems, float lr, float beta1, float beta2, float eps, float wd,
                                                        float bc1, float bc2, float ema_decay, float grad_scale) {
  const OptTensor t = tensors[chunks[2 * blockIdx.x]];
  const long long off = chunks[2 * blockIdx.x + 1];
  const long long end = min(off + (long long)chunk_elems, t.numel);
  AdamScalars c;
  c.lr_wd = 1.0f - lr * wd; c.beta1 = beta1; c.beta2 = beta2; c.omb1 = 1.0f - beta1; c.omb2 = 1.0f - beta2;
  c.step_size = lr / bc1; c.inv_sqrt_bc2 = rsqrtf(bc2); c.eps = eps; c.one_minus_decay = 1.0f - ema_decay; c.grad_scale = grad_scale;
  const unsigned ct = (unsigned)(t.c * t.taps);
  auto gidx = [&](long long i) -> long long {   // p[(r * c + ci) * taps + k]  <->  g[(r * taps + k) * c_pad + ci]
    if (t.taps <= 0) return i;
    const long long r = i / ct;
    const unsigned rem = (unsigned)(i - r * ct);
    const unsigned ci = rem / (unsigned)t.taps, k = rem - ci * (unsigned)t.taps;
    return (r * t.taps + k) * t.c_pad + ci;
  };
  // 128-bit path: four consecutive parameter elements per thread (p, m, v, shadow as float4; the gradient too when it shares the
  // parameter's layout, four gathers when it sits in the weight-gradient GEMM's layout) -- 36 B per parameter of HBM traffic
  const uintptr_t al = (uintptr_t)t.p | (uintptr_t)t.m | (uintptr_t)t.v | (uintptr_t)t.ema | (t.taps <= 0 ? (uintptr_t)t.g : 0);
  if (((t.numel | off) & 3) == 0 && (al & 15) == 0 && !t.w_op) {
    for (long long i = off + 4LL * threadIdx.x; i < end; i += 4LL * blockDim.x) {
      float4 p = *reinterpret_cast<const float4*>(t.p + i), m = *reinterpret_cast<const float4*>(t.m + i);
      float4 v = *reinterpret_cast<const float4*>(t.v + i);
      float4 s = t.ema ? *reinterpret_cast<const float4*>(t.ema + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 g;
      if (t.taps <= 0) g = *reinterpret_cast<const float4*>(t.g + i);
      else g = make_float4(t.g[gidx(i)], t.g[gidx(i + 1)], t.g[gidx(i + 2)], t.g[gidx(i + 3)]);
      adamw_one(c, g.x, p.x, m.x, v.x, t.ema ? &s.x : nullptr);
      adamw_one(c, g.y, p.y, m.y, v.y, t.ema ? &s.y : nullptr);
      adamw_one(c, g.z, p.z, m.z, v.z, t.ema ? &s.z : nullptr);
      adamw_one(c, g.w, p.w, m.w, v.w, t.ema ? &s.w : nullptr);
      *reinterpret_cast<float4*>(t.p + i) = p;
      *reinterpret_cast<float4*>(t.m + i) = m;
      *reinterpret_cast<float4*>(t.v + i) = v;
      if (t.ema) *reinterpret_cast<float4*>(t.ema + i) = s;
    }
    return;
  }
  for (long long i = off + threadIdx.x; i < end; i += blockDim.x) {
    const long long j = gidx(i);
    float p = t.p[i], m = t.m[i], v = t.v[i];
    float s = t.ema ? t.ema[i] : 0.f;
    adamw_one(c, t.g[j], p, m, v, t.ema ? &s : nullptr);
    t.p[i] = p;
    t.m[i] = m;
    t.v[i] = v;
    if (t.w_op) t.w_op[j] = __float2bfloat16(p);
    if (t.ema) t.ema[i] = s;
  }
}

// Operand re-pack after an optimizer step (unet_train.LossBackwardProgram.setup_gather): every packed GEMM operand is a fixed
// rearrangement of parameter elements, dst[i] = arena[map[i]] (map points at a zero slot for padding).  One launch for all
// operands: chunk -> (record, offset); a thread converts 8 consecutive destination elements (32-byte map load, 16-byte store).
struct GatherRec {
  void* dst;
  const int* map;
  long long n;
  int bf16;        // 1: destination is bf16, 0: fp32
  int pad;
};
constexpr int GATHER_CHUNK = 8192;
__global__ void __launch_bounds__(256) gather_repack_kernel(const GatherRec* __restrict__ recs, const long long* __restrict__ chunks,
                                                            const float* __restrict__ arena) {
  const GatherRec r = recs[chunks[2 * blockIdx.x]];
  const long long off = chunks[2 * blockIdx.x + 1];
  const long long end = min(off + (long long)GATHER_CHUNK, r.n);
  for (long long i = off + 8LL * threadIdx.x; i < end; i += 8LL * blockDim.x) {
    if (i + 8 <= end) {
      const int4 m0 = *reinterpret_cast<const int4*>(r.map + i), m1 = *reinterpret_cast<const int4*>(r.map + i + 4);
      const float v0 = arena[m0.x], v1 = arena[m0.y], v2 = arena[m0.z], v3 = arena[m0.w];
      const float v4 = arena[m1.x], v5 = arena[m1.y], v6 = arena[m1.z], v7 = arena[m1.w];
      if (r.bf16) {
        uint4 w;
        w.x = pack_bf16x2(v0, v1);
        w.y = pack_bf16x2(v2, v3);
        w.z = pack_bf16x2(v4, v5);
        w.w = pack_bf16x2(v6, v7);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(r.dst) + i) = w;
      } else {
        float* d = reinterpret_cast<float*>(r.dst) + i;
        *reinterpret_cast<float4*>(d) = make_float4(v0, v1, v2, v3);
        *reinterpret_cast<float4*>(d + 4) = make_float4(v4, v5, v6, v7);
      }
    } else {
      for (long long j = i; j < end; ++j) {
        const float v = arena[r.map[j]];
        if (r.bf16) reinterpret_cast<__nv_bfloat16*>(r.dst)[j] = __float2bfloat16(v);
        else reinterpret_cast<float*>(r.dst)[j] = v;
      }
    }
  }
}

// bicubic resize (A = -0.75, align_corners = False) of the patch position embeddings, HF:57-95
__device__ __forceinline__ void cubic_coeffs(float t, float* w) {
  const float A = -0.75f;
  float x = t + 1.0f;
  w[0] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
  x = t;
  w[1] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
  x = 1.0f - t;
  w[2] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
  x = 2.0f - t;
  w[3] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
}
__global__ void pos_resize_kernel(const float* __restrict__ src, int s, float* __restrict__ dst, int nh, int nw, int D) {
  const long long total = (long long)nh * nw * D;
  const float sh = (float)s / (float)nh, sw = (float)s / (float)nw;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(idx % D);
    const int p = (int)(idx / D);
    const int oy = p / nw, ox = p - oy * nw;
    const float ry = sh * ((float)oy + 0.5f) - 0.5f, rx = sw * ((float)ox + 0.5f) - 0.5f;
    const float fy = floorf(ry), fx = floorf(rx);
    const int iy = (int)fy, ix = (int)fx;
    float wy[4], wx[4];
    cubic_coeffs(ry - fy, wy);
    cubic_coeffs(rx - fx, wx);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), s - 1);
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(ix - 1 + b, 0), s - 1);
        row += src[((long long)yy * s + xx) * D + d] * wx[b];
      }
      acc += row * wy[a];
    }
    dst[idx] = acc;
  }
}

}  // namespace vt
