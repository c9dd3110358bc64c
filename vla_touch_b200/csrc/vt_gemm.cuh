// tcgen05 / TMA implicit-GEMM kernel shared by every contraction on the refinement path:
//   * DinoV2 linears (patch-embed, QKV, attention-out, fc1, fc2)      HF Dinov2Layer, HF:199-201,272-278,312-328
//   * state-encoder / FiLM / time-MLP linears                          bridge_controller.py:42-48, conditional_unet_1D.py:76-80,137-142
//   * Conv1d k5 / k3-stride-2 / 1x1 and ConvTranspose1d k4-stride-2    conditional_unet_1D.py:22-55
// with the op that follows fused into the epilogue: bias, GELU/Mish, LayerScale + residual, or
// GroupNorm(8) + Mish + FiLM / residual (Conv1dBlock + ConditionalResidualBlock1D, :40-105).
//
// One CTA computes a 128 x BN output tile.  Warp 0 = TMA producer, warp 1 = TMEM allocator + UMMA issuer,
// warps 2..5 = epilogue (one TMEM lane quarter each).  A-tiles are fetched with a 5-D tensor map
// (C, phase, T, B, G): a conv tap is a TMA box whose T coordinate is shifted (out-of-bounds rows are
// zero-filled by the TMA unit = the conv's zero padding), a stride-2 conv reads the even/odd phase, and
// v_net / s_net are the G dimension, so no im2col buffer ever exists.
#pragma once
#include "vt_elem.cuh"
#include "vt_ptx.cuh"

namespace vt {

constexpr int GEMM_BM = 128;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_A_STAGE_BYTES = GEMM_BM * 128;
constexpr int GEMM_MAX_TAPS = 8;

enum : int { ACT_NONE = 0, ACT_GELU = 1, ACT_MISH = 2 };
enum : int { EPI_LINEAR = 0, EPI_GN = 1 };

struct GemmArgs {
  CUtensorMap tmA;  // 5-D (C, P, T, B, G), box (KE, 1, Tbox, Bbox, 1), SWIZZLE_128B
  CUtensorMap tmB;  // 2-D (Ktot, G*n_pad), box (KE, BN), SWIZZLE_128B
  // ---- K loop: k-block i -> (pass, tap, cb) ----
  int passes;   // 1; 3 = split-tf32 (hi*hi, lo*hi, hi*lo)
  int taps;     // conv taps (1 for a plain GEMM)
  int cblocks;  // K-blocks per tap
  int a_c0;     // first channel of A (elements)
  int a_plane;  // hi->lo plane distance in A's C dimension (elements), split mode
  int b_plane;  // hi->lo plane distance in B's K dimension (elements), split mode
  int a_g_mul;  // 0: all groups read group 0 of A (shared input), 1: per-group A
  int tap_p[GEMM_MAX_TAPS];
  int tap_t[GEMM_MAX_TAPS];
  int m_t_step, m_b_step;  // A-tile (T, B) coordinate step per m-tile
  int rows_valid;          // rows of the 128-row tile backed by the TMA box (Tbox*Bbox)
  int a_box_bytes;         // bytes one A box delivers (rows_valid * 128)
  int n_pad;               // rows of B per group
  // ---- epilogue ----
  int M_total;  // logical rows per group
  int N;        // valid columns per group
  int row_div;  // logical row -> (q = row / row_div, rem = row % row_div)
  long long out_q, out_r, out_off, out_g;  // out row = q*out_q + rem*out_r + out_off ; group stride in elements
  int ldc;
  long long out_plane;  // >0: also store the tf32 lo-plane at +out_plane (TOut=float split mode)
  int vec;              // 1: rows are 16-byte aligned (ldc/ldres/pointers), 128-bit epilogue accesses allowed
  void* out;
  const float* bias;  // [G][n_pad] or null
  int act;
  const float* colscale;  // [N] LayerScale, or null
  const void* res;        // residual: fp32 (EPI_LINEAR) / activation type (EPI_GN); null = none
  long long res_q, res_r, res_off, res_g;
  long long res_plane;  // >0: the residual is stored as tf32 hi|lo planes; both are added
  int ldres;
  // GroupNorm + Mish + FiLM (EPI_GN)
  const float* gn_gamma;  // [G][n_pad]
  const float* gn_beta;   // [G][n_pad]
  int gn_gs_log2;         // log2(channels per group): 5 or 6
  int gn_rows;            // rows per sample (= Tbox)
  float gn_eps;
  const float* film_c;  // [G][B][film_ld] per-sample FiLM (cond part + bias): scale at n, shift at film_C+n
  const float* film_t;  // [G][film_ld]    per-step FiLM (time part)
  long long film_g;     // group stride of film_c (elements)
  long long film_tg;    // group stride of film_t (elements)
  int film_ld, film_C, film_off;
};

template <typename TIn>
struct InTraits;
template <>
struct InTraits<__nv_bfloat16> {
  static constexpr int KE = 64;  // elements per 128-byte swizzle row
  static constexpr uint32_t FMT = UMMA_FMT_BF16;
};
template <>
struct InTraits<float> {
  static constexpr int KE = 32;
  static constexpr uint32_t FMT = UMMA_FMT_TF32;
};

template <int BN, int STAGES>
constexpr int gemm_smem_bytes() {
  return 1024 /*align slack*/ + STAGES * (GEMM_A_STAGE_BYTES + BN * 128) + 256 /*barriers*/ + 5 * BN * 4 /*col vecs*/ +
         128 * 4 * 8 /*GN partials*/ + 32 * 4 * 8 /*GN stats*/;
}

template <typename TOut>
__device__ __forceinline__ void store_chunk8(TOut* p, const float* y);
template <>
__device__ __forceinline__ void store_chunk8<__nv_bfloat16>(__nv_bfloat16* p, const float* y) {
  uint4 v;
  v.x = pack_bf16x2(y[0], y[1]);
  v.y = pack_bf16x2(y[2], y[3]);
  v.z = pack_bf16x2(y[4], y[5]);
  v.w = pack_bf16x2(y[6], y[7]);
  *reinterpret_cast<uint4*>(p) = v;
}
template <>
__device__ __forceinline__ void store_chunk8<float>(float* p, const float* y) {
  *reinterpret_cast<float4*>(p) = make_float4(y[0], y[1], y[2], y[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(y[4], y[5], y[6], y[7]);
}
__device__ __forceinline__ void load_res8(const float* p, float* r) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
__device__ __forceinline__ void load_res8(const __nv_bfloat16* p, float* r) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    r[2 * i] = f.x;
    r[2 * i + 1] = f.y;
  }
}

template <typename TIn, int BN, int MODE, typename TOut, int STAGES, bool PRECISE>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_tc_kernel(const __grid_constant__ GemmArgs a) {
  constexpr int KE = InTraits<TIn>::KE;
  constexpr int B_STAGE_BYTES = BN * 128;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t IDESC = umma_idesc(InTraits<TIn>::FMT, BN);
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "BN");
  static_assert(MODE != EPI_GN || BN == 128, "GN epilogue assumes a 128-column tile");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * GEMM_A_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* colv = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 256);  // [5][BN]
  float2* gn_part = reinterpret_cast<float2*>(colv + 5 * BN);                       // [128][4]
  float2* gn_stat = gn_part + 128 * 4;                                              // [32][4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, g = blockIdx.z;
  const int nk = a.passes * a.taps * a.cblocks;
  const int n0 = n_tile * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmA);
    tma_prefetch_desc(&a.tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      mbar_init(tmem_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      const int per_pass = a.taps * a.cblocks;
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const int pass = i / per_pass;
        const int j = i - pass * per_pass;
        const int tp = j / a.cblocks;
        const int cb = j - tp * a.cblocks;
        const int pa = (pass == 1) ? a.a_plane : 0;
        const int pb = (pass == 2) ? a.b_plane : 0;
        mbar_arrive_expect_tx(&full[s], a.a_box_bytes + B_STAGE_BYTES);
        tma_load_5d(sA + s * GEMM_A_STAGE_BYTES, &a.tmA, &full[s], a.a_c0 + pa + cb * KE, a.tap_p[tp],
                    m_tile * a.m_t_step + a.tap_t[tp], m_tile * a.m_b_step, g * a.a_g_mul);
        tma_load_2d(sB + s * B_STAGE_BYTES, &a.tmB, &full[s], pb + j * KE, g * a.n_pad + n0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------ UMMA issuer ------------------------------
    if (lane == 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t adesc = umma_smem_desc_sw128(smem_u32(sA + s * GEMM_A_STAGE_BYTES));
        const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(sB + s * B_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x (32 bytes of K) per 128-byte swizzle row
          if constexpr (sizeof(TIn) == 2)
            umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, IDESC, (i | k) != 0);
          else
            umma_tf32(tmem_base, adesc + 2 * k, bdesc + 2 * k, IDESC, (i | k) != 0);
        }
        umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full);
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int et = threadIdx.x - 64;  // 0..127
    const int quarter = warp & 3;     // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;
    // stage per-column vectors while the main loop runs
    {
      const long long gcol = (long long)g * a.n_pad + n0;
      for (int c = et; c < BN; c += 128) {
        colv[c] = a.bias ? a.bias[gcol + c] : 0.f;
        if (MODE == EPI_LINEAR) {
          colv[BN + c] = (a.colscale && (n0 + c) < a.N) ? a.colscale[n0 + c] : 1.f;
        } else {
          colv[BN + c] = a.gn_gamma[gcol + c];
          colv[2 * BN + c] = a.gn_beta[gcol + c];
          const bool f = a.film_t != nullptr;
          const long long fo = (long long)g * a.film_tg + a.film_off + n0 + c;
          colv[3 * BN + c] = f ? a.film_t[fo] : 0.f;
          colv[4 * BN + c] = f ? a.film_t[fo + a.film_C] : 0.f;
        }
      }
    }
    named_bar_sync(1, 128);

    const long long grow = (long long)m_tile * a.rows_valid + r;
    const bool valid = (r < a.rows_valid) && (grow < a.M_total);
    const int q = (int)(grow / a.row_div);
    const int rem = (int)(grow - (long long)q * a.row_div);
    TOut* outp = reinterpret_cast<TOut*>(a.out) + (long long)g * a.out_g +
                 ((long long)q * a.out_q + (long long)rem * a.out_r + a.out_off) * a.ldc + n0;
    const long long res_row = (long long)q * a.res_q + (long long)rem * a.res_r + a.res_off;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);

    mbar_wait(tmem_full, 0);
    tc_fence_after();

    if constexpr (MODE == EPI_LINEAR) {
      const float* resp =
          a.res ? reinterpret_cast<const float*>(a.res) + (long long)g * a.res_g + res_row * a.ldres + n0 : nullptr;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j8 = 0; j8 < 32; j8 += 8) {
            const int cc = c + j8;
            if (n0 + cc >= a.N) break;
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float x = __uint_as_float(v[j8 + j]) + colv[cc + j];
              if (a.act == ACT_GELU) x = gelu_erf(x);
              else if (a.act == ACT_MISH) x = PRECISE ? mish_precise(x) : mish_f(x);
              y[j] = x * colv[BN + cc + j];
            }
            if (a.vec && n0 + cc + 8 <= a.N) {
              if (resp) {
                float rr[8];
                load_res8(resp + cc, rr);
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] += rr[j];
              }
              if constexpr (sizeof(TOut) == 4) {
                if (a.out_plane > 0) {
                  float hi[8], lo[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    hi[j] = tf32_hi(y[j]);
                    lo[j] = y[j] - hi[j];
                  }
                  store_chunk8<TOut>(outp + cc, hi);
                  store_chunk8<TOut>(outp + a.out_plane + cc, lo);
                } else {
                  store_chunk8<TOut>(outp + cc, y);
                }
              } else {
                store_chunk8<TOut>(outp + cc, y);
              }
            } else {  // ragged last columns (N not a multiple of 8): scalar
              for (int j = 0; j < 8 && n0 + cc + j < a.N; ++j) {
                float yy = y[j] + (resp ? resp[cc + j] : 0.f);
                if constexpr (sizeof(TOut) == 4) {
                  if (a.out_plane > 0) {
                    const float hi = tf32_hi(yy);
                    outp[cc + j] = hi;
                    outp[a.out_plane + cc + j] = yy - hi;
                  } else {
                    outp[cc + j] = yy;
                  }
                } else {
                  outp[cc + j] = __float2bfloat16(yy);
                }
              }
            }
          }
        }
      }
    } else {
      // ---------------- GroupNorm(8) + Mish + FiLM | residual ----------------
      // Tile = all T rows of `rows_valid / gn_rows` samples x 128 channels = whole groups, so the
      // statistics are tile-local.  Pass 1: per-row partial sums per 32-column chunk.
      float s1[4], s2[4];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        tmem_ld32(taddr + ch * 32, v);
        tmem_ld_wait();
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]) + colv[ch * 32 + j];
          a1 += x;
          a2 = fmaf(x, x, a2);
        }
        s1[ch] = valid ? a1 : 0.f;
        s2[ch] = valid ? a2 : 0.f;
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) gn_part[r * 4 + ch] = make_float2(s1[ch], s2[ch]);
      named_bar_sync(1, 128);
      {
        // thread et -> (sample et/4, chunk et%4); chunks are merged pairwise when a group spans 64 channels
        const int smp = et >> 2, ch = et & 3;
        const int nsamp = a.rows_valid / a.gn_rows;
        float t1 = 0.f, t2 = 0.f;
        if (smp < nsamp) {
          const int r0 = smp * a.gn_rows;
          for (int t = 0; t < a.gn_rows; ++t) {
            float2 p = gn_part[(r0 + t) * 4 + ch];
            t1 += p.x;
            t2 += p.y;
          }
        }
        if (a.gn_gs_log2 == 6) {
          t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
          t2 += __shfl_xor_sync(0xffffffffu, t2, 1);
        }
        if (smp < nsamp && smp < 32) {
          const float cnt = (float)(a.gn_rows << a.gn_gs_log2);
          const float mean = t1 / cnt;
          const float var = fmaxf(t2 / cnt - mean * mean, 0.f);
          gn_stat[smp * 4 + ch] = make_float2(mean, rsqrtf(var + a.gn_eps));
        }
      }
      named_bar_sync(1, 128);
      const int smp = valid ? (r / a.gn_rows) : 0;
      const float* filmp = a.film_c ? a.film_c + (long long)g * a.film_g + (long long)q * a.film_ld + a.film_off + n0
                                    : nullptr;
      const TOut* resp =
          a.res ? reinterpret_cast<const TOut*>(a.res) + (long long)g * a.res_g + res_row * a.ldres + n0 : nullptr;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        tmem_ld32(taddr + ch * 32, v);
        tmem_ld_wait();
        if (valid) {
          const float2 st = gn_stat[smp * 4 + ch];
#pragma unroll
          for (int j8 = 0; j8 < 32; j8 += 8) {
            const int cc = ch * 32 + j8;
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float x = __uint_as_float(v[j8 + j]) + colv[cc + j];
              x = (x - st.x) * st.y * colv[BN + cc + j] + colv[2 * BN + cc + j];
              y[j] = PRECISE ? mish_precise(x) : mish_f(x);
            }
            if (filmp) {
              float sc[8], sh[8];
              load_res8(filmp + cc, sc);
              load_res8(filmp + a.film_C + cc, sh);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                y[j] = (sc[j] + colv[3 * BN + cc + j]) * y[j] + (sh[j] + colv[4 * BN + cc + j]);
            }
            if (resp) {
              float rr[8];
              load_res8(resp + cc, rr);
#pragma unroll
              for (int j = 0; j < 8; ++j) y[j] += rr[j];
              if (a.res_plane > 0) {
                load_res8(resp + a.res_plane + cc, rr);
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] += rr[j];
              }
            }
            if constexpr (sizeof(TOut) == 4) {
              if (a.out_plane > 0) {
                float hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  hi[j] = tf32_hi(y[j]);
                  lo[j] = y[j] - hi[j];
                }
                store_chunk8<TOut>(outp + cc, hi);
                store_chunk8<TOut>(outp + a.out_plane + cc, lo);
              } else {
                store_chunk8<TOut>(outp + cc, y);
              }
            } else {
              store_chunk8<TOut>(outp + cc, y);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace vt
