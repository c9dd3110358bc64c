// tcgen05 / TMA implicit-GEMM kernel shared by every contraction on the refinement path:
//   * DinoV2 linears (patch-embed, QKV, attention-out, fc1, fc2)      HF Dinov2Layer, HF:199-201,272-278,312-328
//   * state-encoder / FiLM / time-MLP linears                          bridge_controller.py:42-48, conditional_unet_1D.py:76-80,137-142
//   * Conv1d k5 / k3-stride-2 / 1x1 and ConvTranspose1d k4-stride-2    conditional_unet_1D.py:22-55
// with the op that follows fused into the epilogue: bias, GELU/Mish, LayerScale + residual, or
// GroupNorm(8) + Mish + FiLM / residual (Conv1dBlock + ConditionalResidualBlock1D, :40-105).
//
// PERSISTENT kernel: one CTA per SM (or one CTA PAIR per TPC, CTAS = 2) walks a static round-robin list of output tiles.
//   warp 0      TMA producer: A/B k-blocks through a STAGES-deep mbarrier ring that runs ahead across tiles
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer; two accumulators in TMEM (2 x BN columns)
//   warps 2-9   epilogue: both warpgroups work on every tile (column halves), accumulator = local tile index & 1,
// so the epilogue of tile i (tcgen05.ld, GroupNorm / activation math, global stores) overlaps the main loop of tile i+1.
// CTA pairs: tcgen05 cta_group::2, M = 256 per MMA; each CTA stages its own 128 rows of A and half of the B tile, the
// leader CTA issues the MMAs and owns the `full` barriers, commits are multicast to both CTAs.
// A-tiles are fetched with a 5-D tensor map (C, phase, T, B, G): a conv tap is a TMA box whose T coordinate is
// shifted (out-of-bounds rows are zero-filled by the TMA unit = the conv's zero padding), a stride-2 conv reads the
// even/odd phase, and v_net / s_net are the G dimension, so no im2col buffer ever exists.
#pragma once
#include "vt_elem.cuh"
#include "vt_ptx.cuh"

namespace vt {

constexpr int GEMM_BM = 128;
// warp 0 TMA producer, warp 1 MMA issuer, then EW = 8 or 12 epilogue warps (two or three per TMEM lane quarter)
__host__ __device__ constexpr int GEMM_THREADS(int EW) { return 64 + 32 * EW; }
constexpr int GEMM_A_STAGE_BYTES = GEMM_BM * 128;
constexpr int GEMM_MAX_TAPS = 8;

enum : int { ACT_NONE = 0, ACT_GELU = 1, ACT_MISH = 2 };

// Developer instrumentation.  Release builds (VT_DEBUG_KNOBS == 0, the default) compile every knob and timestamp out of the
// kernels; `python -m vla_touch_b200.build --debug-knobs` builds a library in which the env variable VT_GEMM_DEBUG selects:
// 1 skip TMA loads + MMAs, 2 skip the epilogue body, 4 skip TMA loads only, 8 skip MMAs only, 16 no epilogue stores,
// 32 no TMEM loads / no transposition, 128 / 256 clock64() stamps of one epilogue warp / the MMA thread of CTA 0 (read back with
// vt_debug_timestamps; tools/gemm_debug.py, tools/epi_trace.py).
#ifndef VT_DEBUG_KNOBS
#define VT_DEBUG_KNOBS 0
#endif
constexpr bool kDbg = VT_DEBUG_KNOBS != 0;
constexpr int VT_DBG_TS = 2048;
__device__ long long vt_dbg_ts[VT_DBG_TS];
__device__ int vt_dbg_n;
// Persistent-kernel timeline (VT_GEMM_DEBUG bit 512, tools/persist_trace.py): per worker and local tile PTRACE_SLOTS clock64()
// stamps by the leader CTA's TMA thread, MMA thread and first epilogue thread; vt_ptrace_cal holds (clock64, globaltimer)
// pairs per worker at kernel entry and exit to put the SM clocks on one time base.
constexpr int PTRACE_SLOTS = 12, PTRACE_TILES = 256, PTRACE_WORKERS = 80;
#if VT_DEBUG_KNOBS
__device__ long long vt_ptrace[PTRACE_WORKERS * PTRACE_TILES * PTRACE_SLOTS];
__device__ long long vt_ptrace_cal[PTRACE_WORKERS * 4];
#endif
__device__ __forceinline__ void ptrace(long long* base, int slot) {
#if VT_DEBUG_KNOBS
  if (base) base[slot] = clock64();
#endif
}
__device__ __forceinline__ void ptrace_val(long long* base, int slot, long long v) {
#if VT_DEBUG_KNOBS
  if (base) base[slot] = v;
#endif
}
// The entry counter lives in a register of the recording thread (`n`): a stamp is two fire-and-forget stores.
__device__ __forceinline__ void dbg_stamp(bool on, int& n, int tag) {
  if (kDbg && on && n + 1 < VT_DBG_TS) {
    vt_dbg_ts[n] = tag;
    vt_dbg_ts[n + 1] = clock64();
    n += 2;
    vt_dbg_n = n;
  }
}

enum : int { EPI_LINEAR = 0, EPI_GN = 1 };

struct GemmArgs {
  CUtensorMap tmA;  // 5-D (C, P, T, B, G), box (KE, 1, Tbox, Bbox, 1), SWIZZLE_128B
  CUtensorMap tmB;  // 2-D (Ktot, G*n_pad), box (KE, BN), SWIZZLE_128B
  int debug;        // developer knobs (VT_DEBUG_KNOBS builds only, see above)
  // ---- tiles ----
  int n_tiles, m_tiles, total_tiles;  // tile id = (g * m_tiles + m_tile) * n_tiles + n_tile
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
  int fast;             // 1: every tile is full width, no hi|lo planes, 32-bit row offsets: the transposing epilogue applies
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
  float* raw;          // EPI_GN fast path: fp32 copy of acc + bias at raw[g * raw_g + logical row * raw_ld + n], or null
  long long raw_g;
  int raw_ld;
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

// Epilogue scratch.  LINEAR: two column-vector sets (one per accumulator) of [2][BN].  GroupNorm: one set of [5][BN]
// (bias, gamma, beta, FiLM time scale / shift), per warpgroup the row partials [128][4] and sample sums [64][4] (float2),
// and the staged per-sample FiLM table [GEMM_FILM_SAMPLES][2][BN].
constexpr int GEMM_FILM_SAMPLES = 8;
__host__ __device__ constexpr int GEMM_COLV_FLOATS(int BN, int MODE) { return MODE == 0 ? 2 * BN : 5 * BN; }
__host__ __device__ constexpr int GEMM_SCRATCH_FLOATS(int BN, int MODE) {
  return MODE == 0 ? 2 * GEMM_COLV_FLOATS(BN, MODE)
                   : GEMM_COLV_FLOATS(BN, MODE) + 2 * (128 + 64) * 4 * 2 + GEMM_FILM_SAMPLES * 2 * BN;
}

// (A TMA-store epilogue -- 32-row x 128-byte boxes staged in shared memory -- was measured on B200 and dropped: the main loop
// is bound by the TMA -> MMA -> commit round trip divided by the ring depth, so 64 KB buy more as two extra operand stages.)
// Per-warp 32 x 32 fp32 transposition buffer of the coalescing epilogue (8 epilogue warps x 4 KB).
__host__ __device__ constexpr int GEMM_XPOSE_BYTES(int BN, int MODE, int EW = 8) { return (MODE == 0 && BN >= 128) ? EW * 4096 : 0; }

// KA = number of 128-byte-wide K atoms (64 bf16 / 32 tf32 elements of K each) per pipeline stage.
// One mbarrier wait + tcgen05.commit per stage costs the single issuing thread ~450 cycles (measured), more than the MMA
// time of one atom at BN=128, so a stage carries KA atoms (2 x 4 UMMA instructions per barrier round).
template <int BN, int STAGES, int MODE, int KA, int CTAS = 1, int EW = 8>
constexpr int gemm_smem_bytes() {
  return 1024 /*align slack*/ + STAGES * KA * (GEMM_A_STAGE_BYTES + (BN / CTAS) * 128) +
         GEMM_XPOSE_BYTES(BN, MODE, EW) + 256 /*barriers*/ + GEMM_SCRATCH_FLOATS(BN, MODE) * 4;
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

template <typename TOut>
__device__ __forceinline__ void store_split8(TOut* outp, long long plane, const float* y) {
  if constexpr (sizeof(TOut) == 4) {
    if (plane > 0) {
      float hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        hi[j] = tf32_hi(y[j]);
        lo[j] = y[j] - hi[j];
      }
      store_chunk8<TOut>(outp, hi);
      store_chunk8<TOut>(outp + plane, lo);
      return;
    }
  }
  store_chunk8<TOut>(outp, y);
}

// Per-tile state shared by the epilogue flavours.
struct EpiTile {
  int n0, g, r;            // first column, group, tile row (== TMEM lane)
  long long grow;          // logical row
  bool valid;
  int dbg_n;               // developer instrumentation: timestamps recorded so far (see dbg_stamp)
  long long* tr;           // developer instrumentation: this tile's row of the persistent kernel's timeline (null = off)
  int q, rem;
  uint32_t taddr;          // TMEM address of this thread's lane, column 0 of the accumulator
};

// ------------------------------------------------------------------------------------------------------------
// EPI_LINEAR: bias -> activation -> column scale -> (+ fp32 residual) -> store.  The residual chunk of the NEXT
// 32 columns is requested before the current chunk is processed, the first one before the accumulator is ready.
// ------------------------------------------------------------------------------------------------------------
template <int BN, typename TOut, bool PRECISE>
__device__ __forceinline__ void epilogue_linear(const GemmArgs& a, const EpiTile& t, const float* colv, uint64_t* acc_full,
                                                uint32_t acc_parity, int c_begin, int c_end) {
  // kernel parameters used inside the column loops, pinned in registers (the struct lives in the constant bank and is
  // otherwise re-read after every asm "memory" clobber)
  const int N = a.N, act = a.act, dbg = kDbg ? a.debug : 0;
  const bool vec = a.vec != 0;
  const long long oplane = a.out_plane;

  TOut* outp = reinterpret_cast<TOut*>(a.out) + (long long)t.g * a.out_g +
               ((long long)t.q * a.out_q + (long long)t.rem * a.out_r + a.out_off) * a.ldc + t.n0;
  const long long res_row = (long long)t.q * a.res_q + (long long)t.rem * a.res_r + a.res_off;
  const float* resp = a.res ? reinterpret_cast<const float*>(a.res) + (long long)t.g * a.res_g + res_row * a.ldres + t.n0 : nullptr;
  const bool vres = resp && vec && t.valid;
  float4 rb[8];
  auto fetch = [&](int c) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      rb[i] = (vres && t.n0 + c + 4 * i + 4 <= N) ? *reinterpret_cast<const float4*>(resp + c + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  fetch(c_begin);
  mbar_wait(acc_full, acc_parity);
  tc_fence_after();
  ptrace(t.tr, 7);
#pragma unroll 1
  for (int c = c_begin; c < c_end; c += 32) {
    uint32_t v[32];
    if (dbg & 32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0u;
    } else {
      tmem_ld32(t.taddr + c, v);
    }
    float4 rc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rc[i] = rb[i];
    if (c + 32 < c_end) fetch(c + 32);
    tmem_ld_wait();
    if (t.valid) {
#pragma unroll
      for (int j8 = 0; j8 < 32; j8 += 8) {
        const int cc = c + j8;
        if (t.n0 + cc >= N) break;
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(v[j8 + j]) + colv[cc + j];
        if (act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = PRECISE ? gelu_erf(y[j]) : gelu_fast(y[j]);
        } else if (act == ACT_MISH) {
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = PRECISE ? mish_precise(y[j]) : mish_f(y[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] *= colv[BN + cc + j];
        if (dbg & 16) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc += y[j];
          if (acc == 123.456f) outp[cc] = TOut(acc);   // keeps the math alive without producing output traffic
        } else if (vec && t.n0 + cc + 8 <= N) {
          const float4 r0 = rc[j8 / 4], r1 = rc[j8 / 4 + 1];
          y[0] += r0.x; y[1] += r0.y; y[2] += r0.z; y[3] += r0.w;
          y[4] += r1.x; y[5] += r1.y; y[6] += r1.z; y[7] += r1.w;
          store_split8<TOut>(outp + cc, oplane, y);
        } else {  // ragged last columns / unaligned rows: scalar
          for (int j = 0; j < 8 && t.n0 + cc + j < N; ++j) {
            const float yy = y[j] + (resp ? resp[cc + j] : 0.f);
            if constexpr (sizeof(TOut) == 4) {
              if (oplane > 0) {
                const float hi = tf32_hi(yy);
                outp[cc + j] = hi;
                outp[oplane + cc + j] = yy - hi;
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
}

// ------------------------------------------------------------------------------------------------------------
// EPI_LINEAR, coalescing variant (a.fast): the accumulator arrives one ROW per lane (tcgen05.ld 32x32b), which makes
// every direct global access touch 32 different lines.  Each warp therefore transposes its 32 x 32 chunk through a
// swizzled 4 KB shared-memory buffer; afterwards a lane owns NC = 16 / sizeof(TOut) consecutive columns of 32 / NC rows,
// so that one warp-wide load / store covers whole 128-byte (fp32) or 64-byte (bf16) row segments, the per-column
// parameters are plain (L1-resident) vector loads into registers, and the math is packed f32x2:
//   y = act(acc + bias) * colscale + residual.
// Activation and residual are compile-time so the chunk loop is branch-free.  The residual rows of the next chunk are in
// flight while the current one is processed; the first chunk's are requested before the accumulator is ready.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 gelu_fast2(float2 x) {   // same fit as gelu_fast, two lanes per instruction
  float2 x2 = fmul2(x, x);
  x2.x = fminf(x2.x, 36.f);
  x2.y = fminf(x2.y, 36.f);
  float2 p = ffma2(x2, make_float2(-3.51516795e-4f, -3.51516795e-4f), make_float2(3.70056461e-2f, 3.70056461e-2f));
  p = ffma2(x2, p, make_float2(7.97507884e-1f, 7.97507884e-1f));
  const float2 u = fmul2(x, p);
  const float2 th = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 hx = fmul2(x, make_float2(0.5f, 0.5f));
  return ffma2(hx, th, hx);
}

// STATS: st[2i] / st[2i + 1] accumulate the sum / sum of squares of the values this lane stores for its i-th row (the
// fused MLP kernel turns them into the next LayerNorm's row statistics); rows that do not exist accumulate garbage.
template <int BN, typename TOut, bool PRECISE, int ACT, bool HAS_RES, bool STATS = false>
__device__ __forceinline__ void epilogue_linear_t(const GemmArgs& a, EpiTile& t, uint32_t xbuf, uint64_t* acc_full,
                                                  uint32_t acc_parity, int c_begin, int c_end, float* st = nullptr) {
  constexpr int NC = 16 / sizeof(TOut);   // columns per lane after the transpose
  constexpr int LPR = 32 / NC;            // lanes per row segment (8 / 4)
  constexpr int RPI = 32 / LPR;           // rows per warp-wide access (4 / 8)
  constexpr int R = 32 / RPI;             // rows per lane (8 / 4)
  constexpr int NV = NC / 4;              // float4 pieces per lane row
  const int lane = threadIdx.x & 31;
  const int sr = lane / LPR, cg = lane % LPR;
  const int dbg = kDbg ? a.debug : 0;
  const int wrow0 = (int)t.grow - lane;             // first logical row of this warp
  const int trow0 = t.r - lane;                     // ... and its row inside the tile
  const int rdiv = a.row_div, oq = (int)a.out_q, orr = (int)a.out_r, ooff0 = (int)a.out_off;
  const int rq = (int)a.res_q, rr_ = (int)a.res_r, roff0 = (int)a.res_off;
  TOut* outp = reinterpret_cast<TOut*>(a.out) + (long long)t.g * a.out_g + t.n0 + cg * NC;
  const float* resp = HAS_RES ? reinterpret_cast<const float*>(a.res) + (long long)t.g * a.res_g + t.n0 + cg * NC : nullptr;
  const float* biasp = a.bias ? a.bias + (long long)t.g * a.n_pad + t.n0 + cg * NC : nullptr;
  const float* scalep = a.colscale ? a.colscale + t.n0 + cg * NC : nullptr;
  int ooff[R], roff[R];
  uint32_t rd[R], vmask = 0;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int row = i * RPI + sr;
    const int grow = wrow0 + row;
    const bool ok = (trow0 + row < a.rows_valid) && (grow < a.M_total);
    // a.fast guarantees that rows and element offsets fit 32 bits: no 64-bit divisions / multiplies per tile
    const int q = rdiv == 1 ? grow : (int)((unsigned)grow / (unsigned)rdiv), rem = grow - q * rdiv;
    ooff[i] = ok ? (q * oq + rem * orr + ooff0) * a.ldc : 0;
    roff[i] = (ok && HAS_RES) ? (q * rq + rem * rr_ + roff0) * a.ldres : 0;
    vmask |= ok ? (1u << i) : 0u;
    rd[i] = xbuf + row * 128 + (((NV * cg) ^ (row & 7)) << 4);
  }
  const uint32_t wr = xbuf + lane * 128;
  const int sw = lane & 7;
  float4 rb[HAS_RES ? R * NV : 1];
  auto fetch = [&](int c) {
    if constexpr (HAS_RES) {
#pragma unroll
      for (int i = 0; i < R; ++i) {
#pragma unroll
        for (int h = 0; h < NV; ++h)   // rows that do not exist read row 0 of the tile's column range (valid memory), never stored
          rb[i * NV + h] = *reinterpret_cast<const float4*>(resp + roff[i] + c + 4 * h);
      }
    }
  };
  const bool ts_on = (dbg & 128) && blockIdx.x == 0 && threadIdx.x == 64;
  dbg_stamp(ts_on, t.dbg_n, 1);
  // STATS: the rows are read back from L2 by the LayerNorm warps within one tile time; without the hint 60 % of them have been
  // written back and evicted by then (ncu: +121 MB dram__bytes_read per launch at batch 512)
  const uint64_t st_policy = STATS ? l2_policy_evict_last() : 0ull;
  if constexpr (HAS_RES) {
    // Row `lane` of this warp's 32 rows: ask for its residual columns in L2 now (one bulk prefetch per lane), so that the
    // per-chunk loads below are L2 hits; the first chunk's loads are issued before the accumulator is ready.
    const int grow = wrow0 + lane;
    if (trow0 + lane < a.rows_valid && grow < a.M_total) {
      const int q = rdiv == 1 ? grow : (int)((unsigned)grow / (unsigned)rdiv), rem = grow - q * rdiv;
      bulk_prefetch_l2(reinterpret_cast<const float*>(a.res) + (long long)t.g * a.res_g + (long long)(q * rq + rem * rr_ + roff0) * a.ldres +
                           t.n0 + c_begin, (uint32_t)(c_end - c_begin) * 4u);
    }
  }
  fetch(c_begin);
  mbar_wait(acc_full, acc_parity);
  tc_fence_after();
  ptrace(t.tr, 7);
  dbg_stamp(ts_on, t.dbg_n, 2);
#pragma unroll 1
  for (int c = c_begin; c < c_end; c += 32) {
    uint32_t v[32];
    tmem_ld32(t.taddr + c, v);
    tmem_ld_wait();
    float4 bs[NV], sc[NV];   // requested after the TMEM wait (see below), used after the transposition
#pragma unroll
    for (int h = 0; h < NV; ++h) {
      bs[h] = biasp ? *reinterpret_cast<const float4*>(biasp + c + 4 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
      sc[h] = scalep ? *reinterpret_cast<const float4*>(scalep + c + 4 * h) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
    // tcgen05.wait::ld is a scoreboard wait that also covers global loads issued before it (measured: the wait took a full
    // residual-load latency), so this chunk's residual rows are requested only now; they are L2 hits (prefetch above) and
    // overlap the transposition.
    if (HAS_RES && c != c_begin) fetch(c);
    dbg_stamp(ts_on, t.dbg_n, 3);
    float4 x[R * NV];
    if (dbg & 32) {   // developer knob: no transposition through shared memory (results are wrong, timing only)
#pragma unroll
      for (int i = 0; i < R * NV; ++i)
        x[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) st_shared_v4(wr + ((j ^ sw) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < R; ++i) {
        x[i * NV] = ld_shared_v4f(rd[i]);
        if constexpr (NV == 2) x[i * NV + 1] = ld_shared_v4f(rd[i] ^ 16u);   // piece 2cg+1: same row, the neighbouring 16 bytes
      }
      __syncwarp();
    }
    dbg_stamp(ts_on, t.dbg_n, 4);
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float4 y[NV];
#pragma unroll
      for (int h = 0; h < NV; ++h) {
        const float4 xv = x[i * NV + h];
        float2 z0 = fadd2(make_float2(xv.x, xv.y), make_float2(bs[h].x, bs[h].y));
        float2 z1 = fadd2(make_float2(xv.z, xv.w), make_float2(bs[h].z, bs[h].w));
        if constexpr (ACT == ACT_GELU) {
          if constexpr (PRECISE) {
            z0 = make_float2(gelu_erf(z0.x), gelu_erf(z0.y));
            z1 = make_float2(gelu_erf(z1.x), gelu_erf(z1.y));
          } else {
            z0 = gelu_fast2(z0);
            z1 = gelu_fast2(z1);
          }
        } else if constexpr (ACT == ACT_MISH) {
          z0 = make_float2(PRECISE ? mish_precise(z0.x) : mish_f(z0.x), PRECISE ? mish_precise(z0.y) : mish_f(z0.y));
          z1 = make_float2(PRECISE ? mish_precise(z1.x) : mish_f(z1.x), PRECISE ? mish_precise(z1.y) : mish_f(z1.y));
        }
        float2 y0, y1;
        if constexpr (HAS_RES) {
          const float4 r = rb[i * NV + h];
          y0 = ffma2(z0, make_float2(sc[h].x, sc[h].y), make_float2(r.x, r.y));
          y1 = ffma2(z1, make_float2(sc[h].z, sc[h].w), make_float2(r.z, r.w));
        } else {
          y0 = fmul2(z0, make_float2(sc[h].x, sc[h].y));
          y1 = fmul2(z1, make_float2(sc[h].z, sc[h].w));
        }
        y[h] = make_float4(y0.x, y0.y, y1.x, y1.y);
        if constexpr (STATS) {
          const float2 s2 = fadd2(y0, y1), q2 = ffma2(y0, y0, fmul2(y1, y1));
          st[2 * i] += s2.x + s2.y;
          st[2 * i + 1] += q2.x + q2.y;
        }
      }
      if (dbg & 16) {   // developer knob: keep the math alive, produce no store traffic
        if (y[0].x == 123.456f) outp[ooff[i] + c] = TOut(y[0].y);
      } else if ((vmask >> i) & 1) {
        if constexpr (sizeof(TOut) == 4) {
          if constexpr (STATS) st_global_v4f_hint(reinterpret_cast<float*>(outp + ooff[i] + c), y[0], st_policy);   // read back soon
          else *reinterpret_cast<float4*>(outp + ooff[i] + c) = y[0];
        } else {
          uint4 w;
          w.x = pack_bf16x2(y[0].x, y[0].y);
          w.y = pack_bf16x2(y[0].z, y[0].w);
          w.z = pack_bf16x2(y[NV - 1].x, y[NV - 1].y);
          w.w = pack_bf16x2(y[NV - 1].z, y[NV - 1].w);
          *reinterpret_cast<uint4*>(outp + ooff[i] + c) = w;
        }
      }
    }
    dbg_stamp(ts_on, t.dbg_n, 5);
  }
  dbg_stamp(ts_on, t.dbg_n, 6);
}

// runtime (activation, residual) -> compile-time instantiation of the coalescing epilogue
template <int BN, typename TOut, bool PRECISE>
__device__ __forceinline__ void epilogue_linear_fast(const GemmArgs& a, EpiTile& t, uint32_t xbuf, uint64_t* acc_full,
                                                     uint32_t parity, int c0, int c1) {
  // the three combinations the path uses (host sets a.fast only for these): GELU without residual (fc1), plain (qkv,
  // patch embed, U-Net linear convs), plain + residual (attention out-projection, fc2)
  if (a.act == ACT_GELU) epilogue_linear_t<BN, TOut, PRECISE, ACT_GELU, false>(a, t, xbuf, acc_full, parity, c0, c1);
  else if (a.res != nullptr) epilogue_linear_t<BN, TOut, PRECISE, ACT_NONE, true>(a, t, xbuf, acc_full, parity, c0, c1);
  else epilogue_linear_t<BN, TOut, PRECISE, ACT_NONE, false>(a, t, xbuf, acc_full, parity, c0, c1);
}

// ------------------------------------------------------------------------------------------------------------
// EPI_GN: GroupNorm(8) + Mish + FiLM | residual.  Tile = all T rows of `rows_valid / gn_rows` samples x 128 channels =
// whole groups, so the statistics are tile-local.  Pass 1: per-row partial sums per 32-column chunk, reduced over the
// rows of a sample with warp shuffles (power-of-two T <= 32, or T a multiple of 32) or through shared memory.
// ------------------------------------------------------------------------------------------------------------
// Row partials (s1 = sum, s2 = sum of squares per 32-column chunk) -> totals of the row's sample and GroupNorm group,
// left in every thread of the sample.  Power-of-two T <= 32: segmented warp butterfly; T a multiple of 32: warp butterfly +
// shared memory across the sample's warps; any other T: per-row partials through shared memory.
template <int NCH>
__device__ __forceinline__ void gn_sample_totals(const GemmArgs& a, const EpiTile& t, float (&s1)[NCH], float (&s2)[NCH],
                                                 float2* gn_part, float2* gn_stat, int et, int bar_id) {
  const int T = a.gn_rows;
  const int lane = threadIdx.x & 31;
  if (T <= 32 && (T & (T - 1)) == 0) {
    // samples are aligned groups of T lanes: segmented butterfly, every lane ends with its sample's totals
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      for (int off = T >> 1; off > 0; off >>= 1) {
        s1[ch] += __shfl_xor_sync(0xffffffffu, s1[ch], off);
        s2[ch] += __shfl_xor_sync(0xffffffffu, s2[ch], off);
      }
    }
  } else if ((T & 31) == 0) {
    // a sample spans T/32 whole warps: warp butterfly, then combine the warps through shared memory
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      s1[ch] = warp_sum(s1[ch]);
      s2[ch] = warp_sum(s2[ch]);
    }
    if (lane == 0) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) gn_part[(t.r >> 5) * NCH + ch] = make_float2(s1[ch], s2[ch]);
    }
    named_bar_sync(bar_id, 128);
    const int wpt = T >> 5;
    const int w0 = ((t.r >> 5) / wpt) * wpt;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float t1 = 0.f, t2 = 0.f;
      for (int w = 0; w < wpt; ++w) {
        const float2 p = gn_part[(w0 + w) * NCH + ch];
        t1 += p.x;
        t2 += p.y;
      }
      s1[ch] = t1;
      s2[ch] = t2;
    }
  } else {
    // generic T (e.g. 48, 24, 12): per-row partials through shared memory; thread et sums (sample et / NCH, chunk et % NCH)
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) gn_part[t.r * NCH + ch] = make_float2(s1[ch], s2[ch]);
    named_bar_sync(bar_id, 128);
    {
      const int smp = et / NCH, ch = et % NCH;
      const int nsamp = a.rows_valid / T;
      if (smp < nsamp) {
        float t1 = 0.f, t2 = 0.f;
        const int r0 = smp * T;
        for (int i = 0; i < T; ++i) {
          const float2 p = gn_part[(r0 + i) * NCH + ch];
          t1 += p.x;
          t2 += p.y;
        }
        gn_stat[smp * NCH + ch] = make_float2(t1, t2);
      }
    }
    named_bar_sync(bar_id, 128);
    const int smp = t.valid ? (t.r / T) : 0;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const float2 p = gn_stat[smp * NCH + ch];
      s1[ch] = p.x;
      s2[ch] = p.y;
    }
  }
  if (a.gn_gs_log2 == 6) {  // 64-channel groups: merge chunk pairs
#pragma unroll
    for (int ch = 0; ch < NCH; ch += 2) {
      const float p1 = s1[ch] + s1[ch + 1], p2 = s2[ch] + s2[ch + 1];
      s1[ch] = s1[ch + 1] = p1;
      s2[ch] = s2[ch + 1] = p2;
    }
  }
}

// Mish on two lanes: x tanh(softplus(x)) = x n / (n + 2) = x - 2 x / (n + 2) with n = e^x (e^x + 2); two MUFU (ex2, rcp) per
// element, five packed f32x2 operations per pair.  No clamp is needed in this form: e^x = inf gives 1 / inf = 0 and the result x,
// e^x = 0 gives x - x = 0; for very negative x the cancellation leaves an absolute error of ~|x| 2^-23, far below a bf16 ulp of
// anything the value is added to.
__device__ __forceinline__ float2 mish2(float2 x) {
  const float2 xe = fmul2(x, make_float2(1.4426950408889634f, 1.4426950408889634f));
  const float2 e = make_float2(ex2_approx(xe.x), ex2_approx(xe.y));
  const float2 d = ffma2(e, fadd2(e, make_float2(2.f, 2.f)), make_float2(2.f, 2.f));   // n + 2
  float2 r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
  return ffma2(fmul2(x, make_float2(-2.f, -2.f)), r, x);
}

// ------------------------------------------------------------------------------------------------------------
// EPI_GN, bf16 production variant.  Everything that does not depend on the accumulator is fetched while the main loop
// still runs: the per-(sample, column) FiLM scale / shift (cond part + time part) is staged in shared memory by the
// caller (`films`, [samples][2][BN], null = no FiLM or too many samples) and the residual rows of the first chunk are
// already in registers when the accumulator arrives; the residual of chunk c+1 is in flight while chunk c is computed.
// ------------------------------------------------------------------------------------------------------------
template <int BN>
__device__ __forceinline__ void epilogue_gn_fast(const GemmArgs& a, EpiTile& t, const float* colv, const float* films,
                                                 float2* gn_part, float2* gn_stat, int et, int bar_id, uint64_t* acc_full,
                                                 uint32_t acc_parity, int cb) {
  constexpr int NCH = BN / 64;  // 32-column chunks per half tile
  using bf = __nv_bfloat16;
  bf* outp = reinterpret_cast<bf*>(a.out) + (long long)t.g * a.out_g +
             ((long long)t.q * a.out_q + (long long)t.rem * a.out_r + a.out_off) * a.ldc + t.n0 + cb;
  const long long res_row = (long long)t.q * a.res_q + (long long)t.rem * a.res_r + a.res_off;
  const bf* resp = (a.res && t.valid) ? reinterpret_cast<const bf*>(a.res) + (long long)t.g * a.res_g + res_row * a.ldres + t.n0 + cb : nullptr;
  const float* filmp = (a.film_c && !films) ? a.film_c + (long long)t.g * a.film_g + (long long)t.q * a.film_ld + a.film_off + t.n0 + cb : nullptr;
  const float* fs = films ? films + (t.valid ? (t.r / a.gn_rows) : 0) * 2 * BN + cb : nullptr;   // this row's sample
  float* rawp = (a.raw && t.valid) ? a.raw + (long long)t.g * a.raw_g + ((long long)t.q * a.row_div + t.rem) * a.raw_ld + t.n0 + cb : nullptr;
  // 32-byte accesses when rows and chunk offsets allow (chunks are 64 bytes apart): half the L1 wavefronts of the row-per-lane pattern
  const bool out32 = ((reinterpret_cast<uintptr_t>(outp) | (uintptr_t)((long long)a.ldc * 2)) & 31) == 0;
  const bool res32 = resp && ((reinterpret_cast<uintptr_t>(resp) | (uintptr_t)((long long)a.ldres * 2)) & 31) == 0;
  uint4 rr[4];
  auto fetch_res = [&](int ch) {
    if (res32) {
      ld_global_v8(resp + ch * 32, rr[0], rr[1]);
      ld_global_v8(resp + ch * 32 + 16, rr[2], rr[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) rr[i] = resp ? *reinterpret_cast<const uint4*>(resp + ch * 32 + i * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  fetch_res(0);
  const bool ts_on = kDbg && (a.debug & 128) && blockIdx.x == 0 && threadIdx.x == 64;
  dbg_stamp(ts_on, t.dbg_n, 20);

  mbar_wait(acc_full, acc_parity);
  tc_fence_after();
  ptrace(t.tr, 7);
  dbg_stamp(ts_on, t.dbg_n, 21);
  // ---- pass 1: per-row sums of (acc + bias) over every 32-column chunk ----
  float s1[NCH], s2[NCH];
  static_assert(NCH % 2 == 0, "pass 1 reads two chunks per TMEM round trip");
#pragma unroll
  for (int c2 = 0; c2 < NCH; c2 += 2) {
    uint32_t va[32], vb[32];                 // both loads in flight: one exposed TMEM latency per 64 columns instead of two
    tmem_ld32(t.taddr + cb + c2 * 32, va);
    tmem_ld32(t.taddr + cb + c2 * 32 + 32, vb);
    tmem_ld_wait();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int ch = c2 + u;
      const uint32_t(&v)[32] = u ? vb : va;
      float2 a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(colv + cb + ch * 32 + j);
        const float2 x0 = fadd2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), make_float2(b4.x, b4.y));
        const float2 x1 = fadd2(make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), make_float2(b4.z, b4.w));
        a1 = fadd2(a1, fadd2(x0, x1));
        a2 = ffma2(x0, x0, a2);
        a2 = ffma2(x1, x1, a2);
      }
      s1[ch] = t.valid ? a1.x + a1.y : 0.f;
      s2[ch] = t.valid ? a2.x + a2.y : 0.f;
    }
  }
  dbg_stamp(ts_on, t.dbg_n, 22);
  gn_sample_totals<NCH>(a, t, s1, s2, gn_part, gn_stat, et, bar_id);
  dbg_stamp(ts_on, t.dbg_n, 23);
  ptrace(t.tr, 8);
  const float cnt = (float)(a.gn_rows << a.gn_gs_log2);
  float rstd[NCH], nmr[NCH];   // 1/std and -mean/std of the row's sample, per chunk (group)
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const float mean = s1[ch] / cnt;
    rstd[ch] = rsqrtf(fmaxf(s2[ch] / cnt - mean * mean, 0.f) + a.gn_eps);
    nmr[ch] = -mean * rstd[ch];
  }
  // ---- pass 2: normalise, affine, Mish, FiLM, residual, store ----
  // The accumulator chunk is only read by the first phase (v -> y), so the NEXT chunk's tcgen05.ld is issued right after it and its
  // latency (~400 cycles, and the scoreboard wait for the stores that still read the previous values) hides under the Mish phase.
  uint32_t v[32];
  tmem_ld32(t.taddr + cb, v);
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    float rs = rstd[0], nm = nmr[0];
#pragma unroll
    for (int k = 1; k < NCH; ++k) {
      rs = (ch == k) ? rstd[k] : rs;
      nm = (ch == k) ? nmr[k] : nm;
    }
    const float2 rs2 = make_float2(rs, rs), nm2 = make_float2(nm, nm);
    uint4 rc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) rc[i] = rr[i];
    tmem_ld_wait();
    if (ch + 1 < NCH) fetch_res(ch + 1);   // after the wait: tcgen05.wait::ld also waits for global loads issued before it
    const int c0 = ch * 32;
    float2 y[16];
    {
      // The whole 32-column chunk moves through the phases together (16 independent pairs per phase): the Mish chain
      // (ex2 -> fma -> add -> rcp -> mul) has ~150 cycles of latency and only two warps share a scheduler.
      uint4 xprev = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        const float4 b4 = *reinterpret_cast<const float4*>(colv + cb + c0 + 4 * h);
        const float4 g4 = *reinterpret_cast<const float4*>(colv + BN + cb + c0 + 4 * h);
        const float4 e4 = *reinterpret_cast<const float4*>(colv + 2 * BN + cb + c0 + 4 * h);
        const float2 x0 = fadd2(make_float2(__uint_as_float(v[4 * h]), __uint_as_float(v[4 * h + 1])), make_float2(b4.x, b4.y));
        const float2 x1 = fadd2(make_float2(__uint_as_float(v[4 * h + 2]), __uint_as_float(v[4 * h + 3])), make_float2(b4.z, b4.w));
        if (rawp) {   // training: conv + bias as GroupNorm sees it, 32 bytes per store (rows are 32-byte aligned, checked on the host)
          if (h & 1) st_global_v8(rawp + c0 + 4 * (h - 1), xprev, make_uint4(__float_as_uint(x0.x), __float_as_uint(x0.y), __float_as_uint(x1.x), __float_as_uint(x1.y)));
          else xprev = make_uint4(__float_as_uint(x0.x), __float_as_uint(x0.y), __float_as_uint(x1.x), __float_as_uint(x1.y));
        }
        y[2 * h] = ffma2(ffma2(x0, rs2, nm2), make_float2(g4.x, g4.y), make_float2(e4.x, e4.y));
        y[2 * h + 1] = ffma2(ffma2(x1, rs2, nm2), make_float2(g4.z, g4.w), make_float2(e4.z, e4.w));
      }
    }
    if (ch + 1 < NCH) tmem_ld32(t.taddr + cb + (ch + 1) * 32, v);
    if (t.valid) {
#pragma unroll
      for (int p = 0; p < 16; ++p) y[p] = mish2(y[p]);
      if (fs) {
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const float4 sc = *reinterpret_cast<const float4*>(fs + c0 + 4 * h);
          const float4 sh = *reinterpret_cast<const float4*>(fs + BN + c0 + 4 * h);
          y[2 * h] = ffma2(y[2 * h], make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
          y[2 * h + 1] = ffma2(y[2 * h + 1], make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
        }
      } else if (filmp) {   // more samples per tile than the staging buffer holds: straight from global memory
#pragma unroll
        for (int j8 = 0; j8 < 32; j8 += 8) {
          float sc[8], sh[8];
          load_res8(filmp + c0 + j8, sc);
          load_res8(filmp + a.film_C + c0 + j8, sh);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 s2v = fadd2(make_float2(sc[2 * h], sc[2 * h + 1]),
                                     make_float2(colv[3 * BN + cb + c0 + j8 + 2 * h], colv[3 * BN + cb + c0 + j8 + 2 * h + 1]));
            const float2 h2v = fadd2(make_float2(sh[2 * h], sh[2 * h + 1]),
                                     make_float2(colv[4 * BN + cb + c0 + j8 + 2 * h], colv[4 * BN + cb + c0 + j8 + 2 * h + 1]));
            y[j8 / 2 + h] = ffma2(y[j8 / 2 + h], s2v, h2v);
          }
        }
      }
      if (resp) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 rv = rc[q];
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
          for (int h = 0; h < 4; ++h) y[4 * q + h] = fadd2(y[4 * q + h], __bfloat1622float2(h2[h]));
        }
      }
      uint4 w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        w[q].x = pack_bf16x2(y[4 * q].x, y[4 * q].y);
        w[q].y = pack_bf16x2(y[4 * q + 1].x, y[4 * q + 1].y);
        w[q].z = pack_bf16x2(y[4 * q + 2].x, y[4 * q + 2].y);
        w[q].w = pack_bf16x2(y[4 * q + 3].x, y[4 * q + 3].y);
      }
      if (kDbg && (a.debug & 16)) {   // developer knob: no store traffic (timing only)
        if (w[0].x == 0x12345678u) *reinterpret_cast<uint4*>(outp + c0) = w[0];
      } else if (out32) {
        st_global_v8(outp + c0, w[0], w[1]);
        st_global_v8(outp + c0 + 16, w[2], w[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(outp + c0 + 8 * q) = w[q];
      }
    }
    dbg_stamp(ts_on, t.dbg_n, 24);
  }
}

template <int BN, typename TOut, bool PRECISE>
__device__ __forceinline__ void epilogue_gn(const GemmArgs& a, const EpiTile& t, const float* colv, float2* gn_part,
                                            float2* gn_stat, int et, int bar_id, uint64_t* acc_full, uint32_t acc_parity,
                                            int cb) {
  // `cb`: first tile column of the half this warpgroup handles (0 or BN / 2); colv / TMEM / output columns are tile-relative
  const long long oplane = a.out_plane, rplane = a.res_plane;
  const int filmC = a.film_C;

  static_assert(BN == 128 || BN == 256, "GN epilogue: whole 32/64-channel groups per tile");
  constexpr int NCH = BN / 64;  // 32-column chunks per half tile (2 or 4): whole 32- / 64-channel groups
  TOut* outp = reinterpret_cast<TOut*>(a.out) + (long long)t.g * a.out_g +
               ((long long)t.q * a.out_q + (long long)t.rem * a.out_r + a.out_off) * a.ldc + t.n0 + cb;
  const long long res_row = (long long)t.q * a.res_q + (long long)t.rem * a.res_r + a.res_off;
  const TOut* resp = a.res ? reinterpret_cast<const TOut*>(a.res) + (long long)t.g * a.res_g + res_row * a.ldres + t.n0 + cb : nullptr;
  const float* filmp = a.film_c ? a.film_c + (long long)t.g * a.film_g + (long long)t.q * a.film_ld + a.film_off + t.n0 + cb : nullptr;

  mbar_wait(acc_full, acc_parity);
  tc_fence_after();
  ptrace(t.tr, 7);
  float s1[NCH], s2[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    uint32_t v[32];
    tmem_ld32(t.taddr + cb + ch * 32, v);
    tmem_ld_wait();
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float x = __uint_as_float(v[j]) + colv[cb + ch * 32 + j];
      a1 += x;
      a2 = fmaf(x, x, a2);
    }
    s1[ch] = t.valid ? a1 : 0.f;
    s2[ch] = t.valid ? a2 : 0.f;
  }
  gn_sample_totals<NCH>(a, t, s1, s2, gn_part, gn_stat, et, bar_id);
  const float cnt = (float)(a.gn_rows << a.gn_gs_log2);
  float mean[NCH], rstd[NCH];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    mean[ch] = s1[ch] / cnt;
    rstd[ch] = rsqrtf(fmaxf(s2[ch] / cnt - mean[ch] * mean[ch], 0.f) + a.gn_eps);
  }

#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {   // rolled: a fully unrolled body (NCH x 4 copies) thrashes the instruction cache
    uint32_t v[32];
    tmem_ld32(t.taddr + cb + ch * 32, v);
    float mu = mean[0], rs = rstd[0];
#pragma unroll
    for (int k = 1; k < NCH; ++k) {
      mu = (ch == k) ? mean[k] : mu;
      rs = (ch == k) ? rstd[k] : rs;
    }
    tmem_ld_wait();
    if (t.valid) {
#pragma unroll
      for (int j8 = 0; j8 < 32; j8 += 8) {
        const int cc = ch * 32 + j8;
        float sc[8], sh[8], rr[8], rl[8];
        if (filmp) {
          load_res8(filmp + cc, sc);
          load_res8(filmp + filmC + cc, sh);
        }
        if (resp) {
          load_res8(resp + cc, rr);
          if (rplane > 0) load_res8(resp + rplane + cc, rl);
        }
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float x = __uint_as_float(v[j8 + j]) + colv[cb + cc + j];
          x = (x - mu) * rs * colv[BN + cb + cc + j] + colv[2 * BN + cb + cc + j];
          y[j] = PRECISE ? mish_precise(x) : mish_f(x);
        }
        if (filmp) {
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = (sc[j] + colv[3 * BN + cb + cc + j]) * y[j] + (sh[j] + colv[4 * BN + cb + cc + j]);
        }
        if (resp) {
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] += rr[j];
          if (rplane > 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] += rl[j];
          }
        }
        store_split8<TOut>(outp + cc, oplane, y);
      }
    }
  }
}

// FiLM scale / shift of every sample of a GroupNorm tile (per-sample cond part + per-step time part) -> shared memory
// [samples][2][BN].  128-bit loads; every thread first issues ALL its global loads (up to 2 x 4 float4), then adds and stores: a
// plain load / store loop lets the compiler assume the shared-memory stores alias the following loads, which serialises one L2
// round trip per element (measured: the top stall reason of the epilogue warps, 16 dependent round trips per tile).
template <int BN>
__device__ __forceinline__ void stage_film(const GemmArgs& a, float* films, const float* ft, int g, int n0, int m_tile, int nsamp,
                                           int tid, int nthreads) {
  const long long smp0 = (long long)m_tile * nsamp;
  const long long n_samples = a.M_total / a.gn_rows;
  const float* fc = a.film_c + (long long)g * a.film_g + a.film_off + n0;
  const int n4 = nsamp * 2 * BN / 4;
  constexpr int MAXV = GEMM_FILM_SAMPLES * 2 * BN / 4 / 256;   // float4 pieces per thread with 256 staging threads
  float4 vc[MAXV], vt[MAXV];
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i = (tid + k * nthreads) * 4;
    vc[k] = vt[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid + k * nthreads < n4) {
      const int c = i % BN, which = (i / BN) & 1, smp = i / (2 * BN);
      if (smp0 + smp < n_samples) {
        vc[k] = __ldg(reinterpret_cast<const float4*>(fc + (smp0 + smp) * a.film_ld + which * a.film_C + c));
        if (ft) vt[k] = __ldg(reinterpret_cast<const float4*>(ft + which * a.film_C + c));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i4 = tid + k * nthreads;
    if (i4 < n4)
      *reinterpret_cast<float4*>(films + i4 * 4) = make_float4(vc[k].x + vt[k].x, vc[k].y + vt[k].y, vc[k].z + vt[k].z, vc[k].w + vt[k].w);
  }
  for (int i4 = tid + MAXV * nthreads; i4 < n4; i4 += nthreads) {   // (fewer than 256 staging threads)
    const int i = i4 * 4;
    const int c = i % BN, which = (i / BN) & 1, smp = i / (2 * BN);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (smp0 + smp < n_samples) {
      v = __ldg(reinterpret_cast<const float4*>(fc + (smp0 + smp) * a.film_ld + which * a.film_C + c));
      if (ft) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(ft + which * a.film_C + c));
        v = make_float4(v.x + w.x, v.y + w.y, v.z + w.z, v.w + w.w);
      }
    }
    *reinterpret_cast<float4*>(films + i4 * 4) = v;
  }
}

template <typename TIn, int BN, int MODE, typename TOut, int STAGES, bool PRECISE, int KA, int CTAS, int EW>
__global__ void __launch_bounds__(GEMM_THREADS(EW), 1) gemm_tc_kernel(const __grid_constant__ GemmArgs a) {
  static_assert(EW == 8 || (EW == 12 && MODE == EPI_LINEAR), "epilogue warps: 8, or 12 for the LINEAR epilogue");
  constexpr int KE = InTraits<TIn>::KE;
  constexpr int A_ATOM_BYTES = GEMM_A_STAGE_BYTES;   // one K atom of A: 128 rows x 128 B
  constexpr int B_ROWS = BN / CTAS;                  // rows of B this CTA holds (a pair splits the tile's N)
  constexpr int B_ATOM_BYTES = B_ROWS * 128;
  constexpr int A_STAGE_BYTES = KA * A_ATOM_BYTES;
  constexpr int B_STAGE_BYTES = KA * B_ATOM_BYTES;
  constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 256 ? 256 : 512);  // power of two
  constexpr uint32_t IDESC = umma_idesc(InTraits<TIn>::FMT, BN, 0, 0, 128 * CTAS);
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 192 || BN == 256, "BN");
  static_assert(MODE != EPI_GN || BN == 128 || BN == 256, "GN epilogue: whole groups per tile");
  static_assert(CTAS == 1 || (CTAS == 2 && sizeof(TIn) == 2 && BN % 32 == 0), "CTA pairs: bf16 operands");

  extern __shared__ uint8_t smem_raw[];
  const bool ts_mma = kDbg && (a.debug & 256) && blockIdx.x == 0 && threadIdx.x == 32;   // developer instrumentation: MMA thread of CTA 0
  int ts_n = 0;
  dbg_stamp(ts_mma, ts_n, 30);
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint8_t* sX = sB + STAGES * B_STAGE_BYTES;      // per-warp transposition buffers of the coalescing epilogue
  uint64_t* full = reinterpret_cast<uint64_t*>(sX + GEMM_XPOSE_BYTES(BN, MODE, EW));
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = a.passes * a.taps * a.cblocks;
  // A pair (cluster of two CTAs, CTAS == 2) works on two vertically adjacent 128-row tiles as ONE M = 256 MMA:
  // rank 0 (the leader) issues the MMAs and owns the `full` barriers, both CTAs load their own operand halves.
  const int rank = CTAS == 2 ? (int)cluster_ctarank() : 0;
  const int worker = blockIdx.x / CTAS, n_workers = gridDim.x / CTAS;

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
      for (int s = 0; s < 2; ++s) {
        mbar_init(&acc_full[s], 1);
        mbar_init(&acc_empty[s], EW * CTAS);   // one arrival per epilogue warp of every CTA of the pair
      }
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (CTAS == 2) {
      tmem_alloc_pair(tmem_slot, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  pdl_launch_dependents();   // the next kernel may set itself up on SMs this grid has already left
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  dbg_stamp(ts_mma, ts_n, 31);
  // Barriers, TMEM and descriptors are ready; from here on the producing kernel must have finished (programmatic
  // dependent launch).  Only the producer thread runs ahead: the WEIGHT tiles of its first stages do not depend on the
  // predecessor, so their HBM round trip is started before the wait.
  int pre_atoms = 0;
  if (warp == 0 && lane == 0 && a.passes == 1 && !(kDbg && (a.debug & 5)) && worker < a.total_tiles) {
    const int per_pass = a.taps * a.cblocks;          // == nk
    const int pre_stages = (per_pass + KA - 1) / KA < STAGES ? (per_pass + KA - 1) / KA : STAGES;
    const int n_tile0 = worker % a.n_tiles, g0 = (worker / a.n_tiles) / a.m_tiles;
    const int g_b0 = g0 * a.n_pad + n_tile0 * BN + rank * B_ROWS;
    for (int st = 0; st < pre_stages; ++st) {
      const int left = per_pass - st * KA;
      const int n_at = left < KA ? left : KA;
      if (rank == 0) mbar_arrive_expect_tx(&full[st], n_at * CTAS * (a.a_box_bytes + B_ATOM_BYTES));   // ring is empty: no wait
      for (int at = 0; at < n_at; ++at) {
        const int kb = (st * KA + at) * KE;
        if constexpr (CTAS == 2)
          tma_load_2d_pair(sB + st * B_STAGE_BYTES + at * B_ATOM_BYTES, &a.tmB, mapa_shared(smem_u32(&full[st]), 0), kb, g_b0);
        else
          tma_load_2d(sB + st * B_STAGE_BYTES + at * B_ATOM_BYTES, &a.tmB, &full[st], kb, g_b0);
        ++pre_atoms;
      }
    }
  }
  pdl_wait();
  dbg_stamp(ts_mma, ts_n, 32);

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;   // ring position kept incrementally: no divisions in this loop
      int n_tile = worker % a.n_tiles, rest = worker / a.n_tiles;
      const int dn = n_workers % a.n_tiles, dr = n_workers / a.n_tiles;
      const bool no_tma = kDbg && (a.debug & 5) != 0;
      int at = 0, n_at = 0;   // atoms issued into / planned for the current stage
      int atom_seq = 0;       // atoms handled so far (the first pre_atoms already have their stage opened and B in flight)
      for (int tile = worker; tile < a.total_tiles; tile += n_workers) {
        int left = nk;
        const int m_tile = (rest % a.m_tiles) * CTAS + rank;
        const int g = rest / a.m_tiles;
        const int n0 = n_tile * BN;
        const int t_base = m_tile * a.m_t_step, b_base = m_tile * a.m_b_step, g_a = g * a.a_g_mul;
        const int g_b = g * a.n_pad + n0 + rank * B_ROWS;
        for (int pass = 0; pass < a.passes; ++pass) {
          const int pa = a.a_c0 + ((pass == 1) ? a.a_plane : 0);
          int kb = (pass == 2) ? a.b_plane : 0;
          for (int tp = 0; tp < a.taps; ++tp) {
            const int tap_p = a.tap_p[tp], tap_t = t_base + a.tap_t[tp];
            for (int cb = 0; cb < a.cblocks; ++cb, kb += KE) {
              const bool pre = atom_seq < pre_atoms;   // stage opened and weight tile requested before pdl_wait
              ++atom_seq;
              if (at == 0) {   // open a stage: it will receive min(KA, k-blocks left in this tile) atoms
                n_at = left < KA ? left : KA;
                if (!pre) {
                  mbar_wait(&empty[s], ph ^ 1);
                  if (rank == 0) {   // the leader's barrier counts the bytes of both CTAs
                    if (no_tma) mbar_arrive(&full[s]);
                    else mbar_arrive_expect_tx(&full[s], n_at * CTAS * (a.a_box_bytes + B_ATOM_BYTES));
                  }
                }
              }
              if (!no_tma) {
                if constexpr (CTAS == 2) {
                  const uint32_t fb = mapa_shared(smem_u32(&full[s]), 0);
                  tma_load_5d_pair(sA + s * A_STAGE_BYTES + at * A_ATOM_BYTES, &a.tmA, fb, pa + cb * KE, tap_p, tap_t, b_base, g_a);
                  if (!pre) tma_load_2d_pair(sB + s * B_STAGE_BYTES + at * B_ATOM_BYTES, &a.tmB, fb, kb, g_b);
                } else {
                  tma_load_5d(sA + s * A_STAGE_BYTES + at * A_ATOM_BYTES, &a.tmA, &full[s], pa + cb * KE, tap_p, tap_t, b_base, g_a);
                  if (!pre) tma_load_2d(sB + s * B_STAGE_BYTES + at * B_ATOM_BYTES, &a.tmB, &full[s], kb, g_b);
                }
              }
              --left;
              if (++at == n_at) {
                at = 0;
                if (++s == STAGES) {
                  s = 0;
                  ph ^= 1;
                }
              }
            }
          }
        }
        n_tile += dn;
        rest += dr;
        if (n_tile >= a.n_tiles) {
          n_tile -= a.n_tiles;
          ++rest;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ UMMA issuer (leader CTA only) ------------------------------
    if (lane == 0 && rank == 0) {
      uint32_t lt = 0, ph = 0;
      int s = 0;
      const bool no_mma = kDbg && (a.debug & 9) != 0;
      for (int tile = worker; tile < a.total_tiles; tile += n_workers, ++lt) {
        const uint32_t acc = lt & 1;
        mbar_wait(&acc_empty[acc], ((lt >> 1) & 1) ^ 1);  // the epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int i = 0; i < nk; i += KA) {
          const int n_at = (nk - i) < KA ? (nk - i) : KA;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (lt == 0 && i == 0) dbg_stamp(ts_mma, ts_n, 33);
#pragma unroll
          for (int at = 0; at < KA; ++at) {
            if (at >= n_at || no_mma) break;
            const uint64_t adesc = umma_smem_desc_sw128(smem_u32(sA + s * A_STAGE_BYTES + at * A_ATOM_BYTES));
            const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(sB + s * B_STAGE_BYTES + at * B_ATOM_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // 4 x (32 bytes of K) per 128-byte swizzle row
              if constexpr (CTAS == 2)
                umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (i | at | k) != 0);
              else if constexpr (sizeof(TIn) == 2)
                umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (i | at | k) != 0);
              else
                umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (i | at | k) != 0);
            }
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
          if constexpr (CTAS == 2) umma_commit_pair(&empty[s], 3); else umma_commit(&empty[s]);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        if constexpr (CTAS == 2) umma_commit_pair(&acc_full[acc], 3); else umma_commit(&acc_full[acc]);
        dbg_stamp(ts_mma, ts_n, 34);
      }
    }
  } else {
    // ------------------------------ epilogue warpgroups ------------------------------
    // Both warpgroups work on EVERY tile: warpgroup h takes the tile's columns [h * BN/2, (h+1) * BN/2), so a CTA with
    // one or two tiles (the U-Net layers at batch 256) still uses all eight epilogue warps; accumulator lt & 1.
    const int half = (warp - 2) >> 2;
    const int et = (threadIdx.x - 64) & 127;      // thread index inside the warpgroup
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    // 32-column chunks of the tile are dealt to the EW / 4 warpgroups (a warpgroup without columns only signals)
    constexpr int NPARTS = EW / 4, CHUNKS = (BN + 31) / 32;
    const int my_chunks = CHUNKS / NPARTS + (half < CHUNKS % NPARTS ? 1 : 0);
    const int c_begin = 32 * (half * (CHUNKS / NPARTS) + (half < CHUNKS % NPARTS ? half : CHUNKS % NPARTS));
    const int c_end = c_begin + 32 * my_chunks;
    constexpr int CV = GEMM_COLV_FLOATS(BN, MODE);
    constexpr bool GN_FAST = MODE == EPI_GN && sizeof(TOut) == 2 && !PRECISE;
    float2* gn_part = reinterpret_cast<float2*>(scratch + CV) + half * (128 + 64) * 4;   // [128][4]   (GroupNorm only)
    float2* gn_stat = gn_part + 128 * 4;                                                  // [64][4]
    float* films = scratch + CV + 2 * (128 + 64) * 4 * 2;                                 // [8][2][BN] (GroupNorm only)
    EpiTile t;
    t.r = quarter * 32 + lane;
    t.dbg_n = 0;
    t.tr = nullptr;
    uint32_t lt = 0;
    const bool fast = GEMM_XPOSE_BYTES(BN, MODE, EW) > 0 && a.fast != 0;
    const int et256 = threadIdx.x - 64;
    for (int tile = worker; tile < a.total_tiles; tile += n_workers, ++lt) {
      const uint32_t acc = lt & 1;
      dbg_stamp(kDbg && (a.debug & 128) && blockIdx.x == 0 && threadIdx.x == 64, t.dbg_n, 19);
      const int n_tile = tile % a.n_tiles;
      const int rest = tile / a.n_tiles;
      const int m_tile = (rest % a.m_tiles) * CTAS + rank;
      t.g = rest / a.m_tiles;
      t.n0 = n_tile * BN;
      float* colv = scratch + (MODE == EPI_LINEAR ? acc * CV : 0);   // column vectors of this tile
      bool film_staged = false;
      if (!fast) {
        // Stage the per-column vectors.  LINEAR: set `acc` was last read two tiles ago and every warp has passed the
        // barrier of the tile in between, so only the write -> read edge needs a barrier.  GroupNorm (one set): all
        // readers of the previous tile must be done first.
        if (MODE == EPI_GN) named_bar_sync(1, 32 * EW);
        // (all global loads of a thread are issued before its first shared-memory store, see stage_film)
        const long long gcol = (long long)t.g * a.n_pad + t.n0;
        for (int c = et256; c < BN; c += 32 * EW) {
          if (MODE == EPI_LINEAR) {
            const float b_ = a.bias ? __ldg(a.bias + gcol + c) : 0.f;
            const float s_ = (a.colscale && (t.n0 + c) < a.N) ? __ldg(a.colscale + t.n0 + c) : 1.f;
            colv[c] = b_;
            colv[BN + c] = s_;
          } else {
            const bool f = a.film_t != nullptr;
            const long long fo = (long long)t.g * a.film_tg + a.film_off + t.n0 + c;
            const float b_ = a.bias ? __ldg(a.bias + gcol + c) : 0.f;
            const float g_ = __ldg(a.gn_gamma + gcol + c), e_ = __ldg(a.gn_beta + gcol + c);
            const float f0 = f ? __ldg(a.film_t + fo) : 0.f, f1 = f ? __ldg(a.film_t + fo + a.film_C) : 0.f;
            colv[c] = b_;
            colv[BN + c] = g_;
            colv[2 * BN + c] = e_;
            colv[3 * BN + c] = f0;
            colv[4 * BN + c] = f1;
          }
        }
        if constexpr (GN_FAST) {
          // FiLM scale / shift of every sample of the tile (cond part + time part), while the main loop runs
          const int nsamp = a.rows_valid / a.gn_rows;
          if (a.film_c && nsamp <= GEMM_FILM_SAMPLES) {
            film_staged = true;
            stage_film<BN>(a, films, a.film_t ? a.film_t + (long long)t.g * a.film_tg + a.film_off + t.n0 : nullptr, t.g, t.n0, m_tile, nsamp,
                           et256, 32 * EW);
          }
        }
        named_bar_sync(1, 32 * EW);
      }
      t.grow = (long long)m_tile * a.rows_valid + t.r;
      t.valid = (t.r < a.rows_valid) && (t.grow < a.M_total);
      t.q = a.row_div == 1 ? (int)t.grow : (int)((unsigned)t.grow / (unsigned)a.row_div);   // rows fit 32 bits (M is an int)
      t.rem = (int)(t.grow - (long long)t.q * a.row_div);
      t.taddr = tmem_base + acc * ACC_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
      const uint32_t parity = (lt >> 1) & 1;
      if (c_begin < c_end) {
        if (kDbg && (a.debug & 2)) {
          mbar_wait(&acc_full[acc], parity);
          tc_fence_after();
        } else if constexpr (MODE == EPI_LINEAR) {
          if constexpr (GEMM_XPOSE_BYTES(BN, MODE, EW) > 0) {
            if (fast) epilogue_linear_fast<BN, TOut, PRECISE>(a, t, smem_u32(sX) + (warp - 2) * 4096, &acc_full[acc], parity, c_begin, c_end);
            else epilogue_linear<BN, TOut, PRECISE>(a, t, colv, &acc_full[acc], parity, c_begin, c_end);
          } else {
            epilogue_linear<BN, TOut, PRECISE>(a, t, colv, &acc_full[acc], parity, c_begin, c_end);
          }
        } else if constexpr (GN_FAST) {
          if (a.out_plane == 0 && a.res_plane == 0)
            epilogue_gn_fast<BN>(a, t, colv, film_staged ? films : nullptr, gn_part, gn_stat, et, 2 + half, &acc_full[acc], parity, c_begin);
          else
            epilogue_gn<BN, TOut, PRECISE>(a, t, colv, gn_part, gn_stat, et, 2 + half, &acc_full[acc], parity, c_begin);
        } else {
          epilogue_gn<BN, TOut, PRECISE>(a, t, colv, gn_part, gn_stat, et, 2 + half, &acc_full[acc], parity, c_begin);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {   // this warp's share of the accumulator has been read out
        if constexpr (CTAS == 2) mbar_arrive_remote(mapa_shared(smem_u32(&acc_empty[acc]), 0)); else mbar_arrive(&acc_empty[acc]);
      }
    }
  }

  dbg_stamp(ts_mma, ts_n, 35);
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();   // the peer may still signal / read this CTA
  dbg_stamp(ts_mma, ts_n, 36);
  if (warp == 1) {
    if constexpr (CTAS == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace vt
