// Fused ViT MLP of a DinoV2 block (HF Dinov2MLP + layer_scale2 + residual, HF:312-328,380-386), hidden size D = 384:
//     h += ls2 * ( GELU(xn W1^T + b1) W2^T + b2 ),   xn = LayerNorm2(h) in bf16, W1 [4D, D], W2 [D, 4D]
// as ONE kernel: the [rows, 4D] hidden activation (404 MB per layer at batch 256 x 2 cameras) never leaves the SM.
//
// PERSISTENT CTA pairs (cluster of two, tcgen05 cta_group::2, M = 256 per MMA).  Per pair and 256-row tile:
//   X tile (each CTA its 128 rows x 384, six 64-wide K atoms) stays in shared memory for the whole tile;
//   for each of the twelve 128-wide hidden chunks j:
//     MMA1   H_j = X W1_j^T          (N = 128, K = 384)  -> TMEM columns [384, 512)
//     warps  H_j + b1 -> GELU -> bf16 -> shared memory as the 128B-swizzled K-major A operand of MMA2
//     MMA2   Y  += H_j W2_j^T        (N = 2 x 192, K = 128) -> TMEM columns [0, 384)
//   MMA1 of chunk j+1 is issued as soon as the epilogue warps have pulled H_j out of TMEM, so the tensor pipe runs
//   MMA1(j+1) and MMA2(j) back to back while the warps compute GELU of the next chunk.
//   Y + b2, * ls2, + residual -> h through the coalescing epilogue of the GEMM kernel (vt_gemm.cuh).
// Each CTA stages only HALF of every weight tile (64 of W1_j's 128 rows, 2 x 96 of W2's 384 rows): the 2.36 MB of
// weights stream through a pair once per 256 rows, 28 bytes per clock per SM.
//   warp 0  TMA producer      warp 1  MMA issuer (leader CTA)      warps 4-11  epilogue (quarter = warp & 3, half = (warp-4)/4)
//   warps 2-3  LayerNorm of the finished rows (the next block's norm1), off the critical path.
// 12 warps: every thread keeps the 168-register budget the epilogue needs (no setmaxnreg).
#pragma once
#include "vt_gemm.cuh"

namespace vt {

constexpr int MLP_D = 384, MLP_H = 1536, MLP_CH = 128, MLP_NCH = MLP_H / MLP_CH;   // 12 hidden chunks
constexpr int MLP_THREADS = 384;
constexpr int MLP_LN_W0 = 2, MLP_LN_WARPS = 2, MLP_EPI_W0 = 4;   // first LayerNorm warp, their number, first epilogue warp
constexpr int MLP_X_BYTES = 6 * 16384;          // six K atoms of 128 rows x 128 B
constexpr int MLP_H_BYTES = 2 * 16384;          // H chunk as two K atoms (also the epilogue's transposition scratch)
constexpr int MLP_W1_SLOT = 64 * 128;           // 64 rows x 128 B (this CTA's half of a W1 chunk atom)
constexpr int MLP_W1_SLOTS = 6;
constexpr int MLP_W2_SLOT = 2 * 96 * 128;       // 2 x 96 rows x 128 B (this CTA's half of both 192-column halves)
constexpr int MLP_W2_SLOTS = 2;
constexpr int MLP_SMEM_X = 0;
constexpr int MLP_SMEM_H = MLP_SMEM_X + MLP_X_BYTES;
constexpr int MLP_SMEM_W1 = MLP_SMEM_H + MLP_H_BYTES;
constexpr int MLP_SMEM_W2 = MLP_SMEM_W1 + MLP_W1_SLOTS * MLP_W1_SLOT;
constexpr int MLP_SMEM_BAR = MLP_SMEM_W2 + MLP_W2_SLOTS * MLP_W2_SLOT;
constexpr int MLP_SMEM_BYTES = 1024 + MLP_SMEM_BAR + 512;
static_assert(MLP_SMEM_BYTES <= 227 * 1024, "mlp_fused_kernel shared memory");

struct MlpArgs {
  CUtensorMap tmX;    // 2-D (D, rows) over xn (bf16), box (64, 128)
  CUtensorMap tmW1;   // 2-D (D, 4D) over W1 (bf16, K contiguous), box (64, 64)
  CUtensorMap tmW2;   // 2-D (4D, D) over W2, box (64, 96)
  const float* b1;    // [4D]
  GemmArgs epi;       // output side: out = res = h (fp32, ld D), bias = b2, colscale = ls2, M_total = rows, rows_valid = 128 ...
  int m_tiles;        // 128-row tiles
  int n_pairs;        // ceil(m_tiles / 2) work units
  // optional: LayerNorm of the updated rows (the NEXT block's norm1, HF:367-372) written as bf16 by the CTA that produced them
  const float* ln_gamma;   // [D] or null
  const float* ln_beta;
  __nv_bfloat16* ln_out;   // [rows][ln_ld]
  long long ln_ld;
  float ln_eps;
  int ln_debug;            // developer knobs (VT_DEBUG_KNOBS builds): 1 no stores, 2 no loads, 4 handshake only
};


// ---- LayerNorm of finished 384-wide rows by a warp that is off the critical path (shared with vt_rowproj.cuh) ----
// (sum, sum of squares) of the two column halves of a row -> (rstd, -mean * rstd)
__device__ __forceinline__ float2 ln_finish_stats(float2 p0, float2 p1, float eps) {
  const float mean = (p0.x + p1.x) * (1.0f / 384.0f);
  const float var = fmaxf((p0.y + p1.y) * (1.0f / 384.0f) - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  return make_float2(rstd, -mean * rstd);
}
// NROWS consecutive rows from `row0`.  NROWS = 64: lane l holds the statistics of rows 2l (mr0) and 2l + 1 (mr1); NROWS = 32:
// of row l (mr0).  One warp per row, eight rows in flight from L2, (v * rstd - mean * rstd) * g + b in packed fp32 pairs, bf16
// out: ~35 instructions per row.  HOLD_GB keeps gamma / beta in 24 registers instead of re-reading them from L1.
template <int NROWS, bool HOLD_GB>
__device__ __forceinline__ void ln_warp_rows(const float* hp, long long ldh, long long m_total, long long row0, float2 mr0, float2 mr1,
                                             const float* gamma, const float* beta, __nv_bfloat16* out, long long ldo, int lane,
                                             int ldbg) {
  static_assert(NROWS == 64 || NROWS == 32, "statistics layout");
  const uint64_t pol = l2_policy_evict_first();
  float4 g[3], b[3];
  if constexpr (HOLD_GB) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
      b[i] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
    }
  }
#pragma unroll 1
  for (int r8 = (ldbg & 4) ? 1 << 20 : 0; r8 < NROWS; r8 += 8) {
    float4 v[8][3];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long long row = row0 + r8 + u;
      const float4* xr = reinterpret_cast<const float4*>(hp + row * ldh);
#pragma unroll
      for (int i = 0; i < 3; ++i)   // last use of the fp32 row in this kernel: first in line for eviction
        v[u][i] = (row < m_total && !(ldbg & 2)) ? ld_global_v4f_hint(xr + lane + 32 * i, pol) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long long row = row0 + r8 + u;
      const int owner = NROWS == 64 ? (r8 + u) >> 1 : r8 + u;
      const float rs = __shfl_sync(0xffffffffu, (NROWS == 64 && (u & 1)) ? mr1.x : mr0.x, owner);
      const float nm = __shfl_sync(0xffffffffu, (NROWS == 64 && (u & 1)) ? mr1.y : mr0.y, owner);
      const float2 rs2 = make_float2(rs, rs), nm2 = make_float2(nm, nm);
      if (row < m_total && !((ldbg & 1) && rs != 123.456f)) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if constexpr (!HOLD_GB) {
            g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
            b[i] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
          }
          const float2 t0 = ffma2(make_float2(v[u][i].x, v[u][i].y), rs2, nm2);
          const float2 t1 = ffma2(make_float2(v[u][i].z, v[u][i].w), rs2, nm2);
          const float2 y0 = ffma2(t0, make_float2(g[i].x, g[i].y), make_float2(b[i].x, b[i].y));
          const float2 y1 = ffma2(t1, make_float2(g[i].z, g[i].w), make_float2(b[i].z, b[i].w));
          uint2 pk;
          pk.x = pack_bf16x2(y0.x, y0.y);
          pk.y = pack_bf16x2(y1.x, y1.y);
          *reinterpret_cast<uint2*>(out + row * ldo + 4 * (lane + 32 * i)) = pk;
        }
      }
    }
  }
}
// Epilogue side: st[2i] / st[2i + 1] (sum, sum of squares of this lane's four columns of row i*4 + sr, all chunks) are completed
// over the 8 lanes of a row segment; lanes cg == 0 leave the 32 rows' partials at `dst` (256 B of shared memory).
__device__ __forceinline__ void ln_store_partials(float (&st)[16], uint32_t dst, int lane) {
#pragma unroll
  for (int m = 1; m < 8; m <<= 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) st[i] += __shfl_xor_sync(0xffffffffu, st[i], m);
  }
  if ((lane & 7) == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) st_shared_v2f(dst + (i * 4 + (lane >> 3)) * 8, st[2 * i], st[2 * i + 1]);
  }
}

__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_fused_kernel(const __grid_constant__ MlpArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sX = smem + MLP_SMEM_X;
  uint8_t* sH = smem + MLP_SMEM_H;
  uint8_t* sW1 = smem + MLP_SMEM_W1;
  uint8_t* sW2 = smem + MLP_SMEM_W2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MLP_SMEM_BAR);
  uint64_t* x_full = bars;                 // leader
  uint64_t* x_empty = bars + 1;            // both (multicast commit)
  uint64_t* w1_full = bars + 2;            // [6] leader
  uint64_t* w1_empty = bars + 8;           // [6] both
  uint64_t* w2_full = bars + 14;           // [2] leader
  uint64_t* w2_empty = bars + 16;          // [2] both
  uint64_t* h_full = bars + 18;            // both: MMA1(j) complete
  uint64_t* h_tfree = bars + 19;           // leader: every epilogue warp of the pair has read H_j out of TMEM
  uint64_t* h_ready = bars + 20;           // leader: every epilogue warp of the pair has stored its part of bf16 H_j
  uint64_t* h_sfree = bars + 21;           // both: MMA2(j) has read the shared-memory H_j
  uint64_t* y_full = bars + 22;            // both: MMA2(11) complete
  uint64_t* y_free = bars + 23;            // leader: every epilogue warp of the pair has read Y
  uint64_t* ln_go = bars + 24;             // own CTA: the eight epilogue warps have stored their columns of the tile's rows
  uint64_t* st_free = bars + 25;           // own CTA: the LayerNorm warps have read the row statistics out of the H scratch
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int worker = blockIdx.x >> 1, n_workers = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmX);
    tma_prefetch_desc(&a.tmW1);
    tma_prefetch_desc(&a.tmW2);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(x_full, 1);
      mbar_init(x_empty, 1);
      for (int i = 0; i < MLP_W1_SLOTS; ++i) {
        mbar_init(&w1_full[i], 1);
        mbar_init(&w1_empty[i], 1);
      }
      for (int i = 0; i < MLP_W2_SLOTS; ++i) {
        mbar_init(&w2_full[i], 1);
        mbar_init(&w2_empty[i], 1);
      }
      mbar_init(h_full, 1);
      mbar_init(h_tfree, 16);
      mbar_init(h_ready, 16);
      mbar_init(h_sfree, 1);
      mbar_init(y_full, 1);
      mbar_init(y_free, 16);
      mbar_init(ln_go, 8);
      mbar_init(st_free, MLP_LN_WARPS);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  pdl_launch_dependents();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_Y = tmem_base, tmem_H = tmem_base + MLP_D;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs: own operand halves, leader's barriers) ------------------------------
    if (lane == 0) {
      uint32_t ts = 0, n1 = 0, n2 = 0;   // tile sequence, W1 / W2 atoms issued so far
      // One full / empty barrier pair per weight CHUNK (six W1 atoms, two W2 atoms): the MMA thread, which issues ~40 MMAs
      // per chunk on its own, then waits and commits four times per chunk instead of ten.
      auto load_w1 = [&](int j) {
        mbar_wait(&w1_empty[0], (n1 & 1) ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&w1_full[0], 2 * 6 * MLP_W1_SLOT);
        const uint32_t fb = mapa_shared(smem_u32(&w1_full[0]), 0);
        for (int at = 0; at < 6; ++at)
          tma_load_2d_pair(sW1 + at * MLP_W1_SLOT, &a.tmW1, fb, at * 64, j * MLP_CH + rank * 64);
        ++n1;
      };
      auto load_w2 = [&](int j) {
        mbar_wait(&w2_empty[0], (n2 & 1) ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&w2_full[0], 2 * 2 * MLP_W2_SLOT);
        const uint32_t fb = mapa_shared(smem_u32(&w2_full[0]), 0);
        for (int at = 0; at < 2; ++at) {
          tma_load_2d_pair(sW2 + at * MLP_W2_SLOT, &a.tmW2, fb, j * MLP_CH + at * 64, rank * 96);
          tma_load_2d_pair(sW2 + at * MLP_W2_SLOT + 96 * 128, &a.tmW2, fb, j * MLP_CH + at * 64, 192 + rank * 96);
        }
        ++n2;
      };
      for (int unit = worker; unit < a.n_pairs; unit += n_workers, ++ts) {
        const int row0 = (unit * 2 + rank) * 128;
        mbar_wait(x_empty, (ts & 1) ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(x_full, 2 * MLP_X_BYTES);
        {
          const uint32_t fb = mapa_shared(smem_u32(x_full), 0);
          for (int at = 0; at < 6; ++at) tma_load_2d_pair(sX + at * 16384, &a.tmX, fb, at * 64, row0);
        }
        load_w1(0);
        for (int j = 0; j < MLP_NCH; ++j) {   // the order the MMA warp consumes them: W1(j+1) before W2(j)
          if (j + 1 < MLP_NCH) load_w1(j + 1);
          load_w2(j);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA) ------------------------------
    if (lane == 0 && rank == 0) {
      constexpr uint32_t IDESC1 = umma_idesc(UMMA_FMT_BF16, MLP_CH, 0, 0, 256);
      constexpr uint32_t IDESC2 = umma_idesc(UMMA_FMT_BF16, 192, 0, 0, 256);
      uint32_t ts = 0, cs = 0, n1 = 0, n2 = 0;
      const uint64_t xd0 = umma_smem_desc_sw128(smem_u32(sX)), w1d0 = umma_smem_desc_sw128(smem_u32(sW1));
      const uint64_t hd0 = umma_smem_desc_sw128(smem_u32(sH)), w2d0 = umma_smem_desc_sw128(smem_u32(sW2));
      auto mma1 = [&]() {   // H = X W1_j^T over the six K atoms of the chunk
        mbar_wait(&w1_full[0], n1 & 1);
        tc_fence_after();
#pragma unroll
        for (int at = 0; at < 6; ++at) {
          const uint64_t ad = xd0 + (uint64_t)(at * (16384 >> 4));
          const uint64_t bd = w1d0 + (uint64_t)(at * (MLP_W1_SLOT >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_pair(tmem_H, ad + 2 * k, bd + 2 * k, IDESC1, (at | k) != 0);
        }
        umma_commit_pair(&w1_empty[0], 3);
        umma_commit_pair(h_full, 3);
        ++n1;
      };
      for (int unit = worker; unit < a.n_pairs; unit += n_workers, ++ts) {
        mbar_wait(x_full, ts & 1);
        tc_fence_after();
        if (cs > 0) {   // H of the previous tile's last chunk must have left TMEM
          mbar_wait(h_tfree, (cs - 1) & 1);
          tc_fence_after();
        }
        mma1();
        for (int j = 0; j < MLP_NCH; ++j, ++cs) {
          if (j + 1 < MLP_NCH) {
            mbar_wait(h_tfree, cs & 1);        // the warps hold H_j in registers: TMEM H may be overwritten
            tc_fence_after();
            mma1();
            if (j + 2 == MLP_NCH) umma_commit_pair(x_empty, 3);   // that was the tile's last MMA1: X may be reloaded
          }
          mbar_wait(h_ready, cs & 1);          // bf16 H_j is in shared memory (both CTAs)
          if (j == 0) mbar_wait(y_free, (ts & 1) ^ 1);            // Y of the previous tile has been read out
          tc_fence_after();
          mbar_wait(&w2_full[0], n2 & 1);
          tc_fence_after();
#pragma unroll
          for (int at = 0; at < 2; ++at) {
            const uint64_t ad = hd0 + (uint64_t)(at * (16384 >> 4));
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const uint64_t bd = w2d0 + (uint64_t)((at * MLP_W2_SLOT + hf * 96 * 128) >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_pair(tmem_Y + hf * 192, ad + 2 * k, bd + 2 * k, IDESC2, (j | at | k) != 0);
            }
          }
          umma_commit_pair(&w2_empty[0], 3);
          ++n2;
          umma_commit_pair(h_sfree, 3);
        }
        umma_commit_pair(y_full, 3);
      }
    }
  } else if (warp >= MLP_EPI_W0) {
    // ------------------------------ epilogue warps ------------------------------
    const int quarter = warp & 3, half = (warp - MLP_EPI_W0) >> 2;
    const int r = quarter * 32 + lane;                       // row of the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int sw = r & 7;
    const uint32_t h_row = smem_u32(sH) + half * 16384 + r * 128;       // this thread's row of H atom `half`
    const uint32_t xbuf = smem_u32(sH) + (half * 4 + quarter) * 4096;   // the same 4 KB block, as transposition scratch
    const uint32_t tfree_l = mapa_shared(smem_u32(h_tfree), 0), ready_l = mapa_shared(smem_u32(h_ready), 0);
    const uint32_t yfree_l = mapa_shared(smem_u32(y_free), 0);
    uint32_t ts = 0, cs = 0;
    int dbg_n = 0;
    const bool ts_on = kDbg && (a.epi.debug & 128) && blockIdx.x == 0 && threadIdx.x == 32 * MLP_EPI_W0;
    for (int unit = worker; unit < a.n_pairs; unit += n_workers, ++ts) {
      {
        // The residual rows this thread will read in the Y epilogue (twelve chunks = ~25 us from now) are requested from
        // HBM into L2 now: the fp32 stream does not stay in L2 across kernels, and the epilogue is otherwise exposed to one
        // HBM round trip per 32-column chunk while the tensor pipe waits for Y to be drained.
        const long long grow = (long long)(unit * 2 + rank) * 128 + r;
        if (grow < a.epi.M_total)
          bulk_prefetch_l2(reinterpret_cast<const float*>(a.epi.res) + grow * a.epi.ldres + half * 192, 192 * 4);
      }
      for (int j = 0; j < MLP_NCH; ++j, ++cs) {
        dbg_stamp(ts_on, dbg_n, 10);
        // b1 of this thread's 64 hidden columns: the same addresses in every lane (L1 broadcast), requested before the wait
        const float4* b4 = reinterpret_cast<const float4*>(a.b1 + j * MLP_CH + half * 64);
        mbar_wait(h_full, cs & 1);
        tc_fence_after();
        dbg_stamp(ts_on, dbg_n, 11);
        uint32_t v[64];
        tmem_ld64(tmem_H + lane_off + half * 64, v);
        tmem_ld_wait();
        dbg_stamp(ts_on, dbg_n, 12);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(tfree_l);
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 b = __ldg(b4 + i);
          const float2 g0 = gelu_fast2(fadd2(make_float2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), make_float2(b.x, b.y)));
          const float2 g1 = gelu_fast2(fadd2(make_float2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), make_float2(b.z, b.w)));
          pk[2 * i] = pack_bf16x2(g0.x, g0.y);
          pk[2 * i + 1] = pack_bf16x2(g1.x, g1.y);
        }
        dbg_stamp(ts_on, dbg_n, 13);
        mbar_wait(h_sfree, (cs & 1) ^ 1);     // MMA2 of the previous chunk has read the shared-memory H
        if (j == 0 && ts > 0 && a.ln_out) mbar_wait(st_free, (ts - 1) & 1);   // ... and the LayerNorm warps the statistics in it
        dbg_stamp(ts_on, dbg_n, 14);
#pragma unroll
        for (int p = 0; p < 8; ++p) st_shared_v4(h_row + ((p ^ sw) << 4), pk[4 * p], pk[4 * p + 1], pk[4 * p + 2], pk[4 * p + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(ready_l);
        dbg_stamp(ts_on, dbg_n, 15);
      }
      // ---- Y + b2, * ls2, + residual -> h (coalescing epilogue; this warpgroup takes 192 of the 384 columns) ----
      EpiTile t;
      t.dbg_n = dbg_n;
      t.tr = nullptr;
      t.r = r;
      t.g = 0;
      t.n0 = 0;
      t.grow = (long long)(unit * 2 + rank) * 128 + r;
      t.valid = t.grow < a.epi.M_total;
      t.q = (int)t.grow;
      t.rem = 0;
      t.taddr = tmem_Y + lane_off;
      // the last chunk's shared-memory H (which the scratch aliases) has been consumed once y_full completes; the wait is
      // inside the epilogue.  epilogue_linear_t<.., ACT_NONE, true>: y = (acc + b2) * ls2 + h
      if (!a.ln_out) {
        epilogue_linear_t<MLP_D, float, false, ACT_NONE, true>(a.epi, t, xbuf, y_full, ts & 1, half * 192, half * 192 + 192);
        dbg_n = t.dbg_n;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(yfree_l);
      } else {
        // ... and the statistics of the next block's LayerNorm1 over the stored values.  After the transposition lane
        // (sr, cg) holds columns cg*4.. of rows i*4 + sr (i < 8) of every 32-column chunk: sums and sums of squares ride
        // along in 16 registers, are completed over the 8 lanes of a row segment by shuffles and left in the top 256 B of the
        // warp's scratch for the LayerNorm warps (no barrier between epilogue warps: the two column halves stay decoupled).
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        epilogue_linear_t<MLP_D, float, false, ACT_NONE, true, true>(a.epi, t, xbuf, y_full, ts & 1, half * 192, half * 192 + 192, st);
        dbg_n = t.dbg_n;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(yfree_l);
        ln_store_partials(st, xbuf + 3072, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(ln_go);   // release: the tile's rows (global) and partial statistics (shared) are written
      }
    }
  } else {
    // ------------------------------ LayerNorm warps ------------------------------
    // The next block's norm1 (HF:367-372) of the 128 rows this CTA has just written, while the epilogue warps are already in
    // the next tile: one warp per row, 64 rows per warp, eight rows in flight from L2; lane l combines the two column halves'
    // sums of rows 2l, 2l + 1 into (mean, rstd) and hands them out by shuffle.
    // Replaces a standalone kernel that re-reads the whole fp32 residual stream from HBM (46 us per layer at batch 512).
    // The SM is issue-bound on the epilogue warps' GELU: every instruction spent here costs; three earlier forms of this
    // (tail on the epilogue warps, statistics + tail on them, full two-pass LayerNorm on these warps) all cost 46 us.
    if (a.ln_out) {
      const int w2 = warp - MLP_LN_W0;
      static_assert(MLP_LN_WARPS == 2, "ln_warp_rows<64>: 64 rows per LayerNorm warp");
      const int ldbg = kDbg ? a.ln_debug : 0;
      uint32_t ts = 0;
      for (int unit = worker; unit < a.n_pairs; unit += n_workers, ++ts) {
        const int lrow = w2 * 64 + 2 * lane;     // this lane finishes the statistics of tile rows lrow, lrow + 1
        const uint32_t at = smem_u32(sH) + (lrow >> 5) * 4096 + 3072 + (lrow & 31) * 8;
        mbar_wait_relaxed(ln_go, ts & 1);
        const float2 mr0 = ln_finish_stats(ld_shared_v2f(at), ld_shared_v2f(at + 16384), a.ln_eps);
        const float2 mr1 = ln_finish_stats(ld_shared_v2f(at + 8), ld_shared_v2f(at + 16384 + 8), a.ln_eps);
        __syncwarp();
        if (lane == 0) mbar_arrive(st_free);
        ln_warp_rows<64, true>(reinterpret_cast<const float*>(a.epi.out), a.epi.ldc, a.epi.M_total, (long long)(unit * 2 + rank) * 128 + w2 * 64,
                       mr0, mr1, a.ln_gamma, a.ln_beta, a.ln_out, a.ln_ld, lane, ldbg);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace vt
