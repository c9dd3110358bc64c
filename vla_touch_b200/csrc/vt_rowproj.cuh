// Attention output projection of a DinoV2-S block with whole rows per CTA (HF Dinov2SelfOutput + layer_scale1 + residual,
// HF:238-251,374-378) and the block's norm2 (HF:380-381) out of the same kernel, hidden size D = 384:
//     h += ls1 * (ctx Wo^T + bo),        xn = LayerNorm2(h) as bf16
// The generic GEMM tiles N (256 + 128 columns for D = 384), so no CTA sees a whole row and LayerNorm2 had to be a kernel of
// its own that re-reads the fp32 residual stream from HBM (46 us per layer at batch 512, on the HBM roof).  Here a CTA
// pair owns 256 rows x all 384 columns (tcgen05 cta_group::2, two N = 192 MMAs per K step, accumulator = 384 TMEM columns):
// the row statistics ride along in the epilogue and two warps normalise the rows from L2 while the next tile is computed,
// exactly as in mlp_fused_kernel (vt_mlp.cuh), whose epilogue / LayerNorm helpers this kernel uses.
//
// PERSISTENT CTA pairs.  Per pair and 256-row tile:
//     TMA    X tile (this CTA's 128 rows x 384, six 64-wide K atoms) -- stays in shared memory for the whole tile
//     for each of the six 64-column output chunks c (a ring of three W stages):
//       TMA  W chunk (this CTA's 32 of Wo's rows [64c, 64c + 64) x 384, six K atoms of 4 KB)
//       MMA  Y[:, 64c .. 64c + 64) = X Wo_c^T        (N = 64, 24 K = 16 steps) as soon as the epilogue warps have read
//            those 64 TMEM columns of the PREVIOUS tile
//   Y + bo, * ls1, + residual -> h through the coalescing epilogue of the GEMM kernel (vt_gemm.cuh), 64 columns at a time.
// The accumulator is single-buffered (2 x 384 columns do not fit TMEM) but handed back chunk by chunk, so the 3 us of MMAs
// per tile run under the epilogue of the previous tile (a whole-tile hand-back cost +26 us per launch).  The kernel moves
// 606 MB per launch at batch 512 (ctx 101 + h 202 in, h 202 + xn 101 out), an HBM floor of 93 us, against 505 + 303 MB for
// the two kernels it replaces.
//   warp 0  TMA producer   warp 1  MMA issuer (leader CTA)   warps 4-11  epilogue (quarter = warp & 3, column half = (warp-4)/4)
//   warps 12-15  LayerNorm (32 rows each: two warps cannot keep up with a 14 us tile).  Whole warpgroups per role + setmaxnreg.
#pragma once
#include "vt_mlp.cuh"

namespace vt {

constexpr int RP_D = 384, RP_THREADS = 512, RP_WSTAGES = 3, RP_KATOMS = RP_D / 64, RP_CHUNKS = RP_D / 64;
constexpr int RP_REGS_CTRL = 56, RP_REGS_EPI = 160, RP_REGS_LN = 136;
static_assert(RP_REGS_CTRL * 128 + RP_REGS_EPI * 256 + RP_REGS_LN * 128 <= 128 * RP_THREADS,
              "setmaxnreg can only redistribute the registers the launch allocated (128 per thread at 16 warps)");
constexpr int RP_X_ATOM = 128 * 128;            // 128 rows x 128 B
constexpr int RP_W_ATOM = 32 * 128;             // this CTA's 32 rows of a 64-column chunk, one K atom
constexpr int RP_W_STAGE = RP_KATOMS * RP_W_ATOM;
constexpr int RP_SMEM_W = RP_KATOMS * RP_X_ATOM;
constexpr int RP_SMEM_SCR = RP_SMEM_W + RP_WSTAGES * RP_W_STAGE;   // 8 x 4 KB transposition scratch of the epilogue warps
constexpr int RP_SMEM_ST = RP_SMEM_SCR + 8 * 4096;                 // partial row statistics [half][128] float2
constexpr int RP_SMEM_BAR = RP_SMEM_ST + 2048;
constexpr int RP_SMEM_BYTES = 1024 + RP_SMEM_BAR + 256;
static_assert(RP_X_ATOM % 1024 == 0 && RP_W_ATOM % 1024 == 0, "128B-swizzled atoms are 1024-byte aligned");
static_assert(RP_SMEM_BYTES <= 227 * 1024, "rowproj_kernel shared memory");

struct RowprojArgs {
  CUtensorMap tmX;    // 2-D (D, rows) over ctx (bf16), box (64, 128)
  CUtensorMap tmW;    // 2-D (D, D) over Wo (bf16, K contiguous), box (64, 32)
  GemmArgs epi;       // output side: out = res = h (fp32, ld D), bias = bo, colscale = ls1, M_total = rows, rows_valid = 128 ...
  int m_tiles;        // 128-row tiles
  int n_pairs;        // ceil(m_tiles / 2) work units
  const float* ln_gamma;   // [D] or null: LayerNorm of the updated rows (the block's norm2)
  const float* ln_beta;
  __nv_bfloat16* ln_out;   // [rows][ln_ld]
  long long ln_ld;
  float ln_eps;
  int ln_debug;            // developer knobs (VT_DEBUG_KNOBS builds): 1 no stores, 2 no loads, 4 handshake only
};

__global__ void __launch_bounds__(RP_THREADS, 1) rowproj_kernel(const __grid_constant__ RowprojArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sScr = smem + RP_SMEM_SCR;
  const uint32_t st_tab = smem_u32(smem + RP_SMEM_ST);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RP_SMEM_BAR);
  uint64_t* x_full = bars;                 // leader: both CTAs' X tiles have landed
  uint64_t* x_empty = bars + 1;            // both (multicast commit): the tile's last MMA has read X
  uint64_t* w_full = bars + 2;             // [3] leader
  uint64_t* w_empty = bars + 5;            // [3] both
  uint64_t* c_full = bars + 8;             // [6] both: the MMAs of output chunk c are complete
  uint64_t* c_free = bars + 14;            // [6] leader: the epilogue warps of the pair (4 + 4) have read chunk c out of TMEM
  uint64_t* ln_go = bars + 20;             // own CTA: the eight epilogue warps have stored the tile's rows and partial statistics
  uint64_t* st_free = bars + 21;           // own CTA: the LayerNorm warps have read the partial statistics
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int worker = blockIdx.x >> 1, n_workers = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmX);
    tma_prefetch_desc(&a.tmW);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(x_full, 1);
      mbar_init(x_empty, 1);
      for (int i = 0; i < RP_WSTAGES; ++i) {
        mbar_init(&w_full[i], 1);
        mbar_init(&w_empty[i], 1);
      }
      for (int i = 0; i < RP_CHUNKS; ++i) {
        mbar_init(&c_full[i], 1);
        mbar_init(&c_free[i], 8);
      }
      mbar_init(ln_go, 8);
      mbar_init(st_free, 4);
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
  const uint32_t tmem_Y = *tmem_slot;
  pdl_wait();

  if (warp < 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RP_REGS_CTRL));
   if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs: own operand halves, leader's barriers) ------------------------------
    if (lane == 0) {
      uint32_t n = 0, ts = 0;
      for (int unit = worker; unit < a.n_pairs; unit += n_workers, ++ts) {
        const int row0 = (unit * 2 + rank) * 128;
        mbar_wait(x_empty, (ts & 1) ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(x_full, 2 * RP_KATOMS * RP_X_ATOM);
        {
          const uint32_t fb = mapa_shared(smem_u32(x_full), 0);
          for (int at = 0; at < RP_KATOMS; ++at) tma_load_2d_pair(smem + at * RP_X_ATOM, &a.tmX, fb, at * 64, row0);
        }
        for (int c = 0; c < RP_CHUNKS; ++c, ++n) {
          const uint32_t s = n % RP_WSTAGES;
          uint8_t* st = smem + RP_SMEM_W + s * RP_W_STAGE;
          mbar_wait(&w_empty[s], ((n / RP_WSTAGES) & 1) ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&w_full[s], 2 * RP_W_STAGE);
          const uint32_t fb = mapa_shared(smem_u32(&w_full[s]), 0);
          for (int at = 0; at < RP_KATOMS; ++at) tma_load_2d_pair(st + at * RP_W_ATOM, &a.tmW, fb, at * 64, c * 64 + rank * 32);
        }
      }
    }
   } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA) ------------------------------
    if (lane == 0 && rank == 0) {
      constexpr uint32_t IDESC = umma_idesc(UMMA_FMT_BF16, 64, 0, 0, 256);
      const uint64_t xd0 = umma_smem_desc_sw128(smem_u32(smem));
      const uint64_t wd0 = umma_smem_desc_sw128(smem_u32(smem + RP_SMEM_W));
      uint32_t n = 0, ts = 0;
      for (int unit = worker; unit < a.n_pairs; unit += n_workers, ++ts) {
        mbar_wait(x_full, ts & 1);
        tc_fence_after();
        for (int c = 0; c < RP_CHUNKS; ++c, ++n) {
          const uint32_t s = n % RP_WSTAGES;
          mbar_wait(&w_full[s], (n / RP_WSTAGES) & 1);
          mbar_wait(&c_free[c], (ts & 1) ^ 1);     // the previous tile's chunk c has left TMEM
          tc_fence_after();
#pragma unroll 1
          for (int at = 0; at < RP_KATOMS; ++at) {
            const uint64_t ad = xd0 + (uint64_t)((at * RP_X_ATOM) >> 4);
            const uint64_t bd = wd0 + (uint64_t)((s * RP_W_STAGE + at * RP_W_ATOM) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_pair(tmem_Y + c * 64, ad + 2 * k, bd + 2 * k, IDESC, (at | k) != 0);
          }
          umma_commit_pair(&w_empty[s], 3);
          umma_commit_pair(&c_full[c], 3);
        }
        umma_commit_pair(x_empty, 3);
      }
    }
   }
  } else if (warp < 12) {
    // ------------------------------ epilogue warps ------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RP_REGS_EPI));
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    const int r = quarter * 32 + lane;                       // row of the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t xbuf = smem_u32(sScr) + (half * 4 + quarter) * 4096;
    // (Requesting the residual rows of the NEXT tile into L2 here was measured: 100 MB of them are evicted again before they are
    //  used -- dram__bytes_read 303 -> 402 MB per launch -- so the epilogue's own just-in-time prefetch is all there is.)
    uint32_t ts = 0;
    for (int unit = worker; unit < a.n_pairs; unit += n_workers, ++ts) {
      EpiTile t;
      t.dbg_n = 0;
      t.tr = nullptr;
      t.r = r;
      t.g = 0;
      t.n0 = 0;
      t.grow = (long long)(unit * 2 + rank) * 128 + r;
      t.valid = t.grow < a.epi.M_total;
      t.q = (int)t.grow;
      t.rem = 0;
      t.taddr = tmem_Y + lane_off;
      float st[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) st[i] = 0.f;
      // y = (acc + bo) * ls1 + h, this warp's three 64-column chunks; each goes back to the MMA warp as soon as it is out of TMEM
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        const int c = half * 3 + j;
        epilogue_linear_t<RP_D, float, false, ACT_NONE, true, true>(a.epi, t, xbuf, &c_full[c], ts & 1, c * 64, c * 64 + 64, st);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(&c_free[c]), 0));
      }
      if (a.ln_out) {
        if (ts > 0) mbar_wait(st_free, (ts - 1) & 1);   // the LayerNorm warps have read the previous tile's partials (always)
        ln_store_partials(st, st_tab + half * 1024 + quarter * 256, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(ln_go);   // release: the tile's rows (global) and partial statistics (shared) are written
      }
    }
  } else {
    // ------------------------------ LayerNorm warps (see vt_mlp.cuh) ------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RP_REGS_LN));
    if (a.ln_out) {
      const int w4 = warp - 12;
      uint32_t ts = 0;
      for (int unit = worker; unit < a.n_pairs; unit += n_workers, ++ts) {
        const uint32_t at = st_tab + (w4 * 32 + lane) * 8;   // this lane finishes the statistics of tile row w4*32 + l
        mbar_wait_relaxed(ln_go, ts & 1);
        const float2 mr = ln_finish_stats(ld_shared_v2f(at), ld_shared_v2f(at + 1024), a.ln_eps);
        __syncwarp();
        if (lane == 0) mbar_arrive(st_free);
        ln_warp_rows<32, false>(reinterpret_cast<const float*>(a.epi.out), a.epi.ldc, a.epi.M_total,
                                (long long)(unit * 2 + rank) * 128 + w4 * 32, mr, mr, a.ln_gamma, a.ln_beta, a.ln_out, a.ln_ld, lane, kDbg ? a.ln_debug : 0);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_Y, 512);
}

}  // namespace vt
