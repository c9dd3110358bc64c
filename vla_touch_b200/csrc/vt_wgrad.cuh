// Weight gradients of the convolutions / linears straight from the channels-last activations (the backward of
// conditional_unet_1D.py:22-55 that torch autograd runs for bridge_train.py:330):
//
//     dW[g][r][tap * c_pad + c] = sum over (b, t) of  rows[g][b][t][r] * cols[g][b][t * stride + off_tap][c]
//
// The contraction index is the POSITION (b, t); both operands are stored with the CHANNEL contiguous, i.e. they are exactly the
// "MN-major" operand form of tcgen05.mma: a TMA box of 64 channels x 64 positions (128-byte rows, 128-byte swizzle) is a valid
// UMMA operand tile as it lands in shared memory (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: 64 channels
// per 128-byte row, 8 positions per swizzle atom, SBO = 1024 B between 8-position groups, LBO = one box between 64-channel
// blocks).  So no transposed copy of the operands is ever made: the first version materialised both operands K-major with
// tcol_kernel (2 + 2 bytes per element and tap), which was 34 % of the training program's time at batch 256.
//
// Same machinery as gemm_tc_kernel: persistent CTA pairs (cta_group::2, M = 256 = rows channels, N = 256 = (tap, channel)
// columns), warp 0 TMA producer (5-stage ring, one 64-position K tile per stage = four 8 KB boxes per CTA), warp 1 MMA issuer,
// warps 2-9 the fp32 linear epilogue of vt_gemm.cuh; conv taps are TMA coordinate shifts of the cols operand (out-of-range
// positions are zero-filled = the convolution's padding), stride-2 convolutions address the even / odd phase.
#pragma once
#include "vt_gemm.cuh"

namespace vt {

constexpr int WG_STAGES = 5;
constexpr int WG_KT = 64;                       // positions per K tile
constexpr int WG_BOX_BYTES = WG_KT * 128;       // one 64-channel x 64-position box
constexpr int WG_STAGE_BYTES = 4 * WG_BOX_BYTES;   // per CTA: 2 boxes of the rows operand + 2 boxes of the cols operand
constexpr int WG_THREADS = GEMM_THREADS(8);
constexpr int WG_SMEM_BYTES = 1024 + WG_STAGES * WG_STAGE_BYTES + 8 * 4096 + 256 + 2 * 2 * 256 * 4;

struct WgradArgs {
  CUtensorMap tmR;   // rows operand, 5-D (C, P, T, B, G), box (64, 1, t_box, b_box, 1)
  CUtensorMap tmC;   // cols operand, same form
  GemmArgs epi;      // output mapping for the linear epilogue (M_total = R, N, ldc, out, out_g, fast, vec ...)
  int rows_p, rows_t;            // phase / position offset of the rows operand
  int taps, c_pad;               // N index = tap * c_pad + channel
  int tap_p[GEMM_MAX_TAPS], tap_t[GEMM_MAX_TAPS];
  int k_tiles;                   // K tiles per output tile
  int k_t_step, k_b_step;        // (T, B) coordinate step per K tile
  int box_bytes;                 // bytes one box delivers (t_box * b_box * 128)
  int m_units, n_tiles, total_tiles;   // tile id = (g * m_units + m_unit) * n_tiles + n_tile
};

// shared-memory descriptor of an MN-major 128-byte-swizzled operand tile: `lbo` bytes between 64-element blocks along M / N
__device__ __forceinline__ uint64_t umma_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradArgs a) {
  constexpr int BN = 256;
  constexpr uint32_t ACC_COLS = BN, TMEM_COLS = 512;
  constexpr uint32_t IDESC = umma_idesc(UMMA_FMT_BF16, BN, 1, 1, 256);   // both operands MN-major
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sOp = smem;                                      // [stage][R0 | R1 | C0 | C1] boxes
  uint8_t* sX = smem + WG_STAGES * WG_STAGE_BYTES;           // transposition buffers of the coalescing epilogue
  uint64_t* full = reinterpret_cast<uint64_t*>(sX + 8 * 4096);
  uint64_t* empty = full + WG_STAGES;
  uint64_t* acc_full = empty + WG_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int worker = blockIdx.x / 2, n_workers = gridDim.x / 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.tmR);
    tma_prefetch_desc(&a.tmC);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < WG_STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&acc_full[s], 1);
        mbar_init(&acc_empty[s], 16);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tmem_relinquish_pair();
  }
  pdl_launch_dependents();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = worker; tile < a.total_tiles; tile += n_workers) {
        const int n_tile = tile % a.n_tiles, rest = tile / a.n_tiles;
        const int m_unit = rest % a.m_units, g = rest / a.m_units;
        const int r_c0 = m_unit * 256 + rank * 128;            // this CTA's 128 rows-operand channels
        const int nb0 = (n_tile * BN + rank * 128) / 64;       // ... and its two 64-column blocks of (tap, channel)
        int bt[2], bc[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int col = (nb0 + j) * 64;
          int tap = col / a.c_pad;
          bc[j] = col - tap * a.c_pad;
          bt[j] = tap < a.taps ? tap : 0;                      // columns past the last tap: any valid box (masked by the epilogue)
        }
        for (int k = 0; k < a.k_tiles; ++k) {
          mbar_wait(&empty[s], ph ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2u * 4u * (uint32_t)a.box_bytes);
          const uint32_t fb = mapa_shared(smem_u32(&full[s]), 0);
          uint8_t* st = sOp + s * WG_STAGE_BYTES;
          const int t0 = k * a.k_t_step, b0 = k * a.k_b_step;
          tma_load_5d_pair(st, &a.tmR, fb, r_c0, a.rows_p, t0 + a.rows_t, b0, g);
          tma_load_5d_pair(st + WG_BOX_BYTES, &a.tmR, fb, r_c0 + 64, a.rows_p, t0 + a.rows_t, b0, g);
          tma_load_5d_pair(st + 2 * WG_BOX_BYTES, &a.tmC, fb, bc[0], a.tap_p[bt[0]], t0 + a.tap_t[bt[0]], b0, g);
          tma_load_5d_pair(st + 3 * WG_BOX_BYTES, &a.tmC, fb, bc[1], a.tap_p[bt[1]], t0 + a.tap_t[bt[1]], b0, g);
          if (++s == WG_STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      uint32_t lt = 0, ph = 0;
      int s = 0;
      for (int tile = worker; tile < a.total_tiles; tile += n_workers, ++lt) {
        const uint32_t acc = lt & 1;
        mbar_wait(&acc_empty[acc], ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int k = 0; k < a.k_tiles; ++k) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(sOp + s * WG_STAGE_BYTES);
          const uint64_t adesc = umma_smem_desc_mn_sw128(st, WG_BOX_BYTES);
          const uint64_t bdesc = umma_smem_desc_mn_sw128(st + 2 * WG_BOX_BYTES, WG_BOX_BYTES);
#pragma unroll
          for (int j = 0; j < WG_KT / 16; ++j)   // 16 positions (two 8-position swizzle atoms, 2048 bytes) per instruction
            umma_f16_pair(d_tmem, adesc + (uint64_t)(j * 2048 >> 4), bdesc + (uint64_t)(j * 2048 >> 4), IDESC, (k | j) != 0);
          umma_commit_pair(&empty[s], 3);
          if (++s == WG_STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit_pair(&acc_full[acc], 3);
      }
    }
  } else {
    const int half = (warp - 2) >> 2, quarter = warp & 3;
    const int et256 = threadIdx.x - 64;
    const GemmArgs& e = a.epi;
    EpiTile t;
    t.r = quarter * 32 + lane;
    t.dbg_n = 0;
    t.tr = nullptr;
    uint32_t lt = 0;
    for (int tile = worker; tile < a.total_tiles; tile += n_workers, ++lt) {
      const uint32_t acc = lt & 1;
      const int n_tile = tile % a.n_tiles, rest = tile / a.n_tiles;
      const int m_tile = (rest % a.m_units) * 2 + rank;
      t.g = rest / a.m_units;
      t.n0 = n_tile * BN;
      float* colv = scratch + acc * 2 * BN;
      const bool fast = e.fast != 0;
      if (!fast) {
        for (int c = et256; c < BN; c += 256) {
          colv[c] = 0.f;
          colv[BN + c] = 1.f;
        }
        named_bar_sync(1, 256);
      }
      t.grow = (long long)m_tile * 128 + t.r;
      t.valid = t.grow < e.M_total;
      t.q = (int)t.grow;
      t.rem = 0;
      t.taddr = tmem_base + acc * ACC_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
      const uint32_t parity = (lt >> 1) & 1;
      const int c_begin = half * (BN / 2), c_end = c_begin + BN / 2;
      if (fast) epilogue_linear_fast<BN, float, false>(e, t, smem_u32(sX) + (warp - 2) * 4096, &acc_full[acc], parity, c_begin, c_end);
      else epilogue_linear<BN, float, false>(e, t, colv, &acc_full[acc], parity, c_begin, c_end);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(&acc_empty[acc]), 0));
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, TMEM_COLS);
}

}  // namespace vt
