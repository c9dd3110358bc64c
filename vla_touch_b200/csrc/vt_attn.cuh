// Multi-head attention of the DinoV2 blocks (HF:199-235: softmax(Q K^T / sqrt(64)) V, head_dim 64) over the packed
// qkv rows written by the QKV GEMM.
//
// attn_tc_kernel (bf16): one CTA = one (image, head, 128-query tile).  Flash-style loop over 128-key blocks:
//   warp 8  TMA producer : Q tile once, then K_j / V_j tiles through a 3-slot ring (SWIZZLE_128B boxes of the
//                          same 2-D tensor map over qkv, different column coordinate)
//   warp 9  UMMA issuer  : S = Q K_j^T (M128 x N<=128 x K64, accumulator in TMEM columns [0,128)), then
//                          O_j = P_j V_j (M128 x N64 x K<=128; P from shared memory, V as an MN-major operand)
//                          into TMEM columns [128,192)
//   warps 0-7 softmax    : two threads per query row (64 key columns / 32 output channels each); online max / sum
//                          in fp32, exp2 with the 1/sqrt(d) scale folded in, P written to shared memory as bf16 in
//                          the 128B-swizzled K-major layout the UMMA descriptor expects, running O kept in
//                          registers and rescaled per block.
// 96 KB of shared memory and 256 TMEM columns per CTA -> two CTAs per SM overlap each other's softmax and MMA.
//
// attn_f32_kernel: fp32 CUDA-core variant used by the parity ("precise") mode only.
#pragma once
#include "vt_elem.cuh"
#include "vt_ptx.cuh"

namespace vt {

constexpr int ATT_THREADS = 320;           // warps 0-7 softmax (two threads per query row), warp 8 TMA, warp 9 MMA
constexpr int ATT_TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr int ATT_KV_SLOTS = 3;
constexpr int ATT_SMEM_BYTES = 1024 + ATT_TILE_BYTES * (1 + ATT_KV_SLOTS + 2) + 256 + 2 * 256 * 4;
constexpr int ATT_TAIL_MAX = 16;           // a last query tile with <= this many rows goes to attn_tail_kernel

struct AttnArgs {
  CUtensorMap tm;  // 2-D (3*D, images*tokens), box (64, 128), SWIZZLE_128B
  __nv_bfloat16* ctx;
  long long ctx_ld;
  int tokens, heads, D;
  float scale_log2;  // log2(e) / sqrt(head_dim)
};

__global__ void __launch_bounds__(ATT_THREADS, 2) attn_tc_kernel(const __grid_constant__ AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + ATT_TILE_BYTES;
  uint8_t* sP = sKV + ATT_KV_SLOTS * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + ATT_KV_SLOTS;
  uint64_t* s_full = kv_empty + ATT_KV_SLOTS;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2][256] row max / row sum exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  const int N = a.tokens;
  const int nb = (N + 127) >> 7;
  const int row0 = img * N;

  if (warp == 9) {
    if (lane == 0) {
      mbar_init(q_full, 1);
      for (int i = 0; i < ATT_KV_SLOTS; ++i) {
        mbar_init(&kv_full[i], 1);
        mbar_init(&kv_empty[i], 1);
      }
      mbar_init(s_full, 1);
      mbar_init(p_full, 256);
      mbar_init(o_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  if (warp == 8 && lane == 0) tma_prefetch_desc(&a.tm);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  if (warp == 8) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_2d(sQ, &a.tm, q_full, head * 64, row0 + qt * 128);
      for (int i = 0; i < 2 * nb; ++i) {
        const int slot = i % ATT_KV_SLOTS;
        const uint32_t ph = (i / ATT_KV_SLOTS) & 1;
        mbar_wait(&kv_empty[slot], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[slot], ATT_TILE_BYTES);
        const int col = ((i & 1) ? 2 * a.D : a.D) + head * 64;
        tma_load_2d(sKV + slot * ATT_TILE_BYTES, &a.tm, &kv_full[slot], col, row0 + (i >> 1) * 128);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      mbar_wait(q_full, 0);
      const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(sQ));
      const uint64_t pdesc = umma_smem_desc_sw128(smem_u32(sP));
      for (int j = 0; j < nb; ++j) {
        const int valid = min(128, N - j * 128);
        const int ncols = (valid + 15) & ~15;
        int i = 2 * j;
        int slot = i % ATT_KV_SLOTS;
        mbar_wait(&kv_full[slot], (i / ATT_KV_SLOTS) & 1);
        tc_fence_after();
        {
          const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sKV + slot * ATT_TILE_BYTES));
          const uint32_t idesc = umma_idesc(UMMA_FMT_BF16, (uint32_t)ncols);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_S, qdesc + 2 * k, kdesc + 2 * k, idesc, k != 0);
        }
        umma_commit(&kv_empty[slot]);
        umma_commit(s_full);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        i = 2 * j + 1;
        slot = i % ATT_KV_SLOTS;
        mbar_wait(&kv_full[slot], (i / ATT_KV_SLOTS) & 1);
        tc_fence_after();
        {
          // V tile: rows = keys (the MMA's K index), 128 B = the head's 64 channels (N index): MN-major operand
          const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(sKV + slot * ATT_TILE_BYTES));
          const uint32_t idesc = umma_idesc(UMMA_FMT_BF16, 64, 0, 1);
          for (int kk = 0; kk < ncols / 16; ++kk) {
            const uint64_t ad = pdesc + (uint64_t)((kk >> 2) * (ATT_TILE_BYTES >> 4) + 2 * (kk & 3));
            const uint64_t bd = vdesc + (uint64_t)(kk * (16 * 128 >> 4));
            umma_f16(tmem_O, ad, bd, idesc, kk != 0);
          }
        }
        umma_commit(&kv_empty[slot]);
        umma_commit(o_full);
      }
    }
  } else {
    // ------------------------------ softmax + output ------------------------------
    // Thread (r, half): row r = (warp & 3) * 32 + lane (its TMEM lane), key columns [half*64, half*64+64) of every
    // 128-key block (= one 64-key chunk of P), output channels [half*32, half*32+32).
    const int half = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float sl = a.scale_log2;
    float m = -INFINITY, l = 0.f;
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    const uint32_t sP_row = smem_u32(sP) + half * ATT_TILE_BYTES + r * 128;
    const int sw = r & 7;
    float* my_x = xch + half * 128 + r;
    const float* peer_x = xch + (half ^ 1) * 128 + r;

    for (int j = 0; j < nb; ++j) {
      const int valid = min(128, N - j * 128) - half * 64;   // valid keys among this thread's 64 columns (may be <= 0)
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // S row in four 16-column chunks, two TMEM loads in flight (register budget: 2 CTAs x 320 threads per SM).
      // Chunks without valid keys are skipped (the MMA reads only round16(valid) columns of P), full chunks take a
      // select-free path: the kernel is instruction-issue bound, not MUFU or tensor bound.
      const uint32_t s_addr = tmem_S + lane_off + half * 64;
      uint32_t va[16], vb[16];
      float mx = -INFINITY;
      const bool act0 = valid > 0, act1 = valid > 16, act2 = valid > 32, act3 = valid > 48;
      auto max16 = [&](const uint32_t(&v)[16], int c0) {
        if (c0 + 16 <= valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, (c0 + i < valid) ? __uint_as_float(v[i]) : -INFINITY);
        }
      };
      if (act0) tmem_ld16(s_addr, va);
      tmem_ld_wait();
      if (act1) tmem_ld16(s_addr + 16, vb);
      if (act0) max16(va, 0);
      tmem_ld_wait();
      if (act2) tmem_ld16(s_addr + 32, va);
      if (act1) max16(vb, 16);
      tmem_ld_wait();
      if (act3) tmem_ld16(s_addr + 48, vb);
      if (act2) max16(va, 32);
      tmem_ld_wait();
      if (act3) max16(vb, 48);
      my_x[(j & 1) * 256] = mx;
      if (act0) tmem_ld16(s_addr, va);            // first chunk of the second pass, overlapped with the exchange
      named_bar_sync(1, 256);
      const float m_new = fmaxf(m, fmaxf(mx, peer_x[(j & 1) * 256]));
      const float alpha = ex2_approx((m - m_new) * sl);
      const float msc = m_new * sl;
      float rowsum = 0.f;
      auto exp16 = [&](const uint32_t(&v)[16], int c0) {
        const bool full = c0 + 16 <= valid;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = u * 8 + 2 * e;
            float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), sl, -msc));
            float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sl, -msc));
            if (!full) {
              p0 = (c0 + i < valid) ? p0 : 0.f;
              p1 = (c0 + i + 1 < valid) ? p1 : 0.f;
            }
            rowsum += p0 + p1;
            pk[e] = pack_bf16x2(p0, p1);
          }
          st_shared_v4(sP_row + ((((c0 >> 3) + u) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
      };
      tmem_ld_wait();
      if (act1) tmem_ld16(s_addr + 16, vb);
      if (act0) exp16(va, 0);
      tmem_ld_wait();
      if (act2) tmem_ld16(s_addr + 32, va);
      if (act1) exp16(vb, 16);
      tmem_ld_wait();
      if (act3) tmem_ld16(s_addr + 48, vb);
      if (act2) exp16(va, 32);
      tmem_ld_wait();
      if (act3) exp16(vb, 48);
      l = l * alpha + rowsum;
      m = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      {
        uint32_t w[32];
        tmem_ld32(tmem_O + lane_off + half * 32, w);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], alpha, __uint_as_float(w[i]));
      }
    }
    // combine the two halves' row sums
    my_x[0] = l;
    named_bar_sync(1, 256);
    const float inv = 1.0f / (l + peer_x[0]);
    const int qrow = qt * 128 + r;
    if (qrow < N) {
      __nv_bfloat16* dst = a.ctx + (long long)(row0 + qrow) * a.ctx_ld + head * 64 + half * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 w;
        w.x = pack_bf16x2(o[i] * inv, o[i + 1] * inv);
        w.y = pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
        w.z = pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv);
        w.w = pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
        *reinterpret_cast<uint4*>(dst + i) = w;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------------------
// attn_row_kernel (bf16, 256 <= tokens <= 272, e.g. the 257 tokens of a 224 x 224 image): PERSISTENT, one CTA per SM,
// work unit = one (image, head).  The whole key range fits TMEM, so there is no online-softmax rescaling:
//   warp 8   TMA producer: K tiles (2 x 128 rows + a 16-row tail box; double-buffered across units), V tiles (single
//            buffer: free again one query tile before the next unit needs it), Q tiles (2-deep ring)
//   warp 9   UMMA issuer : S = Q K^T for ALL keys into TMEM columns [0, 272), issued in 64-key slices interleaved with
//            the previous tile's O = P V slices (a slice of S may be overwritten once the p_full barrier of the chunk
//            it holds has completed); two O accumulators (columns 288 / 352) alternate per query tile
//   warps 0-7  softmax   : two threads per query row (keys [0,128) / [128, 272)); the 128 scores are read from TMEM
//            once (two x64 loads in flight) and stay in registers for the max pass and the exp pass;
//            p = 2^(s*c - m*c) -> bf16 P chunk tiles in shared memory (128B-swizzled K-major UMMA operand) + row sums;
//            O of tile t-1 is read out (scaled by 1/sum, 64 contiguous bytes per row) while tile t's MMAs drain
//   warps 10-11  tail rows: the 1..16 query rows behind the last full tile (the CLS-shifted 257th token) on the CUDA
//            cores, straight from the K / V tiles that are already in shared memory - no 128-row tile that is 99 %
//            padding, no second pass over K and V in HBM.
// Per full query tile the MUFU pipe (128 x 257 exponentials, 16 / clock / SM) is the floor: ~2100 of ~3400 cycles.
// ------------------------------------------------------------------------------------------------------------
constexpr int ATR_THREADS = 384;
constexpr int ATR_KBUF = 2 * ATT_TILE_BYTES + 2048;           // two 128-row tiles + one 16-row tail tile
constexpr int ATR_SMEM_K = 0;                                 // [2][ATR_KBUF]
constexpr int ATR_SMEM_V = 2 * ATR_KBUF;                      // [ATR_KBUF]
constexpr int ATR_SMEM_Q = 3 * ATR_KBUF;                      // [2][16 KB]   (also the O staging tile)
constexpr int ATR_SMEM_P = ATR_SMEM_Q + 2 * ATT_TILE_BYTES;   // [5][16 KB]   64-key chunks of P (chunk 4: 16 keys)
constexpr int ATR_SMEM_BAR = ATR_SMEM_P + 5 * ATT_TILE_BYTES; // 32 mbarriers
constexpr int ATR_SMEM_XCH = ATR_SMEM_BAR + 256;              // [2][2][128] floats: row max / row sum exchange
constexpr int ATR_SMEM_TAIL = ATR_SMEM_XCH + 4096;            // tail warps: q[64], p[272], red[8], part[2][64] floats
constexpr int ATR_SMEM_BYTES = 1024 + ATR_SMEM_TAIL + (64 + 272 + 8 + 128) * 4 + 16;
static_assert(ATR_SMEM_BYTES <= 227 * 1024, "attn_row_kernel shared memory");

struct AttnRowArgs {
  CUtensorMap tm;    // 2-D (3*D, images*tokens) over qkv, box (64, 128), SWIZZLE_128B
  CUtensorMap tm16;  // same tensor, box (64, 16): the K / V tail rows
  CUtensorMap tmO;   // 2-D (D, images*tokens) over ctx (row stride ctx_ld), box (64, 128), SWIZZLE_128B
  const __nv_bfloat16* qkv;
  __nv_bfloat16* ctx;
  long long ctx_ld;
  int tokens, heads, D, units;   // units = images * heads
  float scale_log2;
};

__global__ void __launch_bounds__(ATR_THREADS, 1) attn_row_kernel(const __grid_constant__ AttnRowArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sK = smem + ATR_SMEM_K;
  uint8_t* sV = smem + ATR_SMEM_V;
  uint8_t* sQ = smem + ATR_SMEM_Q;
  uint8_t* sP = smem + ATR_SMEM_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATR_SMEM_BAR);
  uint64_t* k_full = bars;            // [2]
  uint64_t* k_empty = bars + 2;       // [2]
  uint64_t* v_full = bars + 4;
  uint64_t* v_empty = bars + 5;
  uint64_t* q_full = bars + 6;        // [2]
  uint64_t* q_empty = bars + 8;       // [2]
  uint64_t* s_full = bars + 10;
  uint64_t* p_full = bars + 12;       // [5]
  uint64_t* o_full = bars + 17;       // [2]  (O accumulators alternate per query tile)
  uint64_t* o_empty = bars + 19;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  float* xch = reinterpret_cast<float*>(smem + ATR_SMEM_XCH);
  float* tq = reinterpret_cast<float*>(smem + ATR_SMEM_TAIL);   // [64]
  float* tp = tq + 64;                                           // [272]
  float* tred = tp + 272;                                        // [8]
  float* tpart = tred + 8;                                       // [2][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.tokens;
  const int n_tail = N - 256;              // 0..16 keys / query rows behind the two full tiles
  const bool has_tail = n_tail > 0;

  if (warp == 9) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&k_empty[i], has_tail ? 3 : 1);   // MMA commit + the two tail warps
        mbar_init(&q_full[i], 1);
        mbar_init(&q_empty[i], 1);
      }
      mbar_init(v_full, 1);
      mbar_init(v_empty, has_tail ? 3 : 1);
      mbar_init(s_full, 1);
      for (int i = 0; i < 5; ++i) mbar_init(&p_full[i], 128);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&o_full[i], 1);
        mbar_init(&o_empty[i], 256);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&a.tm);
    tma_prefetch_desc(&a.tm16);
    tma_prefetch_desc(&a.tmO);
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 288;   // O accumulators at columns 288 and 352
  pdl_wait();

  if (warp == 8) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t us = 0, qs = 0;
      for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x, ++us) {
        const int img = unit / a.heads, head = unit - img * a.heads;
        const int row0 = img * N;
        const uint32_t kb = us & 1;
        uint8_t* kbuf = sK + kb * ATR_KBUF;
        mbar_wait(&k_empty[kb], ((us >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[kb], 2 * ATT_TILE_BYTES + (has_tail ? 2048 : 0));
        tma_load_2d(kbuf, &a.tm, &k_full[kb], a.D + head * 64, row0);
        tma_load_2d(kbuf + ATT_TILE_BYTES, &a.tm, &k_full[kb], a.D + head * 64, row0 + 128);
        if (has_tail) tma_load_2d(kbuf + 2 * ATT_TILE_BYTES, &a.tm16, &k_full[kb], a.D + head * 64, row0 + 256);
        {   // first Q tile before V: it is needed first
          const uint32_t qb = qs & 1;
          mbar_wait(&q_empty[qb], ((qs >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[qb], ATT_TILE_BYTES);
          tma_load_2d(sQ + qb * ATT_TILE_BYTES, &a.tm, &q_full[qb], head * 64, row0);
          ++qs;
        }
        mbar_wait(v_empty, (us & 1) ^ 1);
        mbar_arrive_expect_tx(v_full, 2 * ATT_TILE_BYTES + (has_tail ? 2048 : 0));
        tma_load_2d(sV, &a.tm, v_full, 2 * a.D + head * 64, row0);
        tma_load_2d(sV + ATT_TILE_BYTES, &a.tm, v_full, 2 * a.D + head * 64, row0 + 128);
        if (has_tail) tma_load_2d(sV + 2 * ATT_TILE_BYTES, &a.tm16, v_full, 2 * a.D + head * 64, row0 + 256);
        {
          const uint32_t qb = qs & 1;
          mbar_wait(&q_empty[qb], ((qs >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[qb], ATT_TILE_BYTES);
          tma_load_2d(sQ + qb * ATT_TILE_BYTES, &a.tm, &q_full[qb], head * 64, row0 + 128);
          ++qs;
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------ UMMA issuer ------------------------------
    // S of query tile t+1 is issued in 64-key slices INTERLEAVED with the P V slices of tile t: slice c of S may be
    // overwritten as soon as p_full[c] of tile t has completed (every row has read that slice and published its P), so
    // the next tile's scores are ready when the softmax warps come back from the O read-out instead of one S MMA later.
    if (lane == 0 && blockIdx.x < a.units) {
      const uint32_t idesc_s = umma_idesc(UMMA_FMT_BF16, 64);
      const uint32_t idesc_st = umma_idesc(UMMA_FMT_BF16, 16);
      const uint32_t idesc_o = umma_idesc(UMMA_FMT_BF16, 64, 0, 1);
      const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(sV));
      const uint64_t pdesc = umma_smem_desc_sw128(smem_u32(sP));
      const uint32_t n_tiles = 2u * (uint32_t)((a.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
      // S slice c (c = 4: the 16-key tail) of tile t: keys [64 c, 64 c + 64) are rows of the unit's contiguous K tiles
      auto issue_s = [&](uint32_t t, int c) {
        const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(sQ + (t & 1) * ATT_TILE_BYTES));
        const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sK + ((t >> 1) & 1) * ATR_KBUF)) + (uint64_t)(c * 64 * 128 >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_S + c * 64, qdesc + 2 * k, kdesc + 2 * k, c < 4 ? idesc_s : idesc_st, k != 0);
      };
      auto wait_inputs = [&](uint32_t t) {   // K tiles (first tile of a unit) and Q tile of tile t
        if ((t & 1) == 0) mbar_wait(&k_full[(t >> 1) & 1], (t >> 2) & 1);
        mbar_wait(&q_full[t & 1], (t >> 1) & 1);
        tc_fence_after();
      };
      wait_inputs(0);
      for (int c = 0; c < 4; ++c) issue_s(0, c);
      if (has_tail) issue_s(0, 4);
      umma_commit(s_full);
      umma_commit(&q_empty[0]);
      for (uint32_t t = 0; t < n_tiles; ++t) {
        const bool more = t + 1 < n_tiles;
        if ((t & 1) == 0) mbar_wait(v_full, (t >> 1) & 1);
        const uint32_t o_tmem = tmem_O + (t & 1) * 64;
        mbar_wait(&o_empty[t & 1], ((t >> 1) & 1) ^ 1);   // the O accumulator of two tiles ago has been read out
        if (more) wait_inputs(t + 1);
        tc_fence_after();
        for (int c = 0; c < 4; ++c) {             // 64-key chunks of P, published one by one
          mbar_wait(&p_full[c], t & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(o_tmem, pdesc + (uint64_t)(c * (ATT_TILE_BYTES >> 4)) + 2 * kk,
                     vdesc + (uint64_t)((c * 64 + kk * 16) * 128 >> 4), idesc_o, (c | kk) != 0);
          if (more) issue_s(t + 1, c);
        }
        if (has_tail) {
          mbar_wait(&p_full[4], t & 1);
          tc_fence_after();
          umma_f16(o_tmem, pdesc + (uint64_t)(4 * (ATT_TILE_BYTES >> 4)), vdesc + (uint64_t)(256 * 128 >> 4), idesc_o, 1);
          if (more) issue_s(t + 1, 4);
        }
        umma_commit(&o_full[t & 1]);
        if (more) {
          umma_commit(s_full);
          umma_commit(&q_empty[(t + 1) & 1]);     // S of tile t+1 was the only reader of its Q tile
        }
        if ((t & 1) == 0) umma_commit(&k_empty[(t >> 1) & 1]);   // S of the unit's second tile (just issued) was the last use of K
        else umma_commit(v_empty);                               // last P V of the unit
      }
    }
  } else if (warp < 8) {
    // ------------------------------ softmax + output ------------------------------
    const int half = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl = a.scale_log2;
    const int sw = r & 7;
    const uint32_t sP_row = smem_u32(sP) + half * 2 * ATT_TILE_BYTES + r * 128;   // this half's first chunk tile
    const int my_tail = half ? n_tail : 0;                                        // valid tail keys of this thread
    uint32_t qs = 0;
    float inv_prev = 0.f;
    __nv_bfloat16* dst_prev = nullptr;
    // O accumulator of tile `t` (columns 288 + 64 (t & 1)): this thread's 32 channels, scaled by 1 / row sum, straight to
    // global memory (64 contiguous bytes per row); frees the accumulator for tile t + 2.
    auto read_out = [&](uint32_t t, float inv, __nv_bfloat16* dst) {
      mbar_wait(&o_full[t & 1], (t >> 1) & 1);
      tc_fence_after();
      uint32_t w[32];
      tmem_ld32(tmem_O + (t & 1) * 64 + lane_off + half * 32, w);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&o_empty[t & 1]);
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 o4;
        o4.x = pack_bf16x2(__uint_as_float(w[i]) * inv, __uint_as_float(w[i + 1]) * inv);
        o4.y = pack_bf16x2(__uint_as_float(w[i + 2]) * inv, __uint_as_float(w[i + 3]) * inv);
        o4.z = pack_bf16x2(__uint_as_float(w[i + 4]) * inv, __uint_as_float(w[i + 5]) * inv);
        o4.w = pack_bf16x2(__uint_as_float(w[i + 6]) * inv, __uint_as_float(w[i + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + i) = o4;
      }
    };
    for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x) {
      const int img = unit / a.heads, head = unit - img * a.heads;
      for (int qt = 0; qt < 2; ++qt, ++qs) {
        float* my_x = xch + (qs & 1) * 512 + half * 128 + r;
        const float* peer_x = xch + (qs & 1) * 512 + (half ^ 1) * 128 + r;
        mbar_wait(s_full, qs & 1);
        tc_fence_after();
        const uint32_t s_addr = tmem_S + lane_off + half * 128;
        // ---- the thread's 128 scores in ONE TMEM round trip (two x64 loads in flight), kept in registers for both
        //      passes; only the 16 tail columns are read a second time ----
        uint32_t v0[64], v1[64], vt[16];
        tmem_ld64(s_addr, v0);
        tmem_ld64(s_addr + 64, v1);
        if (my_tail > 0) tmem_ld16(s_addr + 128, vt);
        tmem_ld_wait();
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          m0 = fmaxf(m0, __uint_as_float(v0[i]));
          m1 = fmaxf(m1, __uint_as_float(v0[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(v1[i]));
          m3 = fmaxf(m3, __uint_as_float(v1[i + 1]));
        }
        float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        if (my_tail > 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, i < my_tail ? __uint_as_float(vt[i]) : -INFINITY);
        }
        my_x[0] = mx;
        named_bar_sync(1 + quarter, 64);              // the two warps that share these 32 rows
        const float m = fmaxf(mx, peer_x[0]);
        const float2 msc = make_float2(-m * sl, -m * sl), sl2 = make_float2(sl, sl);
        // ---- pass 2: p = 2^(s*sl - m*sl) -> bf16 P (two 64-key chunk tiles per half) + row sum ----
        float2 rsum = make_float2(0.f, 0.f), rsum2 = make_float2(0.f, 0.f);
        auto exp64 = [&](const uint32_t(&v)[64], uint32_t row_addr) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {               // 8 keys = one 16-byte piece of the 128-byte P row
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = u * 8 + 2 * e;
              const float2 x = ffma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sl2, msc);
              const float2 pp = make_float2(ex2_approx(x.x), ex2_approx(x.y));
              if (e & 1) rsum2 = fadd2(rsum2, pp); else rsum = fadd2(rsum, pp);
              pk[e] = pack_bf16x2(pp.x, pp.y);
            }
            st_shared_v4(row_addr + ((u ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
          }
        };
        exp64(v0, sP_row);
        fence_proxy_async_smem();                     // chunk complete for this row; its S slice is dead
        tc_fence_before();
        mbar_arrive(&p_full[half * 2]);
        exp64(v1, sP_row + ATT_TILE_BYTES);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_full[half * 2 + 1]);
        rsum = fadd2(rsum, rsum2);
        if (has_tail && half == 1) {
          tmem_ld16(s_addr + 128, vt);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int i = 2 * e;
            const float2 x = ffma2(make_float2(__uint_as_float(vt[i]), __uint_as_float(vt[i + 1])), sl2, msc);
            float2 pp = make_float2(ex2_approx(x.x), ex2_approx(x.y));
            pp.x = i < my_tail ? pp.x : 0.f;
            pp.y = i + 1 < my_tail ? pp.y : 0.f;
            rsum = fadd2(rsum, pp);
            pk[e] = pack_bf16x2(pp.x, pp.y);
          }
          const uint32_t row_addr = smem_u32(sP) + 4 * ATT_TILE_BYTES + r * 128;
          st_shared_v4(row_addr + ((0 ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(row_addr + ((1 ^ sw) << 4), pk[4], pk[5], pk[6], pk[7]);
          fence_proxy_async_smem();
        }
        if (has_tail && half == 1) {                  // the tail chunk belongs to the upper half's 128 threads
          tc_fence_before();
          mbar_arrive(&p_full[4]);
        }
        // ---- row sum exchange; O of the PREVIOUS tile is read out now, while this tile's P V (and the next tile's S)
        //      drain through the tensor pipe ----
        my_x[256] = rsum.x + rsum.y;
        named_bar_sync(1 + quarter, 64);
        const float inv = 1.0f / (rsum.x + rsum.y + peer_x[256]);
        if (qs > 0) read_out(qs - 1, inv_prev, dst_prev);
        inv_prev = inv;
        dst_prev = a.ctx + ((long long)img * N + qt * 128 + r) * a.ctx_ld + head * 64 + half * 32;
      }
    }
    if (qs > 0) read_out(qs - 1, inv_prev, dst_prev);
  } else if (has_tail) {
    // ------------------------------ tail query rows on the CUDA cores (warps 10, 11) ------------------------------
    const int tw = warp - 10;                 // 0 / 1
    const int tid = tw * 32 + lane;           // 0..63
    uint32_t us = 0;
    for (int unit = blockIdx.x; unit < a.units; unit += gridDim.x, ++us) {
      const int img = unit / a.heads, head = unit - img * a.heads;
      const uint32_t kb = us & 1;
      const uint32_t kbase = smem_u32(sK + kb * ATR_KBUF), vbase = smem_u32(sV);
      for (int tr = 0; tr < n_tail; ++tr) {
        const long long qrow = (long long)img * N + 256 + tr;
        // the query row straight into registers (every lane reads the same 128 bytes: one L1 line, broadcast)
        float q[64];
        {
          const uint4* q4 = reinterpret_cast<const uint4*>(a.qkv + qrow * (3LL * a.D) + head * 64);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 raw = q4[i];
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(h2[e]);
              q[i * 8 + 2 * e] = f.x * a.scale_log2;     // fold log2(e) / sqrt(d) into the query
              q[i * 8 + 2 * e + 1] = f.y * a.scale_log2;
            }
          }
        }
        if (tr == 0) mbar_wait(&k_full[kb], (us >> 1) & 1);
        // scores of keys tid, tid + 64, ... (5 per thread), two independent accumulators per key
        float sc[5];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int k = tid + 64 * j;
          float acc = -INFINITY;
          if (k < N) {
            const uint32_t row = kbase + k * 128;   // tiles are contiguous: key k lives at row k of the 128-byte rows
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int pc = 0; pc < 8; ++pc) {
              const float4 raw = ld_shared_v4f(row + ((pc ^ (k & 7)) << 4));
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h2[e]);
                a0 = fmaf(q[pc * 8 + 2 * e], f.x, a0);
                a1 = fmaf(q[pc * 8 + 2 * e + 1], f.y, a1);
              }
            }
            acc = a0 + a1;
          }
          sc[j] = acc;
          mx = fmaxf(mx, acc);
        }
        mx = warp_max(mx);
        named_bar_sync(6, 64);                // the previous row's readers of tred / tp / tpart are done
        if (lane == 0) tred[tw] = mx;
        named_bar_sync(6, 64);
        mx = fmaxf(tred[0], tred[1]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int k = tid + 64 * j;
          const float e = k < N ? ex2_approx(sc[j] - mx) : 0.f;
          if (k < 272) tp[k] = e;
          sum += e;
        }
        sum = warp_sum(sum);
        if (lane == 0) tred[2 + tw] = sum;
        if (tr == 0) mbar_wait(v_full, us & 1);
        named_bar_sync(6, 64);
        const float inv = 1.0f / (tred[2] + tred[3]);
        // O[c] = sum_k p[k] V[k][c]: lane -> channel pair (2 lane, 2 lane + 1), warp tw -> keys k = tw (mod 2);
        // four independent accumulator pairs keep the FMA chains short
        float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t vcol = ((lane & 3) << 2), vpc = lane >> 2;
        int k = tw;
        for (; k + 6 < N; k += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int kk = k + 2 * u;
            uint32_t raw;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(raw) : "r"(vbase + kk * 128 + ((vpc ^ (kk & 7)) << 4) + vcol));
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
            const float pk = tp[kk];
            o0[u] = fmaf(pk, f.x, o0[u]);
            o1[u] = fmaf(pk, f.y, o1[u]);
          }
        }
        for (; k < N; k += 2) {
          uint32_t raw;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(raw) : "r"(vbase + k * 128 + ((vpc ^ (k & 7)) << 4) + vcol));
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
          const float pk = tp[k];
          o0[0] = fmaf(pk, f.x, o0[0]);
          o1[0] = fmaf(pk, f.y, o1[0]);
        }
        tpart[tw * 64 + 2 * lane] = (o0[0] + o0[1]) + (o0[2] + o0[3]);
        tpart[tw * 64 + 2 * lane + 1] = (o1[0] + o1[1]) + (o1[2] + o1[3]);
        named_bar_sync(6, 64);
        if (tw == 0) {
          const float r0 = (tpart[2 * lane] + tpart[64 + 2 * lane]) * inv;
          const float r1 = (tpart[2 * lane + 1] + tpart[64 + 2 * lane + 1]) * inv;
          *reinterpret_cast<uint32_t*>(a.ctx + qrow * a.ctx_ld + head * 64 + 2 * lane) = pack_bf16x2(r0, r1);
        }
      }
      // this warp is done with the unit's K and V tiles
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&k_empty[kb]);
        mbar_arrive(v_empty);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

// Last query rows of an image when tokens % 128 <= ATT_TAIL_MAX (e.g. the 257th token at 224x224): one 128-thread block
// per (image, head, row) on the CUDA cores instead of a 128-row tensor-core tile that would be >87% padding.
// Keys are split over the four warps (scores -> shared memory, block-wide max / sum), then thread (w, l) accumulates
// output channels (2l, 2l+1) over the keys k = w (mod 4) and the four partial sums are combined through shared memory.
constexpr int ATTT_THREADS = 128;
__global__ void __launch_bounds__(ATTT_THREADS) attn_tail_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                 __nv_bfloat16* __restrict__ ctx, int images, int tokens,
                                                                 int heads, int D, long long ctx_ld, int first_row,
                                                                 float scale_log2) {
  extern __shared__ float sp[];  // [tokens] scores + [8] reductions + [4][64] partial outputs
  float* red = sp + tokens;
  float* part = red + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tail = tokens - first_row;
  const int qi = first_row + (int)(blockIdx.x % tail);
  const int head = (int)((blockIdx.x / tail) % heads);
  const int img = (int)(blockIdx.x / (tail * heads));
  const long long ld = 3LL * D;
  const __nv_bfloat16* base = qkv + (long long)img * tokens * ld;
  float qr[64];
  {
    const uint4* q4 = reinterpret_cast<const uint4*>(base + (long long)qi * ld + head * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 t = q4[i];
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h2[e]);
        qr[i * 8 + 2 * e] = f.x;
        qr[i * 8 + 2 * e + 1] = f.y;
      }
    }
  }
  float mx = -INFINITY;
  for (int k = threadIdx.x; k < tokens; k += ATTT_THREADS) {
    const uint4* k4 = reinterpret_cast<const uint4*>(base + (long long)k * ld + D + head * 64);
    uint4 kv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) kv[i] = k4[i];
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&kv[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h2[e]);
        acc = fmaf(qr[i * 8 + 2 * e], f.x, acc);
        acc = fmaf(qr[i * 8 + 2 * e + 1], f.y, acc);
      }
    }
    acc *= scale_log2;
    sp[k] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float sum = 0.f;
  for (int k = threadIdx.x; k < tokens; k += ATTT_THREADS) {
    const float e = ex2_approx(sp[k] - mx);
    sp[k] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[4 + warp] = sum;
  __syncthreads();
  const float inv = 1.0f / (red[4] + red[5] + red[6] + red[7]);
  float o0 = 0.f, o1 = 0.f;
  const __nv_bfloat16* vb = base + 2 * D + head * 64 + 2 * lane;
  for (int k = warp; k < tokens; k += 4) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vb + (long long)k * ld));
    const float pk = sp[k];
    o0 = fmaf(pk, f.x, o0);
    o1 = fmaf(pk, f.y, o1);
  }
  part[warp * 64 + 2 * lane] = o0;
  part[warp * 64 + 2 * lane + 1] = o1;
  __syncthreads();
  if (warp == 0) {
    o0 = part[2 * lane] + part[64 + 2 * lane] + part[128 + 2 * lane] + part[192 + 2 * lane];
    o1 = part[2 * lane + 1] + part[64 + 2 * lane + 1] + part[128 + 2 * lane + 1] + part[192 + 2 * lane + 1];
    *reinterpret_cast<uint32_t*>(ctx + ((long long)img * tokens + qi) * ctx_ld + head * 64 + 2 * lane) = pack_bf16x2(o0 * inv, o1 * inv);
  }
}

// fp32 attention on the CUDA cores (parity mode): one warp per (image, head, query row).
constexpr int ATTF_WARPS = 8;
__global__ void __launch_bounds__(ATTF_WARPS * 32) attn_f32_kernel(const float* __restrict__ qkv, float* __restrict__ ctx,
                                                                   int images, int tokens, int heads, int D,
                                                                   long long ctx_ld, long long ctx_plane) {
  extern __shared__ float sp[];  // [ATTF_WARPS][tokens]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * ATTF_WARPS + warp;
  const long long total = (long long)images * heads * tokens;
  if (gw >= total) return;
  const int qi = (int)(gw % tokens);
  const int head = (int)((gw / tokens) % heads);
  const int img = (int)(gw / ((long long)tokens * heads));
  const long long ld = 3LL * D;
  const float* base = qkv + (long long)img * tokens * ld;
  const float* q = base + (long long)qi * ld + head * 64;
  float* p = sp + (long long)warp * tokens;
  float qr[64];
#pragma unroll
  for (int i = 0; i < 64; i += 4) {
    const float4 t = *reinterpret_cast<const float4*>(q + i);
    qr[i] = t.x; qr[i + 1] = t.y; qr[i + 2] = t.z; qr[i + 3] = t.w;
  }
  float mx = -INFINITY;
  for (int k = lane; k < tokens; k += 32) {
    const float* kr = base + (long long)k * ld + D + head * 64;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      const float4 t = *reinterpret_cast<const float4*>(kr + i);
      acc = fmaf(qr[i], t.x, acc);
      acc = fmaf(qr[i + 1], t.y, acc);
      acc = fmaf(qr[i + 2], t.z, acc);
      acc = fmaf(qr[i + 3], t.w, acc);
    }
    acc *= 0.125f;
    p[k] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int k = lane; k < tokens; k += 32) {
    const float e = expf(p[k] - mx);
    p[k] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.0f / sum;
  float o0 = 0.f, o1 = 0.f;
  const float* vb = base + 2 * D + head * 64;
  for (int k = 0; k < tokens; ++k) {
    const float pk = p[k];
    o0 = fmaf(pk, vb[(long long)k * ld + lane], o0);
    o1 = fmaf(pk, vb[(long long)k * ld + 32 + lane], o1);
  }
  const long long orow = ((long long)img * tokens + qi) * ctx_ld + head * 64;
  store_val(ctx, 1, orow + lane, ctx_plane, o0 * inv);
  store_val(ctx, 1, orow + 32 + lane, ctx_plane, o1 * inv);
}

}  // namespace vt
