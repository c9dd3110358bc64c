// attn_pp_kernel: DinoV2 attention at 257 tokens (224 x 224 images; HF:199-235 softmax(Q K^T / 8) V, head_dim 64) with TWO independent
// groups per CTA, each walking its own (image, head) units ("ping-pong" without coupling).
//
// attn_row_kernel (vt_attn.cuh) keeps one score tile in tensor memory and all eight softmax warps walk the chain
// S ready -> TMEM load -> max -> exp -> P published -> P V -> O read-out in lockstep, so the MUFU pipe (the floor of this head
// dimension: one exponential per 256 tensor FLOPs) idles ~65 % of the time: 7.0 k cycles per query tile against a floor of 2.1 k.
// Here a group = 4 softmax warps (one thread per query row) + its own MMA-issuing thread + two warps for the 257th query row, with
// its own K / V / Q buffers and its own 256-column region of tensor memory; group g takes the CTA's units g, g + 2, ...:
//   tensor memory: S of the tile (128 x 256 fp32).  P never goes to shared memory: the thread writes its row's bf16 probabilities
//     back into columns of S it has already consumed (tcgen05.st) and O = P V reads them as the A operand straight from tensor
//     memory (tcgen05.mma [d], [a_tmem], b_desc); O accumulates into the region's columns 192..255, which are dead because the
//     thread keeps the scores of keys 128..255 in registers after the max pass.
//   the 257th key (the reason S would need 272 columns) is handled by the row's own thread on the CUDA cores: one 64-wide dot
//     product for its score (computed for the NEXT tile while this tile's last P V chunk drains), one rank-1 update at read-out.
//   shared memory per group: K and V single-buffered (K is released after the S MMA of the unit's second tile, V after its last
//     P V: the next unit's tiles land while the current one is still in its softmax), Q per tile, the query box of row 256.
//   warp 8 = TMA producer for both groups (polls with mbarrier.test_wait: try_wait may suspend the thread, and a thread that
//     serves several barriers in turn then serves them at that cadence).
// Whole warpgroups per role + setmaxnreg: the softmax threads hold 128 scores + 32 packed probabilities in registers.
// What the first versions taught (tools/attn_trace.py, profiles/r02_attn_pp_timeline.txt): with both tiles of ONE unit in flight the
// two tiles share the unit's K / V buffers, so whichever runs ahead waits for the other to release them; one MMA thread polling
// both tiles spent ~1 k cycles per action; the 257th query row took 10.7 k cycles per unit with volatile shared-memory loads.
#pragma once
#include "vt_attn.cuh"

namespace vt {

constexpr int APP_THREADS = 512;   // per group g: softmax warps 4g..4g+3, MMA warp 9+g, tail-query warps 11+2g, 12+2g; warp 8 TMA, warp 15 idle
constexpr int APP_KBUF = 2 * ATT_TILE_BYTES + 2048;            // two 128-row tiles + one 16-row tail box (row 256)
constexpr int APP_GROUP_BYTES = 2 * APP_KBUF + 2 * ATT_TILE_BYTES;   // per group: K | V | Q tile 0 | Q tile 1
constexpr int APP_SMEM_QT = 2 * APP_GROUP_BYTES;               // [group][2 KB]: the 16-row box that holds query row 256
constexpr int APP_SMEM_BAR = APP_SMEM_QT + 2 * 2048;           // 2 x 14 mbarriers + the TMEM slot
constexpr int APP_SMEM_TAIL = APP_SMEM_BAR + 512;              // per group: q[64], p[272], red[16], part[2][64] floats
constexpr int APP_TAIL_FLOATS = 64 + 272 + 16 + 128;
constexpr int APP_SMEM_BYTES = 1024 + APP_SMEM_TAIL + 2 * APP_TAIL_FLOATS * 4 + 16;
static_assert(APP_SMEM_BYTES <= 227 * 1024, "attn_pp_kernel shared memory");
static_assert(APP_GROUP_BYTES % 1024 == 0 && APP_KBUF % 1024 == 0, "128B-swizzled tiles need 1024-byte alignment");
constexpr int APP_REGS_CTRL = 56, APP_REGS_SOFTMAX = 200;    // 56 * 256 + 200 * 256 = 128 * 512: the pool is what the launch allocated
static_assert(APP_REGS_CTRL * 256 + APP_REGS_SOFTMAX * 256 <= 128 * APP_THREADS, "setmaxnreg can only redistribute the registers the launch allocated (128 per thread at 16 warps)");
// columns inside a group's 256-column region of tensor memory
constexpr int APP_COL_O = 192;
__host__ __device__ constexpr int app_p_col(int chunk) { return chunk < 2 ? chunk * 32 : 128 + (chunk - 2) * 32; }
// barriers of one group
enum : int { APB_K_FULL = 0, APB_K_EMPTY, APB_V_FULL, APB_V_EMPTY, APB_Q_FULL /*2*/, APB_Q_EMPTY = 6 /*2*/, APB_S_FULL = 8, APB_P_FULL /*4*/,
             APB_O_FULL = 13, APB_REGION_FREE, APB_COUNT };

// developer instrumentation (debug-knobs builds, tools/attn_trace.py): (tag, clock64) pairs of CTA 0 in four slices of vt_dbg_ts
constexpr int APP_TS_SLICE = VT_DBG_TS / 4;
__device__ __forceinline__ void app_stamp(int slice, int& n, int tag) {
#if VT_DEBUG_KNOBS
  if (blockIdx.x == 0 && n + 1 < APP_TS_SLICE) {
    vt_dbg_ts[slice * APP_TS_SLICE + n] = tag;
    vt_dbg_ts[slice * APP_TS_SLICE + n + 1] = clock64();
    n += 2;
  }
#endif
}

__global__ void __launch_bounds__(APP_THREADS, 1) attn_pp_kernel(const __grid_constant__ AttnRowArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + APP_SMEM_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * APB_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.tokens;   // 257

  if (warp == 9) {
    if (lane == 0) {
      for (int g = 0; g < 2; ++g) {
        uint64_t* B = bars + g * APB_COUNT;
        mbar_init(&B[APB_K_FULL], 1);
        mbar_init(&B[APB_K_EMPTY], 1 + 4 + 2);     // S of tile 1 committed, 4 softmax warps (tail-key dot of tile 1), 2 tail warps
        mbar_init(&B[APB_V_FULL], 1);
        mbar_init(&B[APB_V_EMPTY], 1 + 4 + 2);     // P V of tile 1 committed, 4 softmax warps (read-out of tile 1), 2 tail warps
        for (int t = 0; t < 2; ++t) {
          mbar_init(&B[APB_Q_FULL + t], 1);
          mbar_init(&B[APB_Q_EMPTY + t], 1 + 4);   // S committed, 4 softmax warps (tail-key dot)
        }
        mbar_init(&B[APB_S_FULL], 1);
        for (int c = 0; c < 4; ++c) mbar_init(&B[APB_P_FULL + c], 128);
        mbar_init(&B[APB_O_FULL], 1);
        mbar_init(&B[APB_REGION_FREE], 128);
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
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int n_units = (int)blockIdx.x < a.units ? (a.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  // group g takes this CTA's units g, g + 2, ...; its j-th unit:
  auto unit_of = [&](int g, int j) { return (int)blockIdx.x + (2 * j + g) * (int)gridDim.x; };
  auto units_of = [&](int g) { return (n_units - g + 1) >> 1; };

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(APP_REGS_CTRL));
    if (warp == 8) {
      // ------------------------------ TMA producer (both groups; polls with test_wait, never blocks on one group) ------------------------------
      if (lane == 0) {
        int jj[2] = {0, 0}, op[2] = {0, 0};      // next load of each group: unit index, 0 = K (+ tail boxes), 1 / 2 = Q tile 0 / 1, 3 = V
        int live = (units_of(0) > 0) + (units_of(1) > 0);
        int ts_n = 0;
        long long t_progress = clock64();
        while (live > 0) {
          if (clock64() - t_progress > 4000000000LL) __trap();   // a protocol bug must trap, never hang the GPU box
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (jj[g] >= units_of(g)) continue;
            uint64_t* B = bars + g * APB_COUNT;
            uint8_t* base = smem + g * APP_GROUP_BYTES;
            const int j = jj[g];
            const uint32_t pe = (uint32_t)(j & 1) ^ 1u;
            const int unit = unit_of(g, j);
            const int img = unit / a.heads, head = unit - img * a.heads;
            const int row0 = img * N;
            if (op[g] == 0) {
              if (!mbar_test_wait(&B[APB_K_EMPTY], pe)) continue;
              mbar_arrive_expect_tx(&B[APB_K_FULL], 2 * ATT_TILE_BYTES + 2048 + 2048);
              tma_load_2d(base, &a.tm, &B[APB_K_FULL], a.D + head * 64, row0);
              tma_load_2d(base + ATT_TILE_BYTES, &a.tm, &B[APB_K_FULL], a.D + head * 64, row0 + 128);
              tma_load_2d(base + 2 * ATT_TILE_BYTES, &a.tm16, &B[APB_K_FULL], a.D + head * 64, row0 + 256);
              tma_load_2d(smem + APP_SMEM_QT + g * 2048, &a.tm16, &B[APB_K_FULL], head * 64, row0 + 256);   // query row 256
              if (g == 0) app_stamp(3, ts_n, 60);
            } else if (op[g] < 3) {
              const int t = op[g] - 1;
              if (!mbar_test_wait(&B[APB_Q_EMPTY + t], pe)) continue;
              mbar_arrive_expect_tx(&B[APB_Q_FULL + t], ATT_TILE_BYTES);
              tma_load_2d(base + 2 * APP_KBUF + t * ATT_TILE_BYTES, &a.tm, &B[APB_Q_FULL + t], head * 64, row0 + t * 128);
            } else {
              if (!mbar_test_wait(&B[APB_V_EMPTY], pe)) continue;
              uint8_t* vbuf = base + APP_KBUF;
              mbar_arrive_expect_tx(&B[APB_V_FULL], 2 * ATT_TILE_BYTES + 2048);
              tma_load_2d(vbuf, &a.tm, &B[APB_V_FULL], 2 * a.D + head * 64, row0);
              tma_load_2d(vbuf + ATT_TILE_BYTES, &a.tm, &B[APB_V_FULL], 2 * a.D + head * 64, row0 + 128);
              tma_load_2d(vbuf + 2 * ATT_TILE_BYTES, &a.tm16, &B[APB_V_FULL], 2 * a.D + head * 64, row0 + 256);
              if (g == 0) app_stamp(3, ts_n, 61);
            }
            t_progress = clock64();
            if (++op[g] == 4) {
              op[g] = 0;
              if (++jj[g] >= units_of(g)) --live;
            }
          }
        }
      }
    } else if (warp == 9 || warp == 10) {
      // ------------------------------ UMMA issuer of group g = warp - 9 ------------------------------
      const int g = warp - 9;
      if (lane == 0 && units_of(g) > 0) {
        uint64_t* B = bars + g * APB_COUNT;
        uint8_t* base = smem + g * APP_GROUP_BYTES;
        const uint32_t idesc_s = umma_idesc(UMMA_FMT_BF16, 256);
        const uint32_t idesc_o = umma_idesc(UMMA_FMT_BF16, 64, 0, 1);
        const uint32_t region = tmem_base + g * 256;
        const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(base));
        const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(base + APP_KBUF));
        int ts_n = g * (APP_TS_SLICE / 2);
        // Group 1 starts half a chain behind group 0 (its first S waits for group 0's last P chunk): two groups that walk the chain
        // in lockstep want the MUFU pipe at the same time and leave it idle at the same time; nothing couples them afterwards.
        if (g == 1) mbar_wait(&bars[APB_P_FULL + 1], 0);
        const int nu = units_of(g);
        for (int j = 0; j < nu; ++j) {
          const uint32_t pj = j & 1;
#pragma unroll 1
          for (int t = 0; t < 2; ++t) {
            const int tt = 2 * j + t;
            if (t == 0) mbar_wait(&B[APB_K_FULL], pj);
            mbar_wait(&B[APB_Q_FULL + t], pj);
            if (tt > 0) mbar_wait(&B[APB_REGION_FREE], (tt - 1) & 1);
            tc_fence_after();
            const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(base + 2 * APP_KBUF + t * ATT_TILE_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(region, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
            umma_commit(&B[APB_S_FULL]);
            umma_commit(&B[APB_Q_EMPTY + t]);
            if (t == 1) umma_commit(&B[APB_K_EMPTY]);
            app_stamp(2, ts_n, 20 + g);
            if (t == 0) mbar_wait(&B[APB_V_FULL], pj);
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
              const int c = (i + 2) & 3;      // chunks are published in the order 2, 3, 0, 1
              mbar_wait(&B[APB_P_FULL + c], tt & 1);
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16_ts(region + APP_COL_O, region + app_p_col(c) + 8 * kk, vdesc + (uint64_t)((c * 64 + kk * 16) * 128 >> 4), idesc_o, (i | kk) != 0);
              app_stamp(2, ts_n, 30 + 10 * g + c);
            }
            umma_commit(&B[APB_O_FULL]);
            if (t == 1) umma_commit(&B[APB_V_EMPTY]);
          }
        }
      }
    } else if (warp < 15) {
      // ------------------------------ the 257th query row on the CUDA cores: warps 11, 12 (group 0), 13, 14 (group 1) ------------------------------
      // Plain (non-volatile) shared-memory loads through generic pointers: the compiler may overlap them and still keeps them behind
      // the mbarrier waits (asm volatile with a memory clobber); the volatile ld.shared helper serialises one shared-memory round
      // trip per 16 bytes (attn_row_kernel's version of this row takes 10.7 k cycles per unit).
      const int g = (warp - 11) >> 1;
      const int tw = (warp - 11) & 1;
      const int tid = tw * 32 + lane;           // 0..63
      uint64_t* B = bars + g * APB_COUNT;
      const uint8_t* kbase = smem + g * APP_GROUP_BYTES;
      const uint8_t* vbase = kbase + APP_KBUF;
      const uint8_t* qt = smem + APP_SMEM_QT + g * 2048;
      float* tq = reinterpret_cast<float*>(smem + APP_SMEM_TAIL) + g * APP_TAIL_FLOATS;   // [64]
      float* tp = tq + 64;                                                                 // [272]
      float* tred = tp + 272;                                                              // [16]
      float* tpart = tred + 16;                                                            // [2][64]
      int ts_n = APP_TS_SLICE / 2;
      const bool ts_on = kDbg && g == 0 && tid == 0;
      const int nu = units_of(g);
      for (int j = 0; j < nu; ++j) {
        const int unit = unit_of(g, j);
        const int img = unit / a.heads, head = unit - img * a.heads;
        const uint32_t pj = j & 1;
        mbar_wait(&B[APB_K_FULL], pj);
        if (ts_on) app_stamp(3, ts_n, 50);
        {   // q_256 * log2(e) / 8 as fp32 in shared memory (the previous unit's readers left at its last barrier)
          const uint32_t raw = *reinterpret_cast<const uint32_t*>(qt + (tid & 31) * 4);   // row 0 of the box: no swizzle offset
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
          if (tid < 32) {
            tq[2 * tid] = f.x * a.scale_log2;
            tq[2 * tid + 1] = f.y * a.scale_log2;
          }
        }
        named_bar_sync(6 + g, 64);
        float mx = -INFINITY;
#pragma unroll 1
        for (int jk = 0; jk < 5; ++jk) {        // (56 registers per thread here: one key at a time, scores parked in tp)
          const int k = tid + 64 * jk;
          float acc = -INFINITY;
          if (k < N) {
            const uint8_t* row = kbase + k * 128;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int pc = 0; pc < 8; ++pc) {
              const float4 raw = *reinterpret_cast<const float4*>(row + ((pc ^ (k & 7)) << 4));
              const float4 q0 = *reinterpret_cast<const float4*>(tq + pc * 8);
              const float4 q1 = *reinterpret_cast<const float4*>(tq + pc * 8 + 4);
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
              const float2 f0 = __bfloat1622float2(h2[0]), f1 = __bfloat1622float2(h2[1]);
              const float2 f2 = __bfloat1622float2(h2[2]), f3 = __bfloat1622float2(h2[3]);
              a0 = fmaf(q0.x, f0.x, a0);
              a1 = fmaf(q0.y, f0.y, a1);
              a2 = fmaf(q0.z, f1.x, a2);
              a3 = fmaf(q0.w, f1.y, a3);
              a0 = fmaf(q1.x, f2.x, a0);
              a1 = fmaf(q1.y, f2.y, a1);
              a2 = fmaf(q1.z, f3.x, a2);
              a3 = fmaf(q1.w, f3.y, a3);
            }
            acc = (a0 + a1) + (a2 + a3);
            tp[k] = acc;
          }
          mx = fmaxf(mx, acc);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&B[APB_K_EMPTY]);      // K and the query box have been read
        mx = warp_max(mx);
        if (lane == 0) tred[tw] = mx;
        named_bar_sync(6 + g, 64);
        mx = fmaxf(tred[0], tred[1]);
        float sum = 0.f;
#pragma unroll
        for (int jk = 0; jk < 5; ++jk) {
          const int k = tid + 64 * jk;
          if (k < N) {
            const float e = ex2_approx(tp[k] - mx);     // the thread's own keys
            tp[k] = e;
            sum += e;
          }
        }
        sum = warp_sum(sum);
        if (lane == 0) tred[8 + tw] = sum;
        mbar_wait(&B[APB_V_FULL], pj);
        named_bar_sync(6 + g, 64);
        if (ts_on) app_stamp(3, ts_n, 51);
        const float inv = 1.0f / (tred[8] + tred[9]);
        // O[c] = sum_k p[k] V[k][c]: lane -> channels (2 lane, 2 lane + 1), warp tw -> keys k = tw (mod 2)
        float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t vcol = ((lane & 3) << 2), vpc = lane >> 2;
        int k = tw;
        for (; k + 6 < N; k += 8) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int kk = k + 2 * u;
            const uint32_t raw = *reinterpret_cast<const uint32_t*>(vbase + kk * 128 + ((vpc ^ (kk & 7)) << 4) + vcol);
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
            const float pk = tp[kk];
            o0[u] = fmaf(pk, f.x, o0[u]);
            o1[u] = fmaf(pk, f.y, o1[u]);
          }
        }
        for (; k < N; k += 2) {
          const uint32_t raw = *reinterpret_cast<const uint32_t*>(vbase + k * 128 + ((vpc ^ (k & 7)) << 4) + vcol);
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
          const float pk = tp[k];
          o0[0] = fmaf(pk, f.x, o0[0]);
          o1[0] = fmaf(pk, f.y, o1[0]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&B[APB_V_EMPTY]);      // V has been read
        tpart[tw * 64 + 2 * lane] = (o0[0] + o0[1]) + (o0[2] + o0[3]);
        tpart[tw * 64 + 2 * lane + 1] = (o1[0] + o1[1]) + (o1[2] + o1[3]);
        named_bar_sync(6 + g, 64);
        if (tw == 0) {
          const float r0 = (tpart[2 * lane] + tpart[64 + 2 * lane]) * inv;
          const float r1 = (tpart[2 * lane + 1] + tpart[64 + 2 * lane + 1]) * inv;
          const long long qrow = (long long)img * N + 256;
          *reinterpret_cast<uint32_t*>(a.ctx + qrow * a.ctx_ld + head * 64 + 2 * lane) = pack_bf16x2(r0, r1);
        }
        if (ts_on) app_stamp(3, ts_n, 52);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(APP_REGS_SOFTMAX));
    // ------------------------------ softmax + output: group g = warp / 4, one thread per query row ------------------------------
    const int g = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    uint64_t* B = bars + g * APB_COUNT;
    const uint8_t* base = smem + g * APP_GROUP_BYTES;
    const uint32_t region = tmem_base + g * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sl = a.scale_log2;
    const float2 sl2 = make_float2(sl, sl);
    const int nu = units_of(g);
    const int n_tiles = 2 * nu;
    // Score of the 257th key for tile tt = 2 j + t: q_r . k_256 from shared memory.  Computed for the NEXT tile while this tile's
    // last P V chunk drains, so it is off the tile's critical chain; releases the Q tile (and, for t = 1, the unit's K).
    auto tail_score = [&](int tt) -> float {
      const int t = tt & 1;
      const uint32_t pj = (tt >> 1) & 1;
      mbar_wait(&B[APB_Q_FULL + t], pj);
      mbar_wait(&B[APB_K_FULL], pj);
      const uint8_t* qrow = base + 2 * APP_KBUF + t * ATT_TILE_BYTES + r * 128;
      const uint8_t* krow = base + 2 * ATT_TILE_BYTES;   // row 256 = row 0 of the tail box
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int pc = 0; pc < 8; ++pc) {
        const float4 qr = *reinterpret_cast<const float4*>(qrow + ((pc ^ (r & 7)) << 4));
        const float4 kr = *reinterpret_cast<const float4*>(krow + (pc << 4));
        const __nv_bfloat162* qh = reinterpret_cast<const __nv_bfloat162*>(&qr);
        const __nv_bfloat162* kh = reinterpret_cast<const __nv_bfloat162*>(&kr);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 qf = __bfloat1622float2(qh[e]), kf = __bfloat1622float2(kh[e]);
          acc[2 * e] = fmaf(qf.x, kf.x, acc[2 * e]);
          acc[2 * e + 1] = fmaf(qf.y, kf.y, acc[2 * e + 1]);
        }
      }
      const float res = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&B[APB_Q_EMPTY + t]);
        if (t == 1) mbar_arrive(&B[APB_K_EMPTY]);
      }
      return res;
    };
    float st = n_tiles > 0 ? tail_score(0) : 0.f;
    int ts_n = 0;
    const bool ts_on = kDbg && quarter == 0 && lane == 0;
    for (int tt = 0; tt < n_tiles; ++tt) {
      const int t = tt & 1, j = tt >> 1;
      const int unit = unit_of(g, j);
      const int img = unit / a.heads, head = unit - img * a.heads;
      const uint32_t pj = j & 1;
      if (ts_on) app_stamp(g, ts_n, 0);
      mbar_wait(&B[APB_S_FULL], tt & 1);
      tc_fence_after();
      if (ts_on) app_stamp(g, ts_n, 1);
      // ---- pass 1: row maximum over the 256 scores in tensor memory; the scores of keys 128..255 stay in registers ----
      uint32_t x0[64], x1[64];
      tmem_ld64(region, x0);
      tmem_ld64(region + 64, x1);
      tmem_ld_wait();
      float m0 = st, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        m0 = fmaxf(m0, __uint_as_float(x0[i]));
        m1 = fmaxf(m1, __uint_as_float(x0[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(x1[i]));
        m3 = fmaxf(m3, __uint_as_float(x1[i + 1]));
      }
      tmem_ld64(region + 128, x0);
      tmem_ld64(region + 192, x1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        m0 = fmaxf(m0, __uint_as_float(x0[i]));
        m1 = fmaxf(m1, __uint_as_float(x0[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(x1[i]));
        m3 = fmaxf(m3, __uint_as_float(x1[i + 1]));
      }
      const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      if (ts_on) app_stamp(g, ts_n, 2);
      const float2 msc = make_float2(-m * sl, -m * sl);
      // ---- pass 2: p = 2^(s*sl - m*sl) as bf16 pairs written back into consumed columns of S; chunk order 2, 3, 0, 1 ----
      float2 rsum = make_float2(0.f, 0.f), rsum2 = make_float2(0.f, 0.f);
      uint32_t pk[32];
      auto exp64 = [&](const uint32_t(&v)[64]) {
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          const float2 x = ffma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sl2, msc);
          const float2 pp = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          if (i & 2) rsum2 = fadd2(rsum2, pp); else rsum = fadd2(rsum, pp);
          pk[i >> 1] = pack_bf16x2(pp.x, pp.y);
        }
      };
      auto publish = [&](int chunk) {
        tmem_st32(region + app_p_col(chunk), pk);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&B[APB_P_FULL + chunk]);
      };
      exp64(x0);                      // keys 128..191
      tmem_ld64(region, x0);          // re-read keys 0..63 while chunk 2 is published and chunk 3 is computed
      publish(2);
      if (ts_on) app_stamp(g, ts_n, 3);
      exp64(x1);                      // keys 192..255
      tmem_ld64(region + 64, x1);
      publish(3);
      if (ts_on) app_stamp(g, ts_n, 4);
      tmem_ld_wait();
      exp64(x0);
      publish(0);
      if (ts_on) app_stamp(g, ts_n, 5);
      exp64(x1);
      publish(1);
      if (ts_on) app_stamp(g, ts_n, 6);
      // the 257th key
      const float pt = ex2_approx(fmaf(st, sl, msc.x));
      const float inv = 1.0f / ((rsum.x + rsum.y) + (rsum2.x + rsum2.y) + pt);
      const float ptb = __bfloat162float(__float2bfloat16(pt));   // the MMA path multiplies bf16 probabilities
      const float st_next = tt + 1 < n_tiles ? tail_score(tt + 1) : 0.f;
      if (ts_on) app_stamp(g, ts_n, 7);
      // ---- O = P V from tensor memory + the rank-1 term of the 257th key, scaled, 128 contiguous bytes per row ----
      mbar_wait(&B[APB_V_FULL], pj);
      mbar_wait(&B[APB_O_FULL], tt & 1);
      tc_fence_after();
      if (ts_on) app_stamp(g, ts_n, 8);
      tmem_ld64(region + APP_COL_O, x0);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&B[APB_REGION_FREE]);
      {
        const uint8_t* vrow = base + APP_KBUF + 2 * ATT_TILE_BYTES;   // V row 256
        __nv_bfloat16* dst = a.ctx + ((long long)img * N + t * 128 + r) * a.ctx_ld + head * 64;
#pragma unroll
        for (int pc = 0; pc < 8; ++pc) {
          const float4 vr = *reinterpret_cast<const float4*>(vrow + (pc << 4));
          const __nv_bfloat162* vh = reinterpret_cast<const __nv_bfloat162*>(&vr);
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 vf = __bfloat1622float2(vh[e]);
            const float o_lo = fmaf(ptb, vf.x, __uint_as_float(x0[pc * 8 + 2 * e])) * inv;
            const float o_hi = fmaf(ptb, vf.y, __uint_as_float(x0[pc * 8 + 2 * e + 1])) * inv;
            w[e] = pack_bf16x2(o_lo, o_hi);
          }
          *reinterpret_cast<uint4*>(dst + pc * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      if (t == 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&B[APB_V_EMPTY]);
      }
      if (ts_on) app_stamp(g, ts_n, 9);
      st = st_next;
    }
  }

#if VT_DEBUG_KNOBS
  if (blockIdx.x == 0 && threadIdx.x == 0) vt_dbg_n = 4 * APP_TS_SLICE;
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

}  // namespace vt
