// attn_pp_kernel: DinoV2 attention at 257 tokens (224 x 224 images; HF:199-235 softmax(Q K^T / 8) V, head_dim 64) with the two
// 128-query tiles of an (image, head) unit in flight at the same time ("ping-pong").
//
// attn_row_kernel (vt_attn.cuh) keeps one score tile in tensor memory and all eight softmax warps walk the chain
// S ready -> TMEM load -> max -> exp -> P published -> P V -> O read-out in lockstep, so the MUFU pipe (the floor of this head
// dimension: one exponential per 256 tensor FLOPs) idles ~65 % of the time: 7.0 k cycles per query tile against a floor of 2.1 k.
// Here each tile has its OWN score buffer and its OWN four softmax warps (one thread per query row):
//   tensor memory (512 columns) = two regions of 256: S of the tile (128 x 256 fp32).  P never goes to shared memory: the thread
//     writes its row's bf16 probabilities back into columns of S it has already consumed (tcgen05.st) and O = P V reads them as
//     the A operand straight from tensor memory (tcgen05.mma with [a_tmem]); O accumulates into the region's columns 192..255,
//     which are dead because the thread keeps the scores of keys 128..255 in registers after the max pass.
//   the 257th key (the reason S would need 272 columns) is handled by the row's own thread on the CUDA cores: one 64-wide dot
//     product for its score, one rank-1 update of the output row at read-out.
//   shared memory = K and V double-buffered per unit (2 x 2 x 34 KB) and Q double-buffered per tile (4 x 16 KB).
//   warps 0-3 / 4-7  softmax of query tile 0 / 1          warp 8  TMA producer
//   warp 9           UMMA issuer for both tiles (polls: whichever tile's next MMA has its inputs goes first)
//   warps 10-11      the 257th query row on the CUDA cores from the K / V tiles in shared memory (as in attn_row_kernel)
// Whole warpgroups per role + setmaxnreg: the softmax threads hold 128 scores + 32 packed probabilities in registers.
#pragma once
#include "vt_attn.cuh"

namespace vt {

constexpr int APP_THREADS = 512;                               // 8 softmax warps, TMA, MMA, 6 warps for the 257th query row
constexpr int APP_TAIL_WARPS = 5, APP_TAIL_THREADS = APP_TAIL_WARPS * 32;
constexpr int APP_KBUF = 2 * ATT_TILE_BYTES + 2048;            // two 128-row tiles + one 16-row tail box (row 256)
constexpr int APP_SMEM_K = 0;                                  // [2][APP_KBUF]
constexpr int APP_SMEM_V = 2 * APP_KBUF;                       // [2][APP_KBUF]
constexpr int APP_SMEM_Q = 4 * APP_KBUF;                       // [tile 2][buffer 2][16 KB]
constexpr int APP_SMEM_QT = APP_SMEM_Q + 4 * ATT_TILE_BYTES;   // [buffer 2][2 KB]: the 16-row box that holds query row 256
constexpr int APP_SMEM_BAR = APP_SMEM_QT + 2 * 2048;           // 24 mbarriers + the TMEM slot
constexpr int APP_SMEM_TAIL = APP_SMEM_BAR + 512;              // tail warps: q[64], p[272], red[16], part[6][64] floats
constexpr int APP_SMEM_BYTES = 1024 + APP_SMEM_TAIL + (64 + 272 + 16 + APP_TAIL_WARPS * 64) * 4 + 16;
static_assert(APP_SMEM_BYTES <= 227 * 1024, "attn_pp_kernel shared memory");
static_assert(APP_KBUF % 1024 == 0, "128B-swizzled tiles need 1024-byte alignment");
constexpr int APP_REGS_CTRL = 56, APP_REGS_SOFTMAX = 200;    // 56 * 256 + 200 * 256 = 128 * 512: the pool is what the launch allocated
static_assert(APP_REGS_CTRL * 256 + APP_REGS_SOFTMAX * 256 <= 128 * APP_THREADS, "setmaxnreg can only redistribute the registers the launch allocated (128 per thread at 16 warps)");
// columns inside a tile's 256-column region
constexpr int APP_COL_O = 192;
__host__ __device__ constexpr int app_p_col(int chunk) { return chunk < 2 ? chunk * 32 : 128 + (chunk - 2) * 32; }

// developer instrumentation (debug-knobs builds, tools/attn_trace.py): (tag, clock64) pairs of CTA 0's first softmax thread of each
// tile and of its MMA thread, in three slices of vt_dbg_ts
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
  uint8_t* sK = smem + APP_SMEM_K;
  uint8_t* sV = smem + APP_SMEM_V;
  uint8_t* sQ = smem + APP_SMEM_Q;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + APP_SMEM_BAR);
  uint64_t* k_full = bars;             // [buffer]
  uint64_t* v_full = bars + 2;         // [buffer]
  uint64_t* q_full = bars + 4;         // [tile * 2 + buffer]
  uint64_t* buf_empty = bars + 8;      // [buffer]: 2 MMA commits + 8 softmax warps + 6 tail warps
  uint64_t* s_full = bars + 10;        // [tile]
  uint64_t* p_full = bars + 12;        // [tile * 4 + chunk], 128 arrivals
  uint64_t* o_full = bars + 20;        // [tile]
  uint64_t* region_free = bars + 22;   // [tile], 128 arrivals: O read out, the next unit's S may overwrite the region
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  float* tq = reinterpret_cast<float*>(smem + APP_SMEM_TAIL);   // [64]
  float* tp = tq + 64;                                           // [272]
  float* tred = tp + 272;                                        // [16]
  float* tpart = tred + 16;                                      // [tail warp][64]
  uint8_t* sQt = smem + APP_SMEM_QT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.tokens;   // 257

  if (warp == 9) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&buf_empty[i], 2 + 8 + APP_TAIL_WARPS);
        mbar_init(&s_full[i], 1);
        mbar_init(&o_full[i], 1);
        mbar_init(&region_free[i], 128);
      }
      for (int i = 0; i < 4; ++i) mbar_init(&q_full[i], 1);
      for (int i = 0; i < 8; ++i) mbar_init(&p_full[i], 128);
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

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(APP_REGS_CTRL));
    if (warp == 8) {
      // ------------------------------ TMA producer ------------------------------
      if (lane == 0) {
        int ts_n = 0;
        for (int us = 0; us < n_units; ++us) {
          const int unit = (int)blockIdx.x + us * (int)gridDim.x;
          const int img = unit / a.heads, head = unit - img * a.heads;
          const int row0 = img * N;
          const int b = us & 1;
          const uint32_t par = (us >> 1) & 1;
          uint8_t* kbuf = sK + b * APP_KBUF;
          uint8_t* vbuf = sV + b * APP_KBUF;
          mbar_wait(&buf_empty[b], par ^ 1);
          app_stamp(3, ts_n, 60 + b);
          mbar_arrive_expect_tx(&k_full[b], 2 * ATT_TILE_BYTES + 2048 + 2048);
          tma_load_2d(kbuf, &a.tm, &k_full[b], a.D + head * 64, row0);
          tma_load_2d(kbuf + ATT_TILE_BYTES, &a.tm, &k_full[b], a.D + head * 64, row0 + 128);
          tma_load_2d(kbuf + 2 * ATT_TILE_BYTES, &a.tm16, &k_full[b], a.D + head * 64, row0 + 256);
          tma_load_2d(sQt + b * 2048, &a.tm16, &k_full[b], head * 64, row0 + 256);     // query row 256 (for the tail warps)
          for (int g = 0; g < 2; ++g) {
            mbar_arrive_expect_tx(&q_full[g * 2 + b], ATT_TILE_BYTES);
            tma_load_2d(sQ + (g * 2 + b) * ATT_TILE_BYTES, &a.tm, &q_full[g * 2 + b], head * 64, row0 + g * 128);
          }
          mbar_arrive_expect_tx(&v_full[b], 2 * ATT_TILE_BYTES + 2048);
          tma_load_2d(vbuf, &a.tm, &v_full[b], 2 * a.D + head * 64, row0);
          tma_load_2d(vbuf + ATT_TILE_BYTES, &a.tm, &v_full[b], 2 * a.D + head * 64, row0 + 128);
          tma_load_2d(vbuf + 2 * ATT_TILE_BYTES, &a.tm16, &v_full[b], 2 * a.D + head * 64, row0 + 256);
        }
      }
    } else if (warp == 9 || warp == 10) {
      // ------------------------------ UMMA issuers: warp 9 for tile 0, warp 10 for tile 1 ------------------------------
      // One thread per tile: a single thread serving both tiles (polling) spent ~1 k cycles per action (4 MMAs + commits) and was the
      // slowest role of the kernel (profiles/r02_attn_pp_timeline.txt).
      const int g = warp - 9;
      if (lane == 0 && n_units > 0) {
        const uint32_t idesc_s = umma_idesc(UMMA_FMT_BF16, 256);
        const uint32_t idesc_o = umma_idesc(UMMA_FMT_BF16, 64, 0, 1);
        const uint32_t region = tmem_base + g * 256;
        int ts_n = g * (APP_TS_SLICE / 2);
        // Tile 1 starts half a period behind tile 0 (its first S waits for tile 0's last P chunk): two tiles that walk the chain in
        // lockstep want the MUFU pipe at the same time and leave it idle at the same time.
        if (g == 1) mbar_wait(&p_full[1], 0);
        for (int us = 0; us < n_units; ++us) {
          const int b = us & 1;
          const uint32_t par = (us >> 1) & 1;
          mbar_wait(&k_full[b], par);
          mbar_wait(&q_full[g * 2 + b], par);
          if (us > 0) mbar_wait(&region_free[g], (us - 1) & 1);
          tc_fence_after();
          const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(sQ + (g * 2 + b) * ATT_TILE_BYTES));
          const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sK + b * APP_KBUF));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(region, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
          umma_commit(&s_full[g]);
          app_stamp(2, ts_n, 20 + g);
          mbar_wait(&v_full[b], par);
          const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(sV + b * APP_KBUF));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = (i + 2) & 3;      // chunks are published in the order 2, 3, 0, 1
            mbar_wait(&p_full[g * 4 + c], us & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16_ts(region + APP_COL_O, region + app_p_col(c) + 8 * kk, vdesc + (uint64_t)((c * 64 + kk * 16) * 128 >> 4), idesc_o, (i | kk) != 0);
            app_stamp(2, ts_n, 30 + 10 * g + c);
          }
          umma_commit(&o_full[g]);
          umma_commit(&buf_empty[b]);
        }
      }
    } else {
      // ------------------------------ the 257th query row on the CUDA cores (warps 11..15) ------------------------------
      // 192 threads per unit: the row gates the release of the unit's K / V buffers, and with two warps (attn_row_kernel's split) it
      // took 10.7 k cycles per unit -- longer than everything else in the kernel (profiles/r02_attn_pp_timeline.txt).
      const int tw = warp - 11;                 // 0..4
      const int tid = tw * 32 + lane;           // 0..159
      int ts_n = APP_TS_SLICE / 2;              // second half of the producer's slice
      const bool ts_on = kDbg && tid == 0;
      for (int us = 0; us < n_units; ++us) {
        const int unit = (int)blockIdx.x + us * (int)gridDim.x;
        const int img = unit / a.heads, head = unit - img * a.heads;
        const int b = us & 1;
        const uint32_t par = (us >> 1) & 1;
        // plain (non-volatile) shared-memory loads through generic pointers: the compiler may overlap them, and it still keeps them
        // behind the mbarrier waits (asm volatile with a memory clobber).  The volatile ld.shared helper serialises one shared-
        // memory round trip per 16 bytes, which is what made this row the slowest part of attn_row_kernel.
        const uint8_t* kbase = sK + b * APP_KBUF;
        const uint8_t* vbase = sV + b * APP_KBUF;
        mbar_wait(&k_full[b], par);
        if (ts_on) app_stamp(3, ts_n, 50);
        // q_256 * log2(e) / 8 as fp32 in shared memory (the previous unit's readers left at its last barrier)
        if (tid < 32) {
          const uint32_t raw = *reinterpret_cast<const uint32_t*>(sQt + b * 2048 + tid * 4);   // row 0 of the box: no swizzle offset
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
          tq[2 * tid] = f.x * a.scale_log2;
          tq[2 * tid + 1] = f.y * a.scale_log2;
        }
        named_bar_sync(6, APP_TAIL_THREADS);
        // scores of keys tid and tid + 192
        float sc[2];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int k = tid + APP_TAIL_THREADS * j;
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
          }
          sc[j] = acc;
          mx = fmaxf(mx, acc);
        }
        mx = warp_max(mx);
        if (lane == 0) tred[tw] = mx;
        named_bar_sync(6, APP_TAIL_THREADS);
        mx = tred[0];
#pragma unroll
        for (int w = 1; w < APP_TAIL_WARPS; ++w) mx = fmaxf(mx, tred[w]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int k = tid + APP_TAIL_THREADS * j;
          if (k < N) {
            const float e = ex2_approx(sc[j] - mx);
            tp[k] = e;
            sum += e;
          }
        }
        sum = warp_sum(sum);
        if (lane == 0) tred[8 + tw] = sum;
        mbar_wait(&v_full[b], par);
        named_bar_sync(6, APP_TAIL_THREADS);
        if (ts_on) app_stamp(3, ts_n, 51);
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < APP_TAIL_WARPS; ++w) tot += tred[8 + w];
        const float inv = 1.0f / tot;
        // O[c] = sum_k p[k] V[k][c]: lane -> channels (2 lane, 2 lane + 1), warp tw -> keys k = tw (mod 6)
        float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t vcol = ((lane & 3) << 2), vpc = lane >> 2;
        int k = tw;
        for (; k + 3 * APP_TAIL_WARPS < N; k += 4 * APP_TAIL_WARPS) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int kk = k + APP_TAIL_WARPS * u;
            const uint32_t raw = *reinterpret_cast<const uint32_t*>(vbase + kk * 128 + ((vpc ^ (kk & 7)) << 4) + vcol);
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
            const float pk = tp[kk];
            o0[u] = fmaf(pk, f.x, o0[u]);
            o1[u] = fmaf(pk, f.y, o1[u]);
          }
        }
        for (; k < N; k += APP_TAIL_WARPS) {
          const uint32_t raw = *reinterpret_cast<const uint32_t*>(vbase + k * 128 + ((vpc ^ (k & 7)) << 4) + vcol);
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw));
          const float pk = tp[k];
          o0[0] = fmaf(pk, f.x, o0[0]);
          o1[0] = fmaf(pk, f.y, o1[0]);
        }
        tpart[tw * 64 + 2 * lane] = (o0[0] + o0[1]) + (o0[2] + o0[3]);
        tpart[tw * 64 + 2 * lane + 1] = (o1[0] + o1[1]) + (o1[2] + o1[3]);
        named_bar_sync(6, APP_TAIL_THREADS);
        if (tw == 0) {
          float r0 = 0.f, r1 = 0.f;
#pragma unroll
          for (int w = 0; w < APP_TAIL_WARPS; ++w) {
            r0 += tpart[w * 64 + 2 * lane];
            r1 += tpart[w * 64 + 2 * lane + 1];
          }
          const long long qrow = (long long)img * N + 256;
          *reinterpret_cast<uint32_t*>(a.ctx + qrow * a.ctx_ld + head * 64 + 2 * lane) = pack_bf16x2(r0 * inv, r1 * inv);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&buf_empty[b]);
        if (ts_on) app_stamp(3, ts_n, 52);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(APP_REGS_SOFTMAX));
    // ------------------------------ softmax + output: tile g = warp / 4, one thread per query row ------------------------------
    const int g = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t region = tmem_base + g * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sl = a.scale_log2;
    const float2 sl2 = make_float2(sl, sl);
    // score of the 257th key of unit `us`: q_r . k_256 from shared memory.  Computed for the NEXT unit while this unit's last
    // P V chunk drains (its Q and K tiles are double-buffered and landed long ago), so it is off the tile's critical chain.
    auto tail_score = [&](int us) -> float {
      const int b = us & 1;
      const uint32_t par = (us >> 1) & 1;
      mbar_wait(&q_full[g * 2 + b], par);
      mbar_wait(&k_full[b], par);
      const uint8_t* qrow = sQ + (g * 2 + b) * ATT_TILE_BYTES + r * 128;
      const uint8_t* krow = sK + b * APP_KBUF + 2 * ATT_TILE_BYTES;   // row 256 = row 0 of the tail box
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
      return ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    };
    float st = n_units > 0 ? tail_score(0) : 0.f;
    int ts_n = 0;
    const bool ts_on = kDbg && quarter == 0 && lane == 0;
    for (int us = 0; us < n_units; ++us) {
      const int unit = (int)blockIdx.x + us * (int)gridDim.x;
      const int img = unit / a.heads, head = unit - img * a.heads;
      const int b = us & 1;
      const uint32_t par = (us >> 1) & 1;
      if (ts_on) app_stamp(g, ts_n, 0);
      mbar_wait(&s_full[g], us & 1);
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
        mbar_arrive(&p_full[g * 4 + chunk]);
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
      const float st_next = us + 1 < n_units ? tail_score(us + 1) : 0.f;
      if (ts_on) app_stamp(g, ts_n, 7);
      // ---- O = P V from tensor memory + the rank-1 term of the 257th key, scaled, 128 contiguous bytes per row ----
      mbar_wait(&v_full[b], par);
      mbar_wait(&o_full[g], us & 1);
      tc_fence_after();
      if (ts_on) app_stamp(g, ts_n, 8);
      tmem_ld64(region + APP_COL_O, x0);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&region_free[g]);
      {
        const uint8_t* vrow = sV + b * APP_KBUF + 2 * ATT_TILE_BYTES;   // V row 256
        __nv_bfloat16* dst = a.ctx + ((long long)img * N + g * 128 + r) * a.ctx_ld + head * 64;
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
      __syncwarp();
      if (lane == 0) mbar_arrive(&buf_empty[b]);
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
