// Persistent multi-layer kernel: a whole U-Net evaluation -- or the whole sde_vs sampling loop, every step of it -- in ONE launch.
//
// Reference: StochasticInterpolants.sde_vs, bridge/bridge_model.py:343-385 (the hot loop: per step two U-Net evaluations,
// DiffusionConditionalUnet1D.forward conditional_unet_1D.py:194-247, then the Euler-Maruyama update) and get_loss' three
// evaluations (:240-242).  The multi-launch form of the same work is 36 gemm_tc_kernel launches + one sde_step_kernel per
// step; every launch pays ~6 us of fixed cost (launch, barrier / TMEM set-up, first operand round trip) and a grid-wide
// drain, which at batch 256 is a quarter of the U-Net's time and at batch 1 nearly all of it.
//
// Here one CTA pair per TPC stays resident and walks a STATIC schedule: the tiles of all layers of all steps are numbered
// consecutively and dealt round-robin to the pairs.  What orders them is data, not launches:
//   * nothing in the U-Net crosses samples (convolutions run along T inside a sample, GroupNorm and FiLM are per sample), so a
//     tile of layer l only needs the tiles of its producer layers that cover the SAME samples;
//   * per (layer, net, sample block) a counter in global memory counts finished epilogue warps; the TMA producer of a consumer
//     tile spins on `counter >= expected * (step + 1)` (ld.acquire.gpu) before it requests the A operand, the epilogue warps
//     do the same before they read a residual, and they publish their own tile with red.release.gpu after their stores;
//   * generic-proxy stores that a later TMA load (async proxy) reads are fenced with fence.proxy.async on both sides.
// A tile only ever waits for tiles with a smaller sequence number and every pair processes its tiles in sequence order, so
// the schedule cannot deadlock as long as all pairs are resident (grid <= SM count, one CTA per SM).
// The Euler-Maruyama update is one more "layer" (kind 1) executed by the epilogue warps; it gates the next step's first layer.
//
// Inside a pair everything is gemm_tc_kernel's machinery (same epilogue code, vt_gemm.cuh): warp 0 TMA producer running ahead
// through a 3 x 2-atom operand ring, warp 1 single-thread tcgen05.mma.cta_group::2 issuer (M = 256) into two TMEM
// accumulators, warps 2-9 epilogue (GroupNorm + Mish + FiLM + residual, or the linear epilogues), so the epilogue of a tile
// overlaps the main loop of the pair's next tile -- which may belong to the next layer.
#pragma once
#include "vt_gemm.cuh"

namespace vt {

constexpr int PERSIST_MAX_LAYERS = 38;
constexpr int PERSIST_MAX_DEPS = 4;
constexpr int PERSIST_STAGES = 3;
constexpr int PERSIST_KA = 2;
constexpr int PERSIST_BN_MAX = 256;
// Warpgroup 0 = {TMA producer, MMA issuer, two idle warps}, warpgroups 1-2 = the eight epilogue warps.  With 10 warps two of the
// four SM sub-partitions host three warps and ptxas caps every thread at 168 registers, which the GroupNorm epilogue exceeds
// (its spill reloads were 39 % of the stall samples of its inner loop, profiles/r02_persist_epilogue_stalls.txt); with whole
// warpgroups per role `setmaxnreg` moves the registers warpgroup 0 does not need to the epilogue warps.
constexpr int PERSIST_THREADS = 128 + 256;
constexpr int PERSIST_EPI_T0 = 128;          // first epilogue thread
constexpr int PERSIST_REGS_CTRL = 56, PERSIST_REGS_EPI = 224;
static_assert(PERSIST_REGS_CTRL * 128 + PERSIST_REGS_EPI * 256 <= 168 * PERSIST_THREADS, "setmaxnreg can only redistribute the registers the launch allocated (168 per thread at 12 warps)");
// operand ring (A 16 KB + B up to 16 KB per atom) + barriers + the GroupNorm epilogue's scratch (the transposition buffers of
// the linear epilogue alias it: 8 warps x 4 KB)
constexpr int PERSIST_SCRATCH_FLOATS = GEMM_SCRATCH_FLOATS(PERSIST_BN_MAX, EPI_GN);
static_assert(PERSIST_SCRATCH_FLOATS * 4 >= 8 * 4096, "the linear epilogue's transposition buffers alias the GroupNorm scratch");
constexpr int PERSIST_SMEM_BYTES =
    1024 + PERSIST_STAGES * PERSIST_KA * (GEMM_A_STAGE_BYTES + (PERSIST_BN_MAX / 2) * 128) + 256 + PERSIST_SCRATCH_FLOATS * 4;
static_assert(PERSIST_SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct PersistLayer {
  GemmArgs g;                      // kind 0: the implicit GEMM (CTA-pair tensor maps)
  int kind;                        // 0 implicit GEMM, 1 Euler-Maruyama update of one sample block
  int mode;                        // EPI_LINEAR / EPI_GN
  int bn;                          // 256 or 32
  int out_f32;                     // linear epilogue writes fp32 (the nets' outputs)
  uint32_t idesc;                  // UMMA instruction descriptor (bf16, N = bn, M = 256)
  int tile_base;                   // sequence number of this layer's tile 0 inside one step
  int n_tiles_total;               // tiles of this layer per step
  int G;                           // groups (nets)
  int spu;                         // samples covered by one tile
  int cnt_off;                     // counters: done[cnt_off + g * n_sb + sb]
  int exp_off;                     // expected[exp_off + sb]: arrivals per step that complete (g, sb)
  int n_dep;
  int dep[PERSIST_MAX_DEPS];       // producing layers ...
  int dep_lag[PERSIST_MAX_DEPS];   // ... 1: their output of the PREVIOUS step (x written by the Euler-Maruyama update)
  long long film_t_step;           // elements between consecutive steps' rows of g.film_t
};

struct PersistSde {                // vt_sde_desc without the per-step scalars
  float* x;
  const float* v;
  const float* s;
  const float* noise;              // [n_steps][rows][A] or null
  long long noise_step;
  int A, T;                        // action dim, rows per sample
  float d;
  unsigned long long seed;
  const unsigned long long* seed_dev;
  void* xpad;
  int xpad_dtype, xpad_ld;
  long long xpad_plane;
};

struct PersistCoef {               // per-step scalars of sde_vs (bridge_model.py:347-385), host table in device memory
  float ginv, dgg, eps, dt, nscale, pad0, pad1, pad2;
};

struct PersistArgs {
  PersistLayer layers[PERSIST_MAX_LAYERS];
  PersistSde sde;
  int n_layers, n_steps, tiles_per_step, n_sb, sbs, B;
  unsigned* done;                  // zeroed before the launch
  const unsigned* expected;
  const PersistCoef* coef;
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy accesses to global memory <-> async-proxy (TMA) accesses to the same addresses
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Block until every producer of (layer L, group g, tile `unit`) has published the sample blocks the tile covers.
__device__ __forceinline__ void persist_wait(const PersistArgs& P, const PersistLayer& L, int step, int g, int unit) {
  const int s0 = unit * L.spu;
  const int s1 = min(s0 + L.spu, P.B) - 1;
  const int sb0 = s0 / P.sbs, sb1 = s1 / P.sbs;
  for (int d = 0; d < L.n_dep; ++d) {
    const PersistLayer& D = P.layers[L.dep[d]];
    const unsigned mult = (unsigned)(step + 1 - L.dep_lag[d]);
    if (mult == 0) continue;     // produced before the launch (the prior x of step 0)
    int g0 = g, g1 = g + 1;
    if (D.G != L.G) {
      g0 = 0;
      g1 = D.G == 1 ? 1 : D.G;   // a shared producer (one group), or a consumer of all groups (the Euler-Maruyama update)
    }
    for (int gp = g0; gp < g1; ++gp) {
      for (int sb = sb0; sb <= sb1; ++sb) {
        const unsigned need = P.expected[D.exp_off + sb] * mult;
        const unsigned* c = P.done + D.cnt_off + gp * P.n_sb + sb;
        if (ld_acquire_gpu(c) >= need) continue;
        const long long t0 = clock64();
        while (ld_acquire_gpu(c) < need) {
          __nanosleep(40);
          if ((clock64() - t0) > 8000000000LL) __trap();   // a protocol bug must trap, never hang the GPU box
        }
      }
    }
  }
}

__device__ __forceinline__ void persist_signal(const PersistArgs& P, const PersistLayer& L, int g, int unit) {
  const int s0 = unit * L.spu;
  const int s1 = min(s0 + L.spu, P.B) - 1;
  for (int sb = s0 / P.sbs; sb <= s1 / P.sbs; ++sb) red_release_gpu_add(P.done + L.cnt_off + g * P.n_sb + sb, 1u);
}

// first tile of a layer this worker owns: tiles are numbered consecutively over layers and steps and dealt round-robin
__device__ __forceinline__ int persist_first_tile(int seq_base, int worker, int n_workers) {
  const int r = seq_base % n_workers;
  return worker >= r ? worker - r : worker - r + n_workers;
}

__global__ void __launch_bounds__(PERSIST_THREADS, 1) unet_persist_kernel(const __grid_constant__ PersistArgs P) {
  using bf = __nv_bfloat16;
  constexpr int STAGES = PERSIST_STAGES, KA = PERSIST_KA, KE = 64;
  constexpr int A_ATOM_BYTES = GEMM_A_STAGE_BYTES;
  constexpr int B_ATOM_SLOT = (PERSIST_BN_MAX / 2) * 128;
  constexpr int A_STAGE_BYTES = KA * A_ATOM_BYTES, B_STAGE_BYTES = KA * B_ATOM_SLOT;
  constexpr uint32_t ACC_COLS = PERSIST_BN_MAX, TMEM_COLS = 512;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int worker = blockIdx.x / 2, n_workers = gridDim.x / 2;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&acc_full[s], 1);
        mbar_init(&acc_empty[s], 8 * 2);   // one arrival per epilogue warp of both CTAs
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything before this line overlaps the predecessor's tail (programmatic dependent launch)
  long long* tr_base = nullptr;   // developer instrumentation (debug-knobs builds, VT_GEMM_DEBUG bit 512): see vt_gemm.cuh ptrace
#if VT_DEBUG_KNOBS
  if ((P.layers[0].g.debug & 512) && rank == 0 && worker < PTRACE_WORKERS) {
    tr_base = vt_ptrace + (size_t)worker * PTRACE_TILES * PTRACE_SLOTS;
    if (threadIdx.x == 0) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      vt_ptrace_cal[worker * 4 + 0] = clock64();
      vt_ptrace_cal[worker * 4 + 1] = (long long)gt;
    }
  }
#endif
  auto tr_row = [&](uint32_t lt) -> long long* { return (tr_base && lt < (uint32_t)PTRACE_TILES) ? tr_base + lt * PTRACE_SLOTS : nullptr; };

  if (warp < 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PERSIST_REGS_CTRL));
   if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0, lt = 0;
      for (int step = 0; step < P.n_steps; ++step) {
        for (int l = 0; l < P.n_layers; ++l) {
          const PersistLayer& L = P.layers[l];
          if (L.kind != 0) continue;
          const GemmArgs& a = L.g;
          const int nk = a.taps * a.cblocks;
          const int b_rows = L.bn >> 1;
          const uint32_t stage_tx = 2u * (uint32_t)(a.a_box_bytes + b_rows * 128);   // both CTAs, per atom
          for (int tile = persist_first_tile(step * P.tiles_per_step + L.tile_base, worker, n_workers); tile < L.n_tiles_total;
               tile += n_workers, ++lt) {
            const int n_tile = tile % a.n_tiles, rest = tile / a.n_tiles;
            const int unit = rest % a.m_tiles, g = rest / a.m_tiles;
            const int m_tile = unit * 2 + rank;
            const int b_base = m_tile * a.m_b_step, t_base = m_tile * a.m_t_step, g_a = g * a.a_g_mul;
            const int g_b = g * a.n_pad + n_tile * L.bn + rank * b_rows;
            const int n_stages = (nk + KA - 1) / KA;
            // The weight tiles depend on nothing: the first ring-full of stages is opened and its B halves are requested BEFORE
            // the wait for the producers of the A operand, so their L2 round trip overlaps the dependency latency.
            const int pre = n_stages < STAGES ? n_stages : STAGES;
            const int s0 = s;
            for (int q = 0; q < pre; ++q) {
              const int n_at = (nk - q * KA) < KA ? (nk - q * KA) : KA;
              mbar_wait(&empty[s], ph ^ 1);
              if (rank == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)n_at * stage_tx);
              const uint32_t fb = mapa_shared(smem_u32(&full[s]), 0);
              for (int at = 0; at < n_at; ++at)
                tma_load_2d_pair(sB + s * B_STAGE_BYTES + at * B_ATOM_SLOT, &a.tmB, fb, (q * KA + at) * KE, g_b);
              if (++s == STAGES) {
                s = 0;
                ph ^= 1;
              }
            }
            ptrace(tr_row(lt), 1);
            persist_wait(P, L, step, g, unit);
            ptrace(tr_row(lt), 2);
            fence_proxy_async_global();
            for (int q = 0; q < n_stages; ++q) {
              const int n_at = (nk - q * KA) < KA ? (nk - q * KA) : KA;
              int slot;
              if (q < pre) {
                slot = s0 + q >= STAGES ? s0 + q - STAGES : s0 + q;
              } else {
                slot = s;
                mbar_wait(&empty[s], ph ^ 1);
                if (rank == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)n_at * stage_tx);
                if (++s == STAGES) {
                  s = 0;
                  ph ^= 1;
                }
              }
              const uint32_t fb = mapa_shared(smem_u32(&full[slot]), 0);
              for (int at = 0; at < n_at; ++at) {
                const int i = q * KA + at;
                const int tp = i / a.cblocks, cb = i - tp * a.cblocks;
                tma_load_5d_pair(sA + slot * A_STAGE_BYTES + at * A_ATOM_BYTES, &a.tmA, fb, a.a_c0 + cb * KE, a.tap_p[tp],
                                 t_base + a.tap_t[tp], b_base, g_a);
                if (q >= pre) tma_load_2d_pair(sB + slot * B_STAGE_BYTES + at * B_ATOM_SLOT, &a.tmB, fb, i * KE, g_b);
              }
            }
          }
        }
      }
    }
   } else if (warp == 1) {
    // ------------------------------ UMMA issuer (leader CTA only) ------------------------------
    if (lane == 0 && rank == 0) {
      uint32_t lt = 0, ph = 0;
      int s = 0;
      for (int step = 0; step < P.n_steps; ++step) {
        for (int l = 0; l < P.n_layers; ++l) {
          const PersistLayer& L = P.layers[l];
          if (L.kind != 0) continue;
          const int nk = L.g.taps * L.g.cblocks;
          const uint32_t idesc = L.idesc;
          for (int tile = persist_first_tile(step * P.tiles_per_step + L.tile_base, worker, n_workers); tile < L.n_tiles_total;
               tile += n_workers, ++lt) {
            const uint32_t acc = lt & 1;
            mbar_wait(&acc_empty[acc], ((lt >> 1) & 1) ^ 1);
            tc_fence_after();
            ptrace(tr_row(lt), 3);
            const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
            for (int i = 0; i < nk; i += KA) {
              const int n_at = (nk - i) < KA ? (nk - i) : KA;
              mbar_wait(&full[s], ph);
              tc_fence_after();
              if (i == 0) ptrace(tr_row(lt), 4);
#pragma unroll
              for (int at = 0; at < KA; ++at) {
                if (at >= n_at) break;
                const uint64_t adesc = umma_smem_desc_sw128(smem_u32(sA + s * A_STAGE_BYTES + at * A_ATOM_BYTES));
                const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(sB + s * B_STAGE_BYTES + at * B_ATOM_SLOT));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (i | at | k) != 0);
              }
              umma_commit_pair(&empty[s], 3);
              if (++s == STAGES) {
                s = 0;
                ph ^= 1;
              }
            }
            umma_commit_pair(&acc_full[acc], 3);
            ptrace(tr_row(lt), 5);
          }
        }
      }
    }
   }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PERSIST_REGS_EPI));
    // ------------------------------ epilogue warps ------------------------------
    constexpr int BN = PERSIST_BN_MAX;
    const int half = (warp - 4) >> 2;
    const int et = (threadIdx.x - PERSIST_EPI_T0) & 127;
    const int et256 = threadIdx.x - PERSIST_EPI_T0;
    const int quarter = warp & 3;
    constexpr int CVG = GEMM_COLV_FLOATS(BN, EPI_GN);
    float2* gn_part = reinterpret_cast<float2*>(scratch + CVG) + half * (128 + 64) * 4;
    float2* gn_stat = gn_part + 128 * 4;
    float* films = scratch + CVG + 2 * (128 + 64) * 4 * 2;
    const uint32_t xbuf = smem_u32(scratch) + (warp - 4) * 4096;
    EpiTile t;
    t.r = quarter * 32 + lane;
    t.dbg_n = 0;
    t.tr = nullptr;
    uint32_t lt = 0;
    for (int step = 0; step < P.n_steps; ++step) {
      for (int l = 0; l < P.n_layers; ++l) {
        const PersistLayer& L = P.layers[l];
        const int first = persist_first_tile(step * P.tiles_per_step + L.tile_base, worker, n_workers);
        if (L.kind == 1) {
          // ---- Euler-Maruyama update of sample block `tile` (bridge_model.py:363-385, reference operation order) ----
          const PersistSde& S = P.sde;
          for (int tile = first; tile < L.n_tiles_total; tile += n_workers) {
            if (lane == 0) persist_wait(P, L, step, 0, tile);
            __syncwarp();
            const PersistCoef c = P.coef[step];
            const long long e0 = (long long)tile * L.spu * S.T * S.A;
            const long long e1 = (long long)min((tile + 1) * L.spu, P.B) * S.T * S.A;
            unsigned long long seed = S.seed;
            if (S.seed_dev) seed += *S.seed_dev;
            const long long vs_stride = (long long)P.B * S.T * S.A;   // v = group 0, s = group 1 of the nets' output
            for (long long idx = e0 + rank * 256 + et256; idx < e1; idx += 512) {
              float z;
              if (S.noise) {
                z = S.noise[(long long)step * S.noise_step + idx];
              } else {
                curandStatePhilox4_32_10_t st;
                curand_init(seed, (unsigned long long)idx, 4ull * (unsigned long long)step, &st);
                z = curand_normal(&st);
              }
              const float sv = __fmul_rn(S.v[vs_stride + idx], c.ginv);
              const float b = __fsub_rn(S.v[idx], __fmul_rn(__fmul_rn(c.dgg, sv), c.eps));
              const float dW = __fmul_rn(S.d, z);
              float nx = __fadd_rn(S.x[idx], __fmul_rn(__fadd_rn(b, __fmul_rn(c.eps, sv)), c.dt));
              nx = __fadd_rn(nx, __fmul_rn(c.nscale, dW));
              S.x[idx] = nx;
              if (S.xpad) {
                const int a_ = (int)(idx % S.A);
                const long long r = idx / S.A;
                store_val(S.xpad, S.xpad_dtype, r * S.xpad_ld + a_, S.xpad_plane, nx);
              }
            }
            fence_proxy_async_global();   // xpad is the next step's first A operand (TMA)
            __syncwarp();
            if (lane == 0) persist_signal(P, L, 0, tile);
          }
          continue;
        }
        const GemmArgs& a = L.g;
        for (int tile = first; tile < L.n_tiles_total; tile += n_workers, ++lt) {
          const uint32_t acc = lt & 1;
          const int n_tile = tile % a.n_tiles, rest = tile / a.n_tiles;
          const int unit = rest % a.m_tiles;
          const int m_tile = unit * 2 + rank;
          t.g = rest / a.m_tiles;
          t.n0 = n_tile * L.bn;
#if VT_DEBUG_KNOBS
          t.tr = threadIdx.x == PERSIST_EPI_T0 ? tr_row(lt) : nullptr;
          ptrace_val(t.tr, 0, ((long long)step << 40) | ((long long)l << 20) | tile);
          ptrace(t.tr, 6);
#endif
          // the residual rows come from another pair's epilogue: acquire before the first read (the A operand's producers were
          // acquired by the TMA thread; this warp's own acquire also covers them)
          if (lane == 0) persist_wait(P, L, step, t.g, unit);
          __syncwarp();
          const bool fast_lin = L.mode == EPI_LINEAR && L.bn == BN && a.fast != 0 && !L.out_f32;
          named_bar_sync(1, 256);   // every warp has left the previous tile: the shared scratch may be rewritten
          float* colv = scratch + (L.mode == EPI_LINEAR ? (int)acc * 2 * L.bn : 0);
          bool film_staged = false;
          if (!fast_lin) {
            // Per-column vectors and the per-sample FiLM table of this tile -> shared memory.  All global loads of a thread are
            // issued before its first shared-memory store: written as a plain load / store loop the compiler must assume that the
            // stores alias the next loads and serialises one L2 round trip per element (ncu: the top stall of the epilogue warps).
            const long long gcol = (long long)t.g * a.n_pad + t.n0;
            if (et256 < L.bn) {
              const int c = et256;
              if (L.mode == EPI_LINEAR) {
                const float b_ = a.bias ? __ldg(a.bias + gcol + c) : 0.f;
                const float s_ = (a.colscale && (t.n0 + c) < a.N) ? __ldg(a.colscale + t.n0 + c) : 1.f;
                colv[c] = b_;
                colv[L.bn + c] = s_;
              } else {
                const bool f = a.film_t != nullptr;
                const long long fo = (long long)t.g * a.film_tg + (long long)step * L.film_t_step + a.film_off + t.n0 + c;
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
            if (L.mode == EPI_GN) {
              const int nsamp = a.rows_valid / a.gn_rows;
              if (a.film_c && nsamp <= GEMM_FILM_SAMPLES && a.out_plane == 0 && a.res_plane == 0) {
                film_staged = true;
                stage_film<BN>(a, films, a.film_t ? a.film_t + (long long)t.g * a.film_tg + (long long)step * L.film_t_step + a.film_off + t.n0 : nullptr,
                               t.g, t.n0, m_tile, nsamp, et256, 256);
              }
            }
            named_bar_sync(1, 256);
          }
          t.grow = (long long)m_tile * a.rows_valid + t.r;
          t.valid = (t.r < a.rows_valid) && (t.grow < a.M_total);
          t.q = a.row_div == 1 ? (int)t.grow : (int)((unsigned)t.grow / (unsigned)a.row_div);
          t.rem = (int)(t.grow - (long long)t.q * a.row_div);
          t.taddr = tmem_base + acc * ACC_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
          const uint32_t parity = (lt >> 1) & 1;
          if (L.mode == EPI_GN) {
            const int c_begin = half * (BN / 2);
            if (a.out_plane == 0 && a.res_plane == 0)
              epilogue_gn_fast<BN>(a, t, colv, film_staged ? films : nullptr, gn_part, gn_stat, et, 2 + half, &acc_full[acc], parity, c_begin);
            else
              epilogue_gn<BN, bf, false>(a, t, colv, gn_part, gn_stat, et, 2 + half, &acc_full[acc], parity, c_begin);
          } else if (L.bn == BN) {
            const int c_begin = half * (BN / 2), c_end = c_begin + BN / 2;
            if (fast_lin) epilogue_linear_fast<BN, bf, false>(a, t, xbuf, &acc_full[acc], parity, c_begin, c_end);
            else if (L.out_f32) epilogue_linear<BN, float, false>(a, t, colv, &acc_full[acc], parity, c_begin, c_end);
            else epilogue_linear<BN, bf, false>(a, t, colv, &acc_full[acc], parity, c_begin, c_end);
          } else if (half == 0) {   // bn == 32: one 32-column chunk, the first warpgroup takes it
            if (L.out_f32) epilogue_linear<32, float, false>(a, t, colv, &acc_full[acc], parity, 0, 32);
            else epilogue_linear<32, bf, false>(a, t, colv, &acc_full[acc], parity, 0, 32);
          }
          ptrace(t.tr, 9);
          tc_fence_before();
          fence_proxy_async_global();   // this tile's rows are A operands (TMA) of the consumer layers
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_remote(mapa_shared(smem_u32(&acc_empty[acc]), 0));
            persist_signal(P, L, t.g, unit);
          }
          ptrace(t.tr, 10);
        }
      }
    }
  }

#if VT_DEBUG_KNOBS
  if (tr_base && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    vt_ptrace_cal[worker * 4 + 2] = clock64();
    vt_ptrace_cal[worker * 4 + 3] = (long long)gt;
  }
#endif
  tc_fence_before();
  cluster_sync_all();   // the peer may still signal / read this CTA
  if (warp == 1) tmem_dealloc_pair(tmem_base, TMEM_COLS);
}

}  // namespace vt
