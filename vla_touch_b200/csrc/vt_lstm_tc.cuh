// nn.LSTM recurrence on the tensor cores (hidden size 256, gate order i,f,g,o; lstm_step_controller.py:66-73,196-204 and the
// back-propagation through time torch autograd runs for lstm_train.py:130).
//
// The CUDA-core kernels of vt_lstm.cuh give every CTA 4 batch rows and re-read the whole fp32 W_hh (1 MB) from L2 at every time
// step: at the LSTM benchmark shape (batch 512, T = 128) that is ~32 GB of L2 traffic per layer and 5 ms (forward) / 13 ms (BPTT).
// Here a CLUSTER of 8 CTAs owns 128 batch rows; CTA r owns hidden units [32 r, 32 r + 32):
//
//   forward   gates_t[:, units r] = xw_t[:, units r] + h_{t-1} @ W_hh[units r]^T     tcgen05.mma  M = 128 rows, N = 128 (4 gates x
//             32 units), K = 256.  The CTA's 64 KB slice of W_hh (bf16) is loaded into shared memory ONCE and stays there for all
//             T steps; h_{t-1} (128 x 256 bf16, 64 KB) and the CTA's slice of the input projections xw_t come in by TMA; the four
//             epilogue warps (one TMEM lane = one batch row per thread) apply the gate non-linearities, keep c in registers and
//             write their 32 units of h_t (bf16) to global memory, where the seven peer CTAs pick it up for step t + 1.
//   backward  d gates_t[:, units r] is elementwise (from the stored gates / cell states, dy_t, and dh from step t + 1); then
//             dh_{t-1}[:, units r] = d gates_t @ W_hh[:, units r]   M = 128, N = 32, K = 1024: the CTA's 64 KB slice of W_hh^T is
//             resident, the full d gates_t row block (256 KB bf16, every CTA's slice) streams through a 2 x 64 KB TMA ring.
//
// One barrier.cluster per time step orders the exchange (release / acquire at cluster scope; generic-proxy stores that a peer's
// TMA load reads are fenced with fence.proxy.async on both sides).  The sequential chain per step is TMA -> MMA -> epilogue ->
// cluster barrier, about 3 us; the layer input projections, weight gradients and d x remain the big GEMMs they were.
// bf16 operands, fp32 accumulation and state; sigmoid / tanh through MUFU tanh.approx (the fp32 parity mode keeps vt_lstm.cuh).
#pragma once
#include "vt_gemm.cuh"

namespace vt {

constexpr int LTC_H = 256;
constexpr int LTC_CLUSTER = 8;
constexpr int LTC_U = LTC_H / LTC_CLUSTER;   // hidden units per CTA: 32
constexpr int LTC_ROWS = 128;                // batch rows per cluster
constexpr int LTC_THREADS = 192;             // 4 epilogue warps, 1 TMA warp, 1 MMA warp
constexpr int LTC_ATOM_A = LTC_ROWS * 128;   // one K atom of a 128-row operand: 16 KB
// forward: W slice (128 rows x 256 K = 4 atoms) + h tile (4 atoms) + xw slice (4 gates x 128 rows x 32 fp32)
constexpr int LTC_FWD_SMEM = 1024 + 4 * LTC_ATOM_A + 4 * LTC_ATOM_A + 4 * LTC_ATOM_A + 256;
// backward: W^T slice (32 rows x 1024 K = 16 atoms of 4 KB) + 2 stages x 4 atoms of d gates
constexpr int LTC_BWD_STAGE_ATOMS = 4;
constexpr int LTC_BWD_SMEM = 1024 + 16 * LTC_U * 128 + 2 * LTC_BWD_STAGE_ATOMS * LTC_ATOM_A + 256;
static_assert(LTC_FWD_SMEM <= 227 * 1024 && LTC_BWD_SMEM <= 227 * 1024, "shared memory budget");

struct LstmTcArgs {
  CUtensorMap tmW;    // bf16 [4H][H], row = unit * 4 + gate (K contiguous); box (64, 128)
  CUtensorMap tmH;    // bf16 h of all steps, 3-D (H, T, B); box (64, 1, 128)
  CUtensorMap tmX;    // fp32 xw, 3-D (4H, T, B); box (32, 1, 128), SWIZZLE_128B
  __nv_bfloat16* hbuf;   // [B][T][H] bf16: h of every step (the recurrence's own copy)
  void* y;               // [B][T][y_ld] hidden outputs in y_dtype (may be null when hbuf is all the caller needs)
  int y_dtype;
  long long y_ld;
  float* gates;          // training: [B][T][4H] activated gates (i, f, g, o), or null
  float* c_all;          // training: [B][T][H], or null
  float* h_out;          // final state [B][H] or null
  float* c_out;
  int B, T;
};

struct LstmBwdTcArgs {
  CUtensorMap tmW;    // bf16 W_hh^T [H][4H] (K = gate * H + unit contiguous); box (64, 32)
  CUtensorMap tmD;    // bf16 d gates of all steps, 3-D (4H, T, B); box (64, 1, 128)
  const float* gates;    // [B][T][4H]
  const float* c_all;    // [B][T][H]
  const float* dy;       // [B][T][dy_ld]
  long long dy_ld;
  float* dgates;         // [B][T][4H] fp32 (consumed by the weight-gradient GEMMs)
  __nv_bfloat16* dgb;    // [B][T][4H] bf16 copy: the A operand of the recurrence
  int B, T;
};

__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

// ------------------------------------------------------------------------------------------------------------
// forward (inference and training): zero initial state
// ------------------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(LTC_CLUSTER, 1, 1) __launch_bounds__(LTC_THREADS, 1) lstm_tc_kernel(const __grid_constant__ LstmTcArgs a) {
  constexpr int H = LTC_H, U = LTC_U;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                       // 4 atoms [128 n][128 B]
  uint8_t* sA = sW + 4 * LTC_ATOM_A;        // 4 atoms [128 rows][128 B]
  uint8_t* sX = sA + 4 * LTC_ATOM_A;        // 4 gates [128 rows][128 B] fp32, 128-byte swizzle
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 4 * LTC_ATOM_A);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* x_full = bars + 2;
  uint64_t* x_empty = bars + 3;
  uint64_t* acc_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int b0 = (blockIdx.x / LTC_CLUSTER) * LTC_ROWS;
  const int u0 = rank * U;
  constexpr uint32_t IDESC = umma_idesc(UMMA_FMT_BF16, 4 * U, 0, 0, 128);

  if (warp == 5) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      mbar_init(a_full, 1);
      mbar_init(x_full, 1);
      mbar_init(x_empty, 4);
      mbar_init(acc_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&a.tmW);
    tma_prefetch_desc(&a.tmH);
    tma_prefetch_desc(&a.tmX);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_arrive_release();   // every CTA of the cluster is set up before the first exchange
  cluster_wait_acquire();

  if (warp == 4) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 4 * LTC_ATOM_A);
      for (int k = 0; k < 4; ++k) tma_load_2d(sW + k * LTC_ATOM_A, &a.tmW, w_full, k * 64, u0 * 4);
    }
    for (int t = 0; t < a.T; ++t) {
      if (lane == 0) {
        if (t > 0) {   // h_{t-1} of all 256 units: every CTA of the cluster has written its 32 (ordered by the cluster barrier)
          fence_proxy_async_all();
          mbar_arrive_expect_tx(a_full, 4 * LTC_ATOM_A);
          for (int k = 0; k < 4; ++k) tma_load_3d(sA + k * LTC_ATOM_A, &a.tmH, a_full, k * 64, t - 1, b0);
          mbar_wait(x_empty, (t - 1) & 1);
        }
        mbar_arrive_expect_tx(x_full, 4 * LTC_ATOM_A);
        for (int g = 0; g < 4; ++g) tma_load_3d(sX + g * LTC_ATOM_A, &a.tmX, x_full, g * H + u0, t, b0);
      }
      __syncwarp();
      cluster_arrive_release();
      cluster_wait_acquire();
    }
  } else if (warp == 5) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) mbar_wait(w_full, 0);
    __syncwarp();
    for (int t = 0; t < a.T; ++t) {
      if (lane == 0 && t > 0) {   // t = 0: h_{-1} = 0, the gates are the input projections alone
        mbar_wait(a_full, (t - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adesc = umma_smem_desc_sw128(smem_u32(sA + k * LTC_ATOM_A));
          const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(sW + k * LTC_ATOM_A));
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_f16(tmem_base, adesc + 2 * j, bdesc + 2 * j, IDESC, (k | j) != 0);
        }
        umma_commit(acc_full);
      }
      __syncwarp();
      cluster_arrive_release();
      cluster_wait_acquire();
    }
  } else {
    // ------------------------------ gate math: one batch row per thread ------------------------------
    const int row = warp * 32 + lane;
    const int b = b0 + row;
    const bool ok = b < a.B;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t xrow = smem_u32(sX) + row * 128;
    const int sw = row & 7;
    float c[U];
#pragma unroll
    for (int i = 0; i < U; ++i) c[i] = 0.f;
    for (int t = 0; t < a.T; ++t) {
      mbar_wait(x_full, t & 1);
      if (t > 0) {
        mbar_wait(acc_full, (t - 1) & 1);
        tc_fence_after();
      }
      const long long bt = (long long)b * a.T + t;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {   // 8 units x 4 gates per 32-column chunk (column = unit * 4 + gate)
        uint32_t v[32];
        if (t > 0) {
          tmem_ld32(taddr + ch * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        float xg[4][8];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 lo = ld_shared_v4f(xrow + g * LTC_ATOM_A + (((2 * ch) ^ sw) << 4));
          const float4 hi = ld_shared_v4f(xrow + g * LTC_ATOM_A + (((2 * ch + 1) ^ sw) << 4));
          xg[g][0] = lo.x; xg[g][1] = lo.y; xg[g][2] = lo.z; xg[g][3] = lo.w;
          xg[g][4] = hi.x; xg[g][5] = hi.y; xg[g][6] = hi.z; xg[g][7] = hi.w;
        }
        float hv[8], gi[8], gf[8], gg[8], go[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          gi[j] = sigmoid_fast(__uint_as_float(v[4 * j]) + xg[0][j]);
          gf[j] = sigmoid_fast(__uint_as_float(v[4 * j + 1]) + xg[1][j]);
          gg[j] = tanh_approx(__uint_as_float(v[4 * j + 2]) + xg[2][j]);
          go[j] = sigmoid_fast(__uint_as_float(v[4 * j + 3]) + xg[3][j]);
          const float cn = fmaf(gf[j], c[ch * 8 + j], gi[j] * gg[j]);
          c[ch * 8 + j] = cn;
          hv[j] = go[j] * tanh_approx(cn);
        }
        if (ok) {
          const int u = u0 + ch * 8;
          uint4 w;
          w.x = pack_bf16x2(hv[0], hv[1]);
          w.y = pack_bf16x2(hv[2], hv[3]);
          w.z = pack_bf16x2(hv[4], hv[5]);
          w.w = pack_bf16x2(hv[6], hv[7]);
          *reinterpret_cast<uint4*>(a.hbuf + bt * H + u) = w;
          if (a.y) {
            if (a.y_dtype == 0) {
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + bt * a.y_ld + u) = w;
            } else {
              float* yp = reinterpret_cast<float*>(a.y) + bt * a.y_ld + u;
              *reinterpret_cast<float4*>(yp) = make_float4(hv[0], hv[1], hv[2], hv[3]);
              *reinterpret_cast<float4*>(yp + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
            }
          }
          if (a.gates) {
            float* gp = a.gates + bt * (4 * H) + u;
            *reinterpret_cast<float4*>(gp) = make_float4(gi[0], gi[1], gi[2], gi[3]);
            *reinterpret_cast<float4*>(gp + 4) = make_float4(gi[4], gi[5], gi[6], gi[7]);
            *reinterpret_cast<float4*>(gp + H) = make_float4(gf[0], gf[1], gf[2], gf[3]);
            *reinterpret_cast<float4*>(gp + H + 4) = make_float4(gf[4], gf[5], gf[6], gf[7]);
            *reinterpret_cast<float4*>(gp + 2 * H) = make_float4(gg[0], gg[1], gg[2], gg[3]);
            *reinterpret_cast<float4*>(gp + 2 * H + 4) = make_float4(gg[4], gg[5], gg[6], gg[7]);
            *reinterpret_cast<float4*>(gp + 3 * H) = make_float4(go[0], go[1], go[2], go[3]);
            *reinterpret_cast<float4*>(gp + 3 * H + 4) = make_float4(go[4], go[5], go[6], go[7]);
            float* cp = a.c_all + bt * H + u;
            *reinterpret_cast<float4*>(cp) = make_float4(c[ch * 8], c[ch * 8 + 1], c[ch * 8 + 2], c[ch * 8 + 3]);
            *reinterpret_cast<float4*>(cp + 4) = make_float4(c[ch * 8 + 4], c[ch * 8 + 5], c[ch * 8 + 6], c[ch * 8 + 7]);
          }
          if (t == a.T - 1 && a.h_out) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              a.h_out[(long long)b * H + u + j] = hv[j];
              a.c_out[(long long)b * H + u + j] = c[ch * 8 + j];
            }
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_all();   // h_t in global memory is the peers' next A operand (TMA)
      __syncwarp();
      if (lane == 0) mbar_arrive(x_empty);
      cluster_arrive_release();
      cluster_wait_acquire();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------------------
// back-propagation through time
// ------------------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(LTC_CLUSTER, 1, 1) __launch_bounds__(LTC_THREADS, 1) lstm_bwd_tc_kernel(const __grid_constant__ LstmBwdTcArgs a) {
  constexpr int H = LTC_H, U = LTC_U, SA = LTC_BWD_STAGE_ATOMS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                          // 16 atoms [32 n][128 B]
  uint8_t* sA = sW + 16 * U * 128;             // 2 stages x 4 atoms [128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 2 * SA * LTC_ATOM_A);
  uint64_t* w_full = bars;
  uint64_t* full = bars + 1;     // [2]
  uint64_t* empty = bars + 3;    // [2]
  uint64_t* acc_full = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int b0 = (blockIdx.x / LTC_CLUSTER) * LTC_ROWS;
  const int u0 = rank * U;
  constexpr uint32_t IDESC = umma_idesc(UMMA_FMT_BF16, U, 0, 0, 128);

  if (warp == 5) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      mbar_init(acc_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&a.tmW);
    tma_prefetch_desc(&a.tmD);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_arrive_release();
  cluster_wait_acquire();

  // Step order: t = T-1 .. 0.  Iteration `it` first produces d gates_t (needs dh from the MMAs of iteration it - 1), exchanges it
  // through global memory, then computes dh_{t-1} = d gates_t @ W_hh for the next iteration.
  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 16 * U * 128);
      for (int k = 0; k < 16; ++k) tma_load_2d(sW + k * U * 128, &a.tmW, w_full, k * 64, u0);
    }
    uint32_t n = 0;   // stages issued so far
    for (int it = 0; it < a.T; ++it) {
      const int t = a.T - 1 - it;
      __syncwarp();
      cluster_arrive_release();   // every CTA has stored its slice of d gates_t
      cluster_wait_acquire();
      if (lane == 0 && t > 0) {   // dh_{-1} is never needed
        fence_proxy_async_all();
        for (int q = 0; q < 16 / SA; ++q, ++n) {
          const int s = n & 1;
          if (n >= 2) mbar_wait(&empty[s], ((n >> 1) - 1) & 1);
          mbar_arrive_expect_tx(&full[s], SA * LTC_ATOM_A);
          for (int k = 0; k < SA; ++k) tma_load_3d(sA + (s * SA + k) * LTC_ATOM_A, &a.tmD, &full[s], (q * SA + k) * 64, t, b0);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) mbar_wait(w_full, 0);
    __syncwarp();
    uint32_t n = 0;
    for (int it = 0; it < a.T; ++it) {
      const int t = a.T - 1 - it;
      __syncwarp();
      cluster_arrive_release();
      cluster_wait_acquire();
      if (lane == 0 && t > 0) {
        for (int q = 0; q < 16 / SA; ++q, ++n) {
          const int s = n & 1;
          mbar_wait(&full[s], (n >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < SA; ++k) {
            const uint64_t adesc = umma_smem_desc_sw128(smem_u32(sA + (s * SA + k) * LTC_ATOM_A));
            const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(sW + (q * SA + k) * U * 128));
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_f16(tmem_base, adesc + 2 * j, bdesc + 2 * j, IDESC, (q | k | j) != 0);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(acc_full);
      }
    }
  } else {
    const int row = warp * 32 + lane;
    const int b = b0 + row;
    const bool ok = b < a.B;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float dc_next[U];
#pragma unroll
    for (int i = 0; i < U; ++i) dc_next[i] = 0.f;
    for (int it = 0; it < a.T; ++it) {
      const int t = a.T - 1 - it;
      uint32_t dh[U];
      if (it > 0) {
        mbar_wait(acc_full, (it - 1) & 1);
        tc_fence_after();
        tmem_ld32(taddr, dh);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < U; ++i) dh[i] = 0u;
      }
      if (ok) {
        const long long bt = (long long)b * a.T + t;
        const float* gp = a.gates + bt * (4 * H) + u0;
        const float* cp = a.c_all + bt * H + u0;
        const float* dyp = a.dy + bt * a.dy_ld + u0;
        float* dp = a.dgates + bt * (4 * H) + u0;
        __nv_bfloat16* db = a.dgb + bt * (4 * H) + u0;
#pragma unroll
        for (int q = 0; q < U / 4; ++q) {
          const float4 ig = *reinterpret_cast<const float4*>(gp + 4 * q), fg = *reinterpret_cast<const float4*>(gp + H + 4 * q);
          const float4 gg = *reinterpret_cast<const float4*>(gp + 2 * H + 4 * q), og = *reinterpret_cast<const float4*>(gp + 3 * H + 4 * q);
          const float4 ct = *reinterpret_cast<const float4*>(cp + 4 * q);
          const float4 cprev = t > 0 ? *reinterpret_cast<const float4*>(cp - H + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 dy4 = *reinterpret_cast<const float4*>(dyp + 4 * q);
          const float i_[4] = {ig.x, ig.y, ig.z, ig.w}, f_[4] = {fg.x, fg.y, fg.z, fg.w}, g_[4] = {gg.x, gg.y, gg.z, gg.w};
          const float o_[4] = {og.x, og.y, og.z, og.w}, c_[4] = {ct.x, ct.y, ct.z, ct.w}, cp_[4] = {cprev.x, cprev.y, cprev.z, cprev.w};
          const float dy_[4] = {dy4.x, dy4.y, dy4.z, dy4.w};
          float d0[4], d1[4], d2[4], d3[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float tc = tanh_approx(c_[j]);
            const float dhv = dy_[j] + __uint_as_float(dh[4 * q + j]);
            const float dc = dc_next[4 * q + j] + dhv * o_[j] * (1.f - tc * tc);
            d0[j] = dc * g_[j] * i_[j] * (1.f - i_[j]);
            d1[j] = dc * cp_[j] * f_[j] * (1.f - f_[j]);
            d2[j] = dc * i_[j] * (1.f - g_[j] * g_[j]);
            d3[j] = dhv * tc * o_[j] * (1.f - o_[j]);
            dc_next[4 * q + j] = dc * f_[j];
          }
          *reinterpret_cast<float4*>(dp + 4 * q) = make_float4(d0[0], d0[1], d0[2], d0[3]);
          *reinterpret_cast<float4*>(dp + H + 4 * q) = make_float4(d1[0], d1[1], d1[2], d1[3]);
          *reinterpret_cast<float4*>(dp + 2 * H + 4 * q) = make_float4(d2[0], d2[1], d2[2], d2[3]);
          *reinterpret_cast<float4*>(dp + 3 * H + 4 * q) = make_float4(d3[0], d3[1], d3[2], d3[3]);
          *reinterpret_cast<uint2*>(db + 4 * q) = make_uint2(pack_bf16x2(d0[0], d0[1]), pack_bf16x2(d0[2], d0[3]));
          *reinterpret_cast<uint2*>(db + H + 4 * q) = make_uint2(pack_bf16x2(d1[0], d1[1]), pack_bf16x2(d1[2], d1[3]));
          *reinterpret_cast<uint2*>(db + 2 * H + 4 * q) = make_uint2(pack_bf16x2(d2[0], d2[1]), pack_bf16x2(d2[2], d2[3]));
          *reinterpret_cast<uint2*>(db + 3 * H + 4 * q) = make_uint2(pack_bf16x2(d3[0], d3[1]), pack_bf16x2(d3[2], d3[3]));
        }
      }
      tc_fence_before();
      fence_proxy_async_all();
      __syncwarp();
      cluster_arrive_release();
      cluster_wait_acquire();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 32);
}

}  // namespace vt
