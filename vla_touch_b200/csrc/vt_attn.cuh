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
