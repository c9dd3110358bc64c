// Developer micro-benchmark (GPU box): tcgen05.ld throughput per SM for the shapes the epilogues use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vla_touch_b200/csrc -o /tmp/tmem_bw tools/micro/tmem_bw.cu -lcuda && /tmp/tmem_bw
#include <cstdio>
#include "vt_ptx.cuh"
using namespace vt;

template <int SHAPE, int INFLIGHT>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* cycles, unsigned* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (uint32_t)((i * 64 + (warp >> 2) * 32) & 255);
    if constexpr (SHAPE == 32) {
      uint32_t v[32], w[32];
      tmem_ld32(base + col, v);
      if (INFLIGHT == 2) tmem_ld32(base + ((col + 128) & 255), w);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) acc ^= v[q];
      if (INFLIGHT == 2) {
#pragma unroll
        for (int q = 0; q < 32; ++q) acc ^= w[q];
      }
    } else if constexpr (SHAPE == 64) {
      uint32_t v[64];
      tmem_ld64(base + col, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 64; ++q) acc ^= v[q];
    } else {
      uint32_t v[16];
      tmem_ld16(base + col, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 16; ++q) acc ^= v[q];
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

template <int SHAPE, int INFLIGHT>
void run(int warps, const char* name) {
  long long* cyc; unsigned* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  k<SHAPE, INFLIGHT><<<148, warps * 32>>>(iters, cyc, sink);
  cudaDeviceSynchronize();
  k<SHAPE, INFLIGHT><<<148, warps * 32>>>(iters, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes = (double)iters * warps * 32 * SHAPE * 4 * INFLIGHT;
  printf("%-28s warps %2d: %8lld cycles, %6.1f B/cycle/SM, %6.1f cycles per load per warp  (%s)\n", name, warps, h[0], bytes / h[0],
         (double)h[0] / iters / INFLIGHT, cudaGetErrorString(e));
}

int main() {
  for (int w : {4, 8, 16}) {
    if (w == 4) { run<32, 1>(4, "32x32b.x32, 1 in flight"); run<32, 2>(4, "32x32b.x32, 2 in flight"); run<64, 1>(4, "32x32b.x64"); run<16, 1>(4, "32x32b.x16"); }
    if (w == 8) { run<32, 1>(8, "32x32b.x32, 1 in flight"); run<32, 2>(8, "32x32b.x32, 2 in flight"); run<64, 1>(8, "32x32b.x64"); run<16, 1>(8, "32x32b.x16"); }
    if (w == 16) { run<32, 1>(16, "32x32b.x32, 1 in flight"); run<64, 1>(16, "32x32b.x64"); }
  }
  return 0;
}
