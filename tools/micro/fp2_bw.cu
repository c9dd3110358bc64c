// Developer micro-benchmark (GPU box): issue rate / latency of packed f32x2 math vs scalar, and of the GELU epilogue math,
// with 2 warps per SM sub-partition (the epilogue configuration).
#include <cstdio>
#include "vt_gemm.cuh"
using namespace vt;

template <int MODE, int CH = 16>
__global__ void __launch_bounds__(256, 1) k(int iters, long long* cycles, float* sink, float seed) {
  float2 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = make_float2(seed + i + threadIdx.x * 1e-3f, seed - i);
  const float2 b = make_float2(seed * 0.5f, seed * 0.25f), s = make_float2(1.0001f, 0.9999f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) x[i] = ffma2(x[i], s, b);                                                   // packed FMA, 16 independent chains
      else if (MODE == 1) { x[i].x = fmaf(x[i].x, s.x, b.x); x[i].y = fmaf(x[i].y, s.y, b.y); }  // scalar FMA, 32 chains
      else if (MODE == 2) x[i] = gelu_fast2(fadd2(x[i], b));                                      // packed GELU
      else if (MODE == 3) { x[i].x = gelu_fast(x[i].x + b.x); x[i].y = gelu_fast(x[i].y + b.y); } // scalar GELU
      else if (MODE == 4) x[i] = mish2(x[i]);
      else { x[i].x = mish_f(x[i].x); x[i].y = mish_f(x[i].y); }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += x[i].x + x[i].y;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
}

template <int MODE, int CH = 16>
void run(const char* name) {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  k<MODE, CH><<<148, 256>>>(iters, cyc, sink, 0.3f);
  cudaDeviceSynchronize();
  k<MODE, CH><<<148, 256>>>(iters, cyc, sink, 0.3f);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double elems = (double)iters * 2 * CH * 256;   // elements per SM
  printf("%-14s %9lld cycles  %6.2f elements/cycle/SM   %6.1f cycles per 32-element step per warp (%s)\n", name, h[0], elems / h[0],
         (double)h[0] / iters, cudaGetErrorString(e));
}

int main() {
  run<0>("FFMA2");
  run<1>("FFMA scalar");
  run<2>("GELU packed");
  run<3>("GELU scalar");
  run<4>("Mish packed");
  run<5>("Mish scalar");
  run<4, 8>("Mish packed, 8 chains");
  run<4, 4>("Mish packed, 4 chains");
  run<4, 2>("Mish packed, 2 chains");
  run<2, 4>("GELU packed, 4 chains");
  run<2, 2>("GELU packed, 2 chains");
  return 0;
}
