// nn.LSTM recurrence (gate order i,f,g,o; lstm_step_controller.py:66-73,196-204,268-271).  The input projections
// W_ih x_t + b_ih + b_hh of all T steps are one tcgen05 GEMM (vt_gemm.cuh); this kernel runs the sequential part.
// One CTA owns LSTM_ROWS batch rows for all T steps (rows are independent, so no grid-wide sync); thread j owns
// hidden unit j.  W_hh is stored transposed ([H][4H], k-major) so that a warp reads 128 contiguous bytes per k.
#pragma once
#include "vt_elem.cuh"

namespace vt {

constexpr int LSTM_ROWS = 4;

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int H>
__global__ void __launch_bounds__(H) lstm_seq_kernel(const float* __restrict__ xw, const float* __restrict__ w_hh_t,
                                                     float* __restrict__ h_state, float* __restrict__ c_state,
                                                     void* __restrict__ y, int y_dtype, long long y_ld, long long y_plane,
                                                     int B, int T) {
  __shared__ float sh[LSTM_ROWS][H];
  const int j = threadIdx.x;
  const int b0 = blockIdx.x * LSTM_ROWS;
  float c[LSTM_ROWS], hreg[LSTM_ROWS];
#pragma unroll
  for (int r = 0; r < LSTM_ROWS; ++r) {
    const bool ok = b0 + r < B;
    c[r] = ok ? c_state[(long long)(b0 + r) * H + j] : 0.f;
    hreg[r] = ok ? h_state[(long long)(b0 + r) * H + j] : 0.f;
    sh[r][j] = hreg[r];
  }
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    float acc[LSTM_ROWS][4];
#pragma unroll
    for (int r = 0; r < LSTM_ROWS; ++r) {
      const bool ok = b0 + r < B;
      const float* g = xw + ((long long)(b0 + r) * T + t) * (4 * H);
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[r][q] = ok ? g[q * H + j] : 0.f;
    }
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
      const float* wr = w_hh_t + (long long)k * (4 * H) + j;
      const float w0 = __ldg(wr), w1 = __ldg(wr + H), w2 = __ldg(wr + 2 * H), w3 = __ldg(wr + 3 * H);
#pragma unroll
      for (int r = 0; r < LSTM_ROWS; ++r) {
        const float hv = sh[r][k];
        acc[r][0] = fmaf(hv, w0, acc[r][0]);
        acc[r][1] = fmaf(hv, w1, acc[r][1]);
        acc[r][2] = fmaf(hv, w2, acc[r][2]);
        acc[r][3] = fmaf(hv, w3, acc[r][3]);
      }
    }
    __syncthreads();  // all reads of sh for this step are done
#pragma unroll
    for (int r = 0; r < LSTM_ROWS; ++r) {
      const float ig = sigmoid_f(acc[r][0]), fg = sigmoid_f(acc[r][1]), gg = tanhf(acc[r][2]), og = sigmoid_f(acc[r][3]);
      c[r] = fg * c[r] + ig * gg;
      hreg[r] = og * tanhf(c[r]);
      sh[r][j] = hreg[r];
      if (b0 + r < B) store_val(y, y_dtype, ((long long)(b0 + r) * T + t) * y_ld + j, y_plane, hreg[r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < LSTM_ROWS; ++r) {
    if (b0 + r < B) {
      c_state[(long long)(b0 + r) * H + j] = c[r];
      h_state[(long long)(b0 + r) * H + j] = hreg[r];
    }
  }
}

}  // namespace vt
