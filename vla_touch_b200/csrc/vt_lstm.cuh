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

// Training forward: the same recurrence from a zero initial state (TactileLSTMController.forward, lstm_step_controller.py:196-204)
// that also keeps what back-propagation through time needs: the activated gates (i, f, g, o) and the cell state of every step.
template <int H>
__global__ void __launch_bounds__(H) lstm_seq_train_kernel(const float* __restrict__ xw, const float* __restrict__ w_hh_t,
                                                           void* __restrict__ y, int y_dtype, long long y_ld,
                                                           float* __restrict__ gates, float* __restrict__ c_all, int B, int T) {
  __shared__ float sh[LSTM_ROWS][H];
  const int j = threadIdx.x;
  const int b0 = blockIdx.x * LSTM_ROWS;
  float c[LSTM_ROWS];
#pragma unroll
  for (int r = 0; r < LSTM_ROWS; ++r) {
    c[r] = 0.f;
    sh[r][j] = 0.f;
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
    __syncthreads();
#pragma unroll
    for (int r = 0; r < LSTM_ROWS; ++r) {
      const float ig = sigmoid_f(acc[r][0]), fg = sigmoid_f(acc[r][1]), gg = tanhf(acc[r][2]), og = sigmoid_f(acc[r][3]);
      c[r] = fg * c[r] + ig * gg;
      const float hv = og * tanhf(c[r]);
      sh[r][j] = hv;
      if (b0 + r < B) {
        const long long bt = (long long)(b0 + r) * T + t;
        store_val(y, y_dtype, bt * y_ld + j, 0, hv);
        float* gp = gates + bt * (4 * H);
        gp[j] = ig;
        gp[H + j] = fg;
        gp[2 * H + j] = gg;
        gp[3 * H + j] = og;
        c_all[bt * H + j] = c[r];
      }
    }
    __syncthreads();
  }
}

// Back-propagation through time of one LSTM layer (the sequential part; `lstm_loss_backward` of oracle/vt_oracle_bwd.py):
//   dh = dy_t + dh_next;  dc = dc_next + dh o (1 - tanh(c_t)^2)
//   d gates (pre-activation) = [dc g i(1-i), dc c_{t-1} f(1-f), dc i (1-g^2), dh tanh(c_t) o(1-o)]
//   dh_next = d gates @ W_hh;  dc_next = dc f
// The weight gradients (W_ih, W_hh, biases) and d x are GEMMs / column sums over the stored d gates of ALL steps.
// Same decomposition as the forward: LSTM_ROWS batch rows per CTA, thread j = hidden unit j; W_hh [4H][H] is read row by row
// (coalesced over j), the 4H gate derivatives of a row are exchanged through shared memory.
template <int H>
__global__ void __launch_bounds__(H) lstm_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_all,
                                                     const float* __restrict__ dy, long long dy_ld, const float* __restrict__ w_hh,
                                                     float* __restrict__ dgates, int B, int T) {
  __shared__ float sdg[LSTM_ROWS][4 * H];
  const int j = threadIdx.x;
  const int b0 = blockIdx.x * LSTM_ROWS;
  float dh_next[LSTM_ROWS], dc_next[LSTM_ROWS];
#pragma unroll
  for (int r = 0; r < LSTM_ROWS; ++r) dh_next[r] = dc_next[r] = 0.f;
  for (int t = T - 1; t >= 0; --t) {
#pragma unroll
    for (int r = 0; r < LSTM_ROWS; ++r) {
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
      if (b0 + r < B) {
        const long long bt = (long long)(b0 + r) * T + t;
        const float* gp = gates + bt * (4 * H);
        const float ig = gp[j], fg = gp[H + j], gg = gp[2 * H + j], og = gp[3 * H + j];
        const float ct = c_all[bt * H + j];
        const float cp = t > 0 ? c_all[(bt - 1) * H + j] : 0.f;
        const float tc = tanhf(ct);
        const float dh = dy[bt * dy_ld + j] + dh_next[r];
        const float dc = dc_next[r] + dh * og * (1.f - tc * tc);
        d0 = dc * gg * ig * (1.f - ig);
        d1 = dc * cp * fg * (1.f - fg);
        d2 = dc * ig * (1.f - gg * gg);
        d3 = dh * tc * og * (1.f - og);
        dc_next[r] = dc * fg;
        float* dp = dgates + bt * (4 * H);
        dp[j] = d0;
        dp[H + j] = d1;
        dp[2 * H + j] = d2;
        dp[3 * H + j] = d3;
      }
      sdg[r][j] = d0;
      sdg[r][H + j] = d1;
      sdg[r][2 * H + j] = d2;
      sdg[r][3 * H + j] = d3;
    }
    __syncthreads();
    float acc[LSTM_ROWS];
#pragma unroll
    for (int r = 0; r < LSTM_ROWS; ++r) acc[r] = 0.f;
#pragma unroll 4
    for (int m = 0; m < 4 * H; ++m) {
      const float w = __ldg(w_hh + (long long)m * H + j);
#pragma unroll
      for (int r = 0; r < LSTM_ROWS; ++r) acc[r] = fmaf(sdg[r][m], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < LSTM_ROWS; ++r) dh_next[r] = acc[r];
    __syncthreads();
  }
}

}  // namespace vt
