// pad_and_resize_for_siglip on the GPU: the camera-frame preprocessing in front of the controller in the deployment script
// (scripts/utils_eef.py:44-77, called at scripts/franka_inference_eef.py:329-330 and data/create_controller_dataset_episode.py:198):
// zero-pad the H x W x C uint8 frame to a centred square, then cv2.resize(..., (target, target), interpolation=cv2.INTER_AREA).
//
// OpenCV (un-vendored dependency; restated from its published algorithm, imgproc/resize.cpp, and pinned bit-for-bit against
// cv2 4.13 outputs in tests/golden/resize_*.npz) computes INTER_AREA down-scaling in two ways:
//   * integer scale factor: every output pixel is the sum of its scale x scale block, (sum + 2) >> 2 for 2 x 2, otherwise
//     saturate_cast<uchar>(sum * (1.f / area));
//   * fractional scale >= 1: per axis a table of (source index, weight) entries -- a partial first cell, whole cells of weight
//     1 / cellWidth, a partial last cell, cellWidth = min(scale, size - dx * scale); per source row a float row buffer
//     buf[dx] += S[sx] * alpha in table order, per output row sum = beta_0 * buf_0, then sum += beta_j * buf_j, and the result is
//     saturate_cast<uchar>(sum) = round-half-even, clamped.  Every multiply and add is a separately rounded fp32 operation.
// One thread per output element reproduces exactly that operation order (__fmul_rn / __fadd_rn: no FMA contraction), so the result
// is bit-identical.  The padding is never materialised: canvas coordinates outside the frame read as zero.
// Up-scaling with INTER_AREA (frames smaller than `target`) is a different OpenCV code path and is rejected by the host.
#pragma once
#include <cstdint>

namespace vt {

struct ResizeArgs {
  const uint8_t* src;      // [n][h][w][c]
  uint8_t* dst;            // [n][target][target][c]
  int n, h, w, c, target;
  int side;                // canvas side = max(h, w)
  int pad_y, pad_x;        // (side - h) / 2, (side - w) / 2
  int iscale;              // > 0: integer scale factor (fast path); 0: table path
  float inv_area;          // 1.f / (iscale * iscale)
  const int* tab_off;      // [target + 1] entry ranges of the axis table (square canvas: the same table for both axes)
  const int* tab_si;       // source index of an entry
  const float* tab_alpha;  // weight of an entry
};

__global__ void __launch_bounds__(256) pad_resize_area_kernel(const ResizeArgs a) {
  const long long total = (long long)a.n * a.target * a.target * a.c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % a.c);
    long long r = i / a.c;
    const int dx = (int)(r % a.target);
    r /= a.target;
    const int dy = (int)(r % a.target);
    const int img = (int)(r / a.target);
    const uint8_t* S = a.src + (long long)img * a.h * a.w * a.c;
    auto px = [&](int sy, int sx) -> int {      // canvas pixel: zero outside the centred frame
      const int y = sy - a.pad_y, x = sx - a.pad_x;
      return (y >= 0 && y < a.h && x >= 0 && x < a.w) ? (int)S[((long long)y * a.w + x) * a.c + ch] : 0;
    };
    int out;
    if (a.iscale > 0) {
      int s = 0;
      for (int yy = 0; yy < a.iscale; ++yy)
        for (int xx = 0; xx < a.iscale; ++xx) s += px(dy * a.iscale + yy, dx * a.iscale + xx);
      out = a.iscale == 2 ? (s + 2) >> 2 : __float2int_rn(__fmul_rn((float)s, a.inv_area));
    } else {
      float sum = 0.f;
      const int y0 = a.tab_off[dy], y1 = a.tab_off[dy + 1], x0 = a.tab_off[dx], x1 = a.tab_off[dx + 1];
      for (int j = y0; j < y1; ++j) {
        const int sy = a.tab_si[j];
        const float beta = a.tab_alpha[j];
        float buf = 0.f;
        for (int k = x0; k < x1; ++k) buf = __fadd_rn(buf, __fmul_rn((float)px(sy, a.tab_si[k]), a.tab_alpha[k]));
        sum = j == y0 ? __fmul_rn(beta, buf) : __fadd_rn(sum, __fmul_rn(beta, buf));
      }
      out = __float2int_rn(sum);
    }
    a.dst[i] = (uint8_t)min(max(out, 0), 255);
  }
}

}  // namespace vt
