// Image preprocessing on the GPU: the reference's utils/image_utils.py:5-23 --
// torchvision Resize((S, S)) of a PIL RGB image, ToTensor, ImageNet Normalize -- for an already decoded RGB8 image.
// Resize is Pillow's two-pass antialiased bilinear resampling on 8-bit pixels (libImaging/Resample.c: triangle filter
// stretched by the scale factor, double-precision weights rounded to 22-bit fixed point, uint8 intermediate between the
// horizontal and the vertical pass); it is integer arithmetic, so the result is bit-identical to Pillow's, and the
// float32 tail ((u8 / 255 - mean) / std with IEEE division) is bit-identical to torchvision's.
//   pass 1: tmp[y][xx][c] = clip8((2^21 + sum_t in[y][x0(xx)+t][c] * kx[xx][t]) >> 22)          one thread per (y, xx)
//   pass 2: out[c][yy][xx] = ((clip8((2^21 + sum_t tmp[y0(yy)+t][xx][c] * ky[yy][t]) >> 22) / 255) - mean[c]) / std[c]
// The coefficient tables depend only on (input size, output size); they are computed on the host in double precision
// exactly as Pillow does and cached on the device by the engine.
#include <cmath>
#include <vector>
#include "kernels.h"
#include "common.cuh"

namespace xn {

constexpr int kPrecisionBits = 32 - 8 - 2;

void resample_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk, int* ksize_out) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  bounds.assign((size_t)out_size * 2, 0);
  kk.assign((size_t)out_size * ksize, 0);
  std::vector<double> w(ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < ksize; ++x) w[x] = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double v = (x + xmin - center + 0.5) * ss;
      if (v < 0.0) v = -v;
      const double wv = v < 1.0 ? 1.0 - v : 0.0;
      w[x] = wv;
      ww += wv;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) w[x] /= ww;
    for (int x = 0; x < ksize; ++x)
      kk[(size_t)xx * ksize + x] = w[x] < 0 ? (int)(-0.5 + w[x] * (1 << kPrecisionBits)) : (int)(0.5 + w[x] * (1 << kPrecisionBits));
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  *ksize_out = ksize;
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256) resample_h_kernel(const uint8_t* __restrict__ in, int H, int W, int S,
                                                         const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                         uint8_t* __restrict__ tmp) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (xx >= S) return;
  const int x0 = bounds[2 * xx], n = bounds[2 * xx + 1];
  const int* k = kk + (size_t)xx * ksize;
  const uint8_t* row = in + ((size_t)y * W + x0) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int t = 0; t < n; ++t) {
    const int kv = k[t];
    s0 += row[3 * t] * kv; s1 += row[3 * t + 1] * kv; s2 += row[3 * t + 2] * kv;
  }
  uint8_t* o = tmp + ((size_t)y * S + xx) * 3;
  o[0] = (uint8_t)clip8(s0); o[1] = (uint8_t)clip8(s1); o[2] = (uint8_t)clip8(s2);
}

__global__ void __launch_bounds__(256) resample_v_norm_kernel(const uint8_t* __restrict__ tmp, int S,
                                                              const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                              float* __restrict__ out) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x, yy = blockIdx.y;
  if (xx >= S) return;
  const int y0 = bounds[2 * yy], n = bounds[2 * yy + 1];
  const int* k = kk + (size_t)yy * ksize;
  int s[3] = {1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1)};
  for (int t = 0; t < n; ++t) {
    const uint8_t* p = tmp + ((size_t)(y0 + t) * S + xx) * 3;
    const int kv = k[t];
    s[0] += p[0] * kv; s[1] += p[1] * kv; s[2] += p[2] * kv;
  }
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __fdiv_rn((float)clip8(s[c]), 255.0f);                 // ToTensor
    out[((size_t)c * S + yy) * S + xx] = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);   // Normalize
  }
}

// ---- batched form: one launch pair for N images of arbitrary sizes.  Every image has an item record on the device
// (source, intermediate and output pointers, its two coefficient tables); blockIdx.z selects the image, rows beyond an
// image's height exit at once (the grid is sized for the tallest image of the batch).
__global__ void __launch_bounds__(256) resample_h_batch_kernel(const PreItem* __restrict__ items, int S) {
  const PreItem it = items[blockIdx.z];
  const int xx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (y >= it.H || xx >= S) return;
  const int x0 = it.bx[2 * xx], n = it.bx[2 * xx + 1];
  const int* k = it.kx + (size_t)xx * it.ksx;
  const uint8_t* row = it.src + ((size_t)y * it.W + x0) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int t = 0; t < n; ++t) {
    const int kv = k[t];
    s0 += row[3 * t] * kv; s1 += row[3 * t + 1] * kv; s2 += row[3 * t + 2] * kv;
  }
  uint8_t* o = it.tmp + ((size_t)y * S + xx) * 3;
  o[0] = (uint8_t)clip8(s0); o[1] = (uint8_t)clip8(s1); o[2] = (uint8_t)clip8(s2);
}

__global__ void __launch_bounds__(256) resample_v_norm_batch_kernel(const PreItem* __restrict__ items, int S) {
  const PreItem it = items[blockIdx.z];
  const int xx = blockIdx.x * blockDim.x + threadIdx.x, yy = blockIdx.y;
  if (xx >= S) return;
  const int y0 = it.by[2 * yy], n = it.by[2 * yy + 1];
  const int* k = it.ky + (size_t)yy * it.ksy;
  int s[3] = {1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1)};
  for (int t = 0; t < n; ++t) {
    const uint8_t* p = it.tmp + ((size_t)(y0 + t) * S + xx) * 3;
    const int kv = k[t];
    s[0] += p[0] * kv; s[1] += p[1] * kv; s[2] += p[2] * kv;
  }
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __fdiv_rn((float)clip8(s[c]), 255.0f);                 // ToTensor
    it.out[((size_t)c * S + yy) * S + xx] = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);   // Normalize
  }
}

cudaError_t launch_preprocess_rgb8_batch(const PreItem* items_dev, int n, int max_h, int S, cudaStream_t st) {
  if (n <= 0 || max_h <= 0 || S <= 0 || n > 65535 || max_h > 65535) return cudaErrorInvalidValue;
  resample_h_batch_kernel<<<dim3((S + 255) / 256, max_h, n), 256, 0, st>>>(items_dev, S);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  resample_v_norm_batch_kernel<<<dim3((S + 255) / 256, S, n), 256, 0, st>>>(items_dev, S);
  return cudaGetLastError();
}

cudaError_t launch_preprocess_rgb8(const uint8_t* rgb_dev, int H, int W, int S, const int* bounds_x, const int* kk_x, int ksize_x,
                                   const int* bounds_y, const int* kk_y, int ksize_y, uint8_t* tmp_dev, float* out_dev,
                                   cudaStream_t st) {
  if (H <= 0 || W <= 0 || S <= 0) return cudaErrorInvalidValue;
  resample_h_kernel<<<dim3((S + 255) / 256, H), 256, 0, st>>>(rgb_dev, H, W, S, bounds_x, kk_x, ksize_x, tmp_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  resample_v_norm_kernel<<<dim3((S + 255) / 256, S), 256, 0, st>>>(tmp_dev, S, bounds_y, kk_y, ksize_y, out_dev);
  return cudaGetLastError();
}

}  // namespace xn
