// Bandwidth-bound kernels: LayerNorm (+2x2 merge gather), patch embedding, static-expansion
// weight normalisation, selector mix, casts.  All reductions are fp32.
#include "kernels.h"
#include "common.cuh"
#include "mma16_frag.cuh"

namespace xn {

constexpr float kLnEps = 1e-5f;     // nn.LayerNorm default, used by every LN on the path
constexpr float kExpEps = 1e-9f;    // expansion eps, reference models/layers.py:106,208

// ------------------------------------------------------------------------------------------
// LayerNorm over the last dim: one warp per row, three L1-resident passes (mean, var, write).
// Replaces nn.LayerNorm at swin:342,355,645,778 and layers.py:108-109,210-212, End...:108-110.
// ------------------------------------------------------------------------------------------
template <typename OutT>
__device__ __forceinline__ void store4(OutT* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store4<bf16>(bf16* p, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
}

template <>
__device__ __forceinline__ void store4<f16>(f16* p, float a, float b, float c, float d) {
  __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
}

// Generic row LayerNorm: one warp per row, three L1-resident passes (used for the gathered PatchMerging rows and
// for widths without a specialised kernel).
template <typename OutT, typename SrcFn>
__device__ __forceinline__ void ln_row(SrcFn src, const float* __restrict__ g, const float* __restrict__ b,
                                       OutT* __restrict__ yr, int C, int lane) {
  float s = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = src(c);
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = src(c);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + kLnEps);
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = src(c);
    const float4 gg = *reinterpret_cast<const float4*>(g + c);
    const float4 bb = *reinterpret_cast<const float4*>(b + c);
    store4<OutT>(yr + c, (v.x - mean) * rstd * gg.x + bb.x, (v.y - mean) * rstd * gg.y + bb.y,
                 (v.z - mean) * rstd * gg.z + bb.z, (v.w - mean) * rstd * gg.w + bb.w);
  }
}

template <typename OutT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long ldx,
                                                        const float* __restrict__ g, const float* __restrict__ b,
                                                        OutT* __restrict__ y, long ldy, long rows, int C) {
  pdl_wait();
  pdl_trigger();
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  ln_row<OutT>([&](int c) { return *reinterpret_cast<const float4*>(xr + c); }, g, b, y + row * ldy, C,
               threadIdx.x & 31);
}

// Split-K consumer (decoder-step ff2 / reduce GEMMs, 16-bit modes): the GEMM ran as `nparts` K-slices that left fp32
// partial products part[s][row][c]; this kernel finishes it -- x = res + bias + sum_s part[s] in slice order (fixed, so
// results do not depend on scheduling), written to xout (may alias res) -- and, with g != nullptr, applies the
// LayerNorm that follows in the graph anyway (one warp per row, the row in registers; C % 128 == 0, C <= 1024).
template <typename OutT>
__global__ void __launch_bounds__(256) layernorm_sum_kernel(const float* __restrict__ part, int nparts, long pstride, long ldp,
                                                            const float* __restrict__ bias, const float* res, long ldr,
                                                            float* xout, long ldxo, const float* __restrict__ g,
                                                            const float* __restrict__ b, OutT* __restrict__ y, long ldy,
                                                            long rows, int C) {
  pdl_wait();
  pdl_trigger();
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float4 v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < C) {
      float4 a = bias ? *reinterpret_cast<const float4*>(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int sl = 0; sl < nparts; ++sl) {
        const float4 q = *reinterpret_cast<const float4*>(part + sl * pstride + row * ldp + c);
        a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
      }
      if (res) {
        const float4 r = *reinterpret_cast<const float4*>(res + row * ldr + c);
        a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
      }
      v[i] = a;
      *reinterpret_cast<float4*>(xout + row * ldxo + c) = a;
      s += (a.x + a.y) + (a.z + a.w);
    }
  }
  if (!g) return;
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (lane * 4 + i * 128 < C) {
      const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
      q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + kLnEps);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane * 4 + i * 128;
    if (c < C) {
      const float4 gg = *reinterpret_cast<const float4*>(g + c), bb = *reinterpret_cast<const float4*>(b + c);
      store4<OutT>(y + row * ldy + c, (v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y,
                   (v[i].z - mean) * rstd * gg.z + bb.z, (v[i].w - mean) * rstd * gg.w + bb.w);
    }
  }
}
template <typename OutT>
cudaError_t launch_layernorm_sum(const float* part, int nparts, long pstride, long ldp, const float* bias, const float* res, long ldr,
                                 float* xout, long ldxo, const float* gamma, const float* beta, OutT* y, long ldy, long rows, int C,
                                 cudaStream_t st) {
  if (C % 128 || C > 1024 || (ldp & 3) || (ldr & 3) || (ldxo & 3) || (pstride & 3) || (gamma && (ldy & 3))) return cudaErrorInvalidValue;
  launch_k(layernorm_sum_kernel<OutT>, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, part, nparts, pstride, ldp, bias, res, ldr, xout,
           ldxo, gamma, beta, y, ldy, rows, C);
  return cudaGetLastError();
}
template cudaError_t launch_layernorm_sum<bf16>(const float*, int, long, long, const float*, const float*, long, float*, long, const float*, const float*, bf16*, long, long, int, cudaStream_t);
template cudaError_t launch_layernorm_sum<f16>(const float*, int, long, long, const float*, const float*, long, float*, long, const float*, const float*, f16*, long, long, int, cudaStream_t);

// Specialised persistent LayerNorm for C = 128*VPL' widths: the row lives in exactly VPL float4 registers per lane,
// gamma/beta are loaded once per warp, each warp strides over rows and keeps RIF rows in flight (memory-level
// parallelism), so HBM latency is covered without relying on thousands of tiny CTAs.
template <typename OutT, int VPL, int RIF>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const float* __restrict__ x, long ldx,
                                                             const float* __restrict__ g, const float* __restrict__ b,
                                                             OutT* __restrict__ y, long ldy, long rows, int C) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long warp_g = (long)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (long)gridDim.x * 8;
  constexpr bool kAffineInRegs = VPL <= 3;      // wide rows re-read gamma/beta (L1-resident) to keep occupancy up
  float4 gg[kAffineInRegs ? VPL : 1], bb[kAffineInRegs ? VPL : 1];
  if (kAffineInRegs) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane * 4 + i * 128;
      gg[i] = c < C ? *reinterpret_cast<const float4*>(g + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      bb[i] = c < C ? *reinterpret_cast<const float4*>(b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float inv_c = 1.0f / (float)C;
  for (long r0 = warp_g * RIF; r0 < rows; r0 += nwarps * RIF) {
    float4 v[RIF][VPL];
#pragma unroll
    for (int k = 0; k < RIF; ++k)
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = lane * 4 + i * 128;
        v[k][i] = (r0 + k < rows && c < C) ? *reinterpret_cast<const float4*>(x + (r0 + k) * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int k = 0; k < RIF; ++k) {
      if (r0 + k >= rows) break;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) s += (v[k][i].x + v[k][i].y) + (v[k][i].z + v[k][i].w);   // padding lanes hold zeros
      const float mean = warp_sum(s) * inv_c;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        if (lane * 4 + i * 128 < C) {
          const float d0 = v[k][i].x - mean, d1 = v[k][i].y - mean, d2 = v[k][i].z - mean, d3 = v[k][i].w - mean;
          q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) * inv_c + kLnEps);
      OutT* yr = y + (r0 + k) * ldy;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = lane * 4 + i * 128;
        if (c < C) {
          const float4 ga = kAffineInRegs ? gg[kAffineInRegs ? i : 0] : *reinterpret_cast<const float4*>(g + c);
          const float4 be = kAffineInRegs ? bb[kAffineInRegs ? i : 0] : *reinterpret_cast<const float4*>(b + c);
          store4<OutT>(yr + c, (v[k][i].x - mean) * rstd * ga.x + be.x, (v[k][i].y - mean) * rstd * ga.y + be.y,
                       (v[k][i].z - mean) * rstd * ga.z + be.z, (v[k][i].w - mean) * rstd * ga.w + be.w);
        }
      }
    }
  }
}

template <typename OutT>
cudaError_t launch_layernorm(const float* x, long ldx, const float* gamma, const float* beta, OutT* y, long ldy,
                             long rows, int C, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  if ((C & 3) || (ldx & 3) || (ldy & 3)) return cudaErrorInvalidValue;
  const int vpl = (C + 127) / 128;
  if (rows >= 4096 && vpl <= 12) {
    const unsigned grid = 148 * 6;
#define XN_LN(V, R) launch_k(layernorm_rows_kernel<OutT, V, R>, dim3(grid), dim3(256), 0, st, x, ldx, gamma, beta, y, ldy, rows, C)
    if (vpl <= 2) XN_LN(2, 4);
    else if (vpl <= 3) XN_LN(3, 4);
    else if (vpl <= 4) XN_LN(4, 2);
    else if (vpl <= 6) XN_LN(6, 2);
    else XN_LN(12, 1);
#undef XN_LN
    return cudaGetLastError();
  }
  launch_k(layernorm_kernel<OutT>, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, x, ldx, gamma, beta, y, ldy, rows, C);
  return cudaGetLastError();
}
template cudaError_t launch_layernorm<float>(const float*, long, const float*, const float*, float*, long, long, int, cudaStream_t);
template cudaError_t launch_layernorm<bf16>(const float*, long, const float*, const float*, bf16*, long, long, int, cudaStream_t);
template cudaError_t launch_layernorm<f16>(const float*, long, const float*, const float*, f16*, long, long, int, cudaStream_t);

// PatchMerging gather + LN(4C): reference swin:482-501.  Output row (b,i,j) concatenates the
// input tokens (2i,2j), (2i+1,2j), (2i,2j+1), (2i+1,2j+1) in that order.
template <typename OutT>
__global__ void __launch_bounds__(256) merge_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                              const float* __restrict__ b, OutT* __restrict__ y,
                                                              int B, int H, int C) {
  pdl_wait();
  pdl_trigger();
  const int H2 = H / 2;
  const long rows = (long)B * H2 * H2;
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int bi = (int)(row / (H2 * H2)), rem = (int)(row % (H2 * H2));
  const int i = rem / H2, j = rem % H2;
  const float* base = x + (long)bi * H * H * C;
  auto src = [&](int c) {
    const int q = c / C, cc = c - q * C;
    const int dy = q & 1, dx = q >> 1;
    return *reinterpret_cast<const float4*>(base + ((long)(2 * i + dy) * H + (2 * j + dx)) * C + cc);
  };
  ln_row<OutT>(src, g, b, y + row * 4L * C, 4 * C, threadIdx.x & 31);
}

// The same with the 4C-wide row held in registers (VPL float4 per lane, C = 32 * VPL): the four source tokens are read
// once, all loads of a row in flight together, and the quadrant of a column is a compile-time division.  Summation
// order and formulas are those of ln_row, so the results are bit-identical to the generic kernel.
template <typename OutT, int VPL>
__global__ void __launch_bounds__(256) merge_layernorm_reg_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                                  const float* __restrict__ b, OutT* __restrict__ y,
                                                                  int B, int H) {
  pdl_wait();
  pdl_trigger();
  constexpr int C = 32 * VPL, C4 = 4 * C;
  const int H2 = H / 2, lane = threadIdx.x & 31;
  const long rows = (long)B * H2 * H2;
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int bi = (int)(row / (H2 * H2)), rem = (int)(row % (H2 * H2));
  const int i = rem / H2, j = rem % H2;
  const float* base = x + (((long)bi * H + 2 * i) * H + 2 * j) * C;
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int c = lane * 4 + k * 128, q = c / C, cc = c - q * C;
    v[k] = *reinterpret_cast<const float4*>(base + ((long)(q & 1) * H + (q >> 1)) * C + cc);
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  const float mean = warp_sum(s) / (float)C4;
  float qs = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const float d0 = v[k].x - mean, d1 = v[k].y - mean, d2 = v[k].z - mean, d3 = v[k].w - mean;
    qs += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(qs) / (float)C4 + kLnEps);
  OutT* yr = y + row * (long)C4;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int c = lane * 4 + k * 128;
    const float4 gg = *reinterpret_cast<const float4*>(g + c);
    const float4 bb = *reinterpret_cast<const float4*>(b + c);
    store4<OutT>(yr + c, (v[k].x - mean) * rstd * gg.x + bb.x, (v[k].y - mean) * rstd * gg.y + bb.y,
                 (v[k].z - mean) * rstd * gg.z + bb.z, (v[k].w - mean) * rstd * gg.w + bb.w);
  }
}

template <typename OutT>
cudaError_t launch_merge_layernorm(const float* x, const float* gamma, const float* beta, OutT* y, int B, int H,
                                   int C, cudaStream_t st) {
  if (C & 3) return cudaErrorInvalidValue;
  const long rows = (long)B * (H / 2) * (H / 2);
#define XN_MLN(V) if (C == 32 * V) { launch_k(merge_layernorm_reg_kernel<OutT, V>, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, x, gamma, beta, y, B, H); return cudaGetLastError(); }
  XN_MLN(6) XN_MLN(12) XN_MLN(24)
#undef XN_MLN
  launch_k(merge_layernorm_kernel<OutT>, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, x, gamma, beta, y, B, H, C);
  return cudaGetLastError();
}
template cudaError_t launch_merge_layernorm<float>(const float*, const float*, const float*, float*, int, int, int, cudaStream_t);
template cudaError_t launch_merge_layernorm<bf16>(const float*, const float*, const float*, bf16*, int, int, int, cudaStream_t);
template cudaError_t launch_merge_layernorm<f16>(const float*, const float*, const float*, f16*, int, int, int, cudaStream_t);

// ------------------------------------------------------------------------------------------
// PatchEmbed (reference swin:611-654): Conv2d(Cin->E, k=stride=P) + flatten + LayerNorm(E).
// One CTA per (image, patch row): the Cin*P input rows of that patch row are staged in smem
// (coalesced), the (E x Cin*P*P) filter bank is staged transposed, one warp per patch.
// ------------------------------------------------------------------------------------------
constexpr int kPeRows = 4;     // patch rows per CTA (amortises staging the filter bank)

__global__ void __launch_bounds__(256) patch_embed_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const float* __restrict__ g,
                                                          const float* __restrict__ be, float* __restrict__ out,
                                                          int Cin, int S, int P, int E) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  const int G = S / P, K = Cin * P * P;
  float* slab = sm;                       // [Cin*P][S]   one patch row of input pixels
  float* wt = slab + Cin * P * S;         // [K][E]       filter bank, transposed
  int* koff = reinterpret_cast<int*>(wt + K * E);   // [K] slab offset of tap k = (c, dy, dx)
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const int c = k / (P * P), r = k % (P * P);
    koff[k] = (c * P + r / P) * S + r % P;
  }
  const int groups = (G + kPeRows - 1) / kPeRows;
  const int b = blockIdx.x / groups, py0 = (blockIdx.x % groups) * kPeRows;
  // w is (E, Cin, P, P) = [e][k]; smem writes are contiguous in e (conflict-free), the strided reads hit L2
  for (int i = threadIdx.x; i < K * E; i += blockDim.x) {
    const int k = i / E, e = i % E;
    wt[i] = w[e * K + k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int EP = E / 32;                  // channels per lane (E % 32 == 0, E <= 256)
  float gam[8], bet[8], bia[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int e = lane + 32 * j;
    gam[j] = (j < EP) ? g[e] : 0.f; bet[j] = (j < EP) ? be[e] : 0.f; bia[j] = (j < EP) ? bias[e] : 0.f;
  }
  for (int pr = 0; pr < kPeRows && py0 + pr < G; ++pr) {
    const int py = py0 + pr;
    __syncthreads();
    for (int i = threadIdx.x; i < Cin * P * (S / 4); i += blockDim.x) {
      const int r = i / (S / 4), x4 = (i % (S / 4)) * 4, c = r / P, dy = r % P;
      *reinterpret_cast<float4*>(&slab[r * S + x4]) =
          *reinterpret_cast<const float4*>(&img[(((long)b * Cin + c) * S + (py * P + dy)) * S + x4]);
    }
    __syncthreads();
    constexpr int PB = 4;                  // patches per warp pass: each filter tap is read from smem once for PB patches
    for (int px0 = warp * PB; px0 < G; px0 += nw * PB) {
      float acc[PB][8];
#pragma unroll
      for (int q = 0; q < PB; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[q][j] = bia[j];
      for (int k = 0; k < K; ++k) {
        const int ko = koff[k];
        float v[PB];
#pragma unroll
        for (int q = 0; q < PB; ++q) v[q] = (px0 + q < G) ? slab[ko + (px0 + q) * P] : 0.f;
        const float* wk = wt + k * E + lane;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < EP) {
            const float wv = wk[32 * j];
#pragma unroll
            for (int q = 0; q < PB; ++q) acc[q][j] = fmaf(v[q], wv, acc[q][j]);
          }
      }
#pragma unroll
      for (int q = 0; q < PB; ++q) {
        const int px = px0 + q;
        if (px >= G) break;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < EP) s += acc[q][j];
        const float mean = warp_sum(s) / (float)E;
        float qv = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < EP) { const float dd = acc[q][j] - mean; qv += dd * dd; }
        const float rstd = 1.0f / sqrtf(warp_sum(qv) / (float)E + kLnEps);
        float* o = out + (((long)b * G + py) * G + px) * E;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < EP) o[lane + 32 * j] = (acc[q][j] - mean) * rstd * gam[j] + bet[j];
      }
    }
  }
}

// Patch width 4 (every Swin-L/384 call site): the four dx taps of a (channel, dy) row segment are one float4 in the
// slab, and the filter bank is pre-arranged as wq[(c*4+dy)][e] = float4 over dx (derived once per weight load), so four
// taps cost one LDS.128 per patch plus one per channel: 12 FMAs per shared-memory load with 6 patches x 6 channels in
// registers (the generic kernel above issues 10 loads per 24 FMAs and re-transposes the filter bank in every CTA).
constexpr int kPe4Patches = 6;

// x16 != nullptr (16-bit modes, folded LayerNorm): the kernel is also the PRODUCER of the first Swin block's norm1 -- it
// writes the rows rounded to 16 bits and their sum / sum of squares in the 64-bit fixed-point format of the tcgen05
// GEMM's producer epilogue (gemm_tcgen05.cu), so the first qkv GEMM runs against the folded weights like every other
// block and the stage-1 LayerNorm launch (the largest one: 453 MB read) disappears.
// EPT > 0: E / 32 known at compile time (Swin-L: 6), so the per-lane channel loops carry no guards (the runtime-EP
// form spent half of its 413 M warp instructions on ISETP / BRA / CS2R around 190 M FFMAs: profiles/round2_s3_ncu_misc_kernels_summary.txt)
template <int EPT>
__global__ void __launch_bounds__(256) patch_embed4_kernel(const float* __restrict__ img, const float4* __restrict__ wq,
                                                           const float* __restrict__ bias, const float* __restrict__ g,
                                                           const float* __restrict__ be, float* __restrict__ out,
                                                           int Cin, int S, int E, void* __restrict__ x16, int fp16,
                                                           unsigned long long* __restrict__ stats) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float4 sm4[];
  const int G = S / 4, KG = Cin * 4;            // KG (channel, dy) groups of four dx taps
  float4* slab = sm4;                           // [KG][G]   one patch row of input pixels, a float4 per patch
  float4* wt = slab + KG * G;                   // [KG][E]
  const int groups = (G + kPeRows - 1) / kPeRows;
  const int b = blockIdx.x / groups, py0 = (blockIdx.x % groups) * kPeRows;
  for (int i = threadIdx.x; i < KG * E; i += blockDim.x) wt[i] = wq[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int EP = EPT > 0 ? EPT : E / 32;        // channels per lane (E % 32 == 0, E <= 256)
  float gam[8], bet[8], bia[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int e = lane + 32 * j;
    gam[j] = (j < EP) ? g[e] : 0.f; bet[j] = (j < EP) ? be[e] : 0.f; bia[j] = (j < EP) ? bias[e] : 0.f;
  }
  for (int pr = 0; pr < kPeRows && py0 + pr < G; ++pr) {
    const int py = py0 + pr;
    __syncthreads();
    for (int i = threadIdx.x; i < KG * G; i += blockDim.x) {
      const int r = i / G, px = i % G, c = r >> 2, dy = r & 3;
      slab[i] = *reinterpret_cast<const float4*>(&img[(((long)b * Cin + c) * S + (py * 4 + dy)) * S + px * 4]);
    }
    __syncthreads();
    for (int px0 = warp * kPe4Patches; px0 < G; px0 += nw * kPe4Patches) {
      float acc[kPe4Patches][8];
#pragma unroll
      for (int q = 0; q < kPe4Patches; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[q][j] = bia[j];
      for (int r = 0; r < KG; ++r) {
        float4 v[kPe4Patches];
#pragma unroll
        for (int q = 0; q < kPe4Patches; ++q) v[q] = (px0 + q < G) ? slab[r * G + px0 + q] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* wr = wt + r * E + lane;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < EP) {
            const float4 w4 = wr[32 * j];
#pragma unroll
            for (int q = 0; q < kPe4Patches; ++q) {
              // same tap order as the generic kernel: k = (c, dy, dx) ascending
              acc[q][j] = fmaf(v[q].x, w4.x, acc[q][j]);
              acc[q][j] = fmaf(v[q].y, w4.y, acc[q][j]);
              acc[q][j] = fmaf(v[q].z, w4.z, acc[q][j]);
              acc[q][j] = fmaf(v[q].w, w4.w, acc[q][j]);
            }
          }
      }
#pragma unroll
      for (int q = 0; q < kPe4Patches; ++q) {
        const int px = px0 + q;
        if (px >= G) break;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < EP) s += acc[q][j];
        const float mean = warp_sum(s) / (float)E;
        float qv = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < EP) { const float dd = acc[q][j] - mean; qv += dd * dd; }
        const float rstd = 1.0f / sqrtf(warp_sum(qv) / (float)E + kLnEps);
        const long row = ((long)b * G + py) * G + px;
        float* o = out + row * E;
        float ys = 0.f, yq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < EP) {
            const float y = (acc[q][j] - mean) * rstd * gam[j] + bet[j];
            o[lane + 32 * j] = y;
            if (x16) {
              ys += y; yq = fmaf(y, y, yq);
              if (fp16) reinterpret_cast<__half*>(x16)[row * E + lane + 32 * j] = __float2half_rn(y);
              else reinterpret_cast<__nv_bfloat16*>(x16)[row * E + lane + 32 * j] = __float2bfloat16_rn(y);
            }
          }
        if (x16) {                                  // integer adds: the statistics do not depend on the reduction order
          long long s_sum = __float2ll_rn(ys * 16777216.0f), s_sq = __float2ll_rn(yq * 65536.0f);
#pragma unroll
          for (int o2 = 16; o2 > 0; o2 >>= 1) {
            s_sum += __shfl_xor_sync(0xffffffffu, s_sum, o2);
            s_sq += __shfl_xor_sync(0xffffffffu, s_sq, o2);
          }
          if (lane == 0) { stats[2 * row] = (unsigned long long)s_sum; stats[2 * row + 1] = (unsigned long long)s_sq; }
        }
      }
    }
  }
}

// (E, Cin, 4, 4) conv filter -> wq[(c*4+dy)*E + e] = (w[e][c][dy][0..3])
__global__ void patch_filter_pack4_kernel(const float* __restrict__ w, float4* __restrict__ wq, int Cin, int E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cin * 4 * E) return;
  const int r = i / E, e = i % E;
  wq[i] = *reinterpret_cast<const float4*>(w + ((long)e * Cin * 4 + r) * 4);
}
cudaError_t launch_patch_filter_pack4(const float* w, float* wq, int Cin, int E, cudaStream_t st) {
  patch_filter_pack4_kernel<<<(Cin * 4 * E + 255) / 256, 256, 0, st>>>(w, reinterpret_cast<float4*>(wq), Cin, E);
  return cudaGetLastError();
}

cudaError_t launch_patch_embed4(const float* img, const float* wq, const float* b, const float* gamma, const float* beta,
                                float* out, int B, int Cin, int S, int E, cudaStream_t st, void* x16, int fp16, float* stats) {
  if (E % 32 || E > 256 || S % 16) return cudaErrorInvalidValue;
  const int G = S / 4;
  const size_t smem = ((size_t)Cin * 4 * G + (size_t)Cin * 4 * E) * sizeof(float4);
  const int groups = (G + kPeRows - 1) / kPeRows;
  if (E == 192) {
    static DynSmemState smem_state6;
    if (cudaError_t e = ensure_dyn_smem(patch_embed4_kernel<6>, smem, smem_state6)) return e;
    return launch_k(patch_embed4_kernel<6>, dim3(B * groups), dim3(256), smem, st, img, reinterpret_cast<const float4*>(wq), b, gamma, beta,
                    out, Cin, S, E, x16, fp16, reinterpret_cast<unsigned long long*>(stats));
  }
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(patch_embed4_kernel<0>, smem, smem_state)) return e;
  return launch_k(patch_embed4_kernel<0>, dim3(B * groups), dim3(256), smem, st, img, reinterpret_cast<const float4*>(wq), b, gamma, beta,
                  out, Cin, S, E, x16, fp16, reinterpret_cast<unsigned long long*>(stats));
}

// ------------------------------------------------------------------------------------------
// Patch width 4 on the tensor cores (16-bit modes): the convolution is a (patches x 16 Cin) . (16 Cin x E) contraction;
// per warp one m16 tile of patches against all E channels with mma.sync.m16n8k8 TF32 (inputs rounded to the 10-bit
// TF32 significand -- the same 2^-11 relative rounding every other operand of the 16-bit modes gets; the fp32 parity mode
// keeps the exact CUDA-core kernels above).  The CUDA-core kernel is bound by shared-memory wavefronts (61.6 M per
// launch for 190 M FFMAs: profiles/round2_s3_ncu_misc_kernels_summary.txt); here a k8 step is 4 + 2 NT conflict-free LDS.32
// for NT tensor-core instructions, and the kernel is bound by its HBM traffic.
// One CTA = one patch row at a time (G / 16 warps, one tile each), kPeRows rows per CTA.  Fragment layouts (PTX ISA,
// m16n8k8 .tf32): g = lane >> 2, t = lane & 3; A: a0 (g, t) a1 (g + 8, t) a2 (g, t + 4) a3 (g + 8, t + 4);
// B: b0 (k = t, n = g) b1 (k = t + 4, n = g); C: c0 (g, 2t) c1 (g, 2t + 1) c2 (g + 8, 2t) c3 (g + 8, 2t + 1).
// Taps of k8 step ks: (c, dy) rows 2 ks and 2 ks + 1 of the slab, dx = 0..3 -- the conv weight's own (c, dy, dx) order.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
constexpr int kPeTcWPitch = 8;        // weight row pitch = E + 8 words: the 4 k rows of a B fragment hit different banks

template <int NT>                     // NT = E / 8 channel tiles
__global__ void __launch_bounds__(256) patch_embed4_tf32_kernel(const float* __restrict__ img, const float4* __restrict__ wq,
                                                                const float* __restrict__ bias, const float* __restrict__ g,
                                                                const float* __restrict__ be, float* __restrict__ out,
                                                                int Cin, int S, void* __restrict__ x16, int fp16,
                                                                unsigned long long* __restrict__ stats) {
  pdl_wait();
  pdl_trigger();
  constexpr int E = NT * 8, WP = E + kPeTcWPitch;
  extern __shared__ float4 sm4[];
  const int G = S / 4, KG = Cin * 4;
  float* slab0 = reinterpret_cast<float*>(sm4);                // [2][KG][G][4]   two patch rows of pixels (double buffer)
  uint32_t* wt = reinterpret_cast<uint32_t*>(slab0 + 2 * KG * G * 4);   // [KG * 4][WP]  tf32 filter bank, k-major
  float* prm = reinterpret_cast<float*>(wt + KG * 4 * WP);     // bias | gamma | beta  [3][E]
  const int groups = (G + kPeRows - 1) / kPeRows;
  const int b = blockIdx.x / groups, py0 = (blockIdx.x % groups) * kPeRows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int i = tid; i < KG * E; i += blockDim.x) {
    const int r = i / E, e = i - r * E;
    const float4 v = wq[i];
    wt[(r * 4 + 0) * WP + e] = to_tf32(v.x); wt[(r * 4 + 1) * WP + e] = to_tf32(v.y);
    wt[(r * 4 + 2) * WP + e] = to_tf32(v.z); wt[(r * 4 + 3) * WP + e] = to_tf32(v.w);
  }
  for (int i = tid; i < E; i += blockDim.x) { prm[i] = bias[i]; prm[E + i] = g[i]; prm[2 * E + i] = be[i]; }
  const int gq = lane >> 2, t = lane & 3;
  const int tiles = G / 16;
  const uint32_t slab_u = (uint32_t)__cvta_generic_to_shared(slab0);
  // the pixels of patch row py0 + pr stream into buffer pr & 1 (cp.async) while the previous row is on the tensor cores
  auto load_row = [&](int pr) {
    const int py = py0 + pr;
    if (pr < kPeRows && py < G) {
      const uint32_t dst = slab_u + (uint32_t)((pr & 1) * KG * G * 16);
      for (int i = tid; i < KG * G; i += blockDim.x) {
        const int r = i / G, px = i - r * G, c = r >> 2, dy = r & 3;
        cp_async16(dst + i * 16, &img[(((long)b * Cin + c) * S + (py * 4 + dy)) * S + px * 4]);
      }
    }
    cp_async_commit();
  };
  load_row(0);
  for (int pr = 0; pr < kPeRows && py0 + pr < G; ++pr) {
    const int py = py0 + pr;
    load_row(pr + 1);
    cp_async_wait<1>();
    __syncthreads();
    const float* slab = slab0 + (pr & 1) * KG * G * 4;
    for (int tile = warp; tile < tiles; tile += nw) {
      const int p0 = tile * 16;
      float acc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float2 bb = *reinterpret_cast<const float2*>(prm + 8 * j + 2 * t);
        acc[j][0] = bb.x; acc[j][1] = bb.y; acc[j][2] = bb.x; acc[j][3] = bb.y;
      }
      for (int ks = 0; ks < KG / 2; ++ks) {
        const float* s0 = slab + ((2 * ks) * G + p0 + gq) * 4 + t;
        const float* s1 = s0 + G * 4;
        const uint32_t a0 = to_tf32(s0[0]), a1 = to_tf32(s0[32]), a2 = to_tf32(s1[0]), a3 = to_tf32(s1[32]);
        const uint32_t* wr = wt + (8 * ks + t) * WP + gq;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const uint32_t b0 = wr[8 * j], b1 = wr[4 * WP + 8 * j];
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
                       : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
      }
      // LayerNorm over the E channels of rows gq and gq + 8: quad reductions
      float s_lo = 0.f, s_hi = 0.f;
#pragma unroll
      for (int j = 0; j < NT; ++j) { s_lo += acc[j][0] + acc[j][1]; s_hi += acc[j][2] + acc[j][3]; }
      s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 1); s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 2);
      s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 1); s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 2);
      const float m_lo = s_lo / (float)E, m_hi = s_hi / (float)E;
      float q_lo = 0.f, q_hi = 0.f;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float d0 = acc[j][0] - m_lo, d1 = acc[j][1] - m_lo, d2 = acc[j][2] - m_hi, d3 = acc[j][3] - m_hi;
        q_lo += d0 * d0 + d1 * d1; q_hi += d2 * d2 + d3 * d3;
      }
      q_lo += __shfl_xor_sync(0xffffffffu, q_lo, 1); q_lo += __shfl_xor_sync(0xffffffffu, q_lo, 2);
      q_hi += __shfl_xor_sync(0xffffffffu, q_hi, 1); q_hi += __shfl_xor_sync(0xffffffffu, q_hi, 2);
      const float r_lo = 1.0f / sqrtf(q_lo / (float)E + kLnEps), r_hi = 1.0f / sqrtf(q_hi / (float)E + kLnEps);
      const long row_lo = ((long)b * G + py) * G + p0 + gq, row_hi = row_lo + 8;
      float* o_lo = out + row_lo * E + 2 * t;
      float* o_hi = out + row_hi * E + 2 * t;
      float ys_lo = 0.f, yq_lo = 0.f, ys_hi = 0.f, yq_hi = 0.f;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float2 ga = *reinterpret_cast<const float2*>(prm + E + 8 * j + 2 * t), bt = *reinterpret_cast<const float2*>(prm + 2 * E + 8 * j + 2 * t);
        const float y0 = (acc[j][0] - m_lo) * r_lo * ga.x + bt.x, y1 = (acc[j][1] - m_lo) * r_lo * ga.y + bt.y;
        const float y2 = (acc[j][2] - m_hi) * r_hi * ga.x + bt.x, y3 = (acc[j][3] - m_hi) * r_hi * ga.y + bt.y;
        *reinterpret_cast<float2*>(o_lo + 8 * j) = make_float2(y0, y1);
        *reinterpret_cast<float2*>(o_hi + 8 * j) = make_float2(y2, y3);
        if (x16) {
          ys_lo += y0 + y1; yq_lo = fmaf(y0, y0, fmaf(y1, y1, yq_lo));
          ys_hi += y2 + y3; yq_hi = fmaf(y2, y2, fmaf(y3, y3, yq_hi));
          if (fp16) {
            *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(x16) + row_lo * E + 8 * j + 2 * t) = __floats2half2_rn(y0, y1);
            *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(x16) + row_hi * E + 8 * j + 2 * t) = __floats2half2_rn(y2, y3);
          } else {
            *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(x16) + row_lo * E + 8 * j + 2 * t) = __floats2bfloat162_rn(y0, y1);
            *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(x16) + row_hi * E + 8 * j + 2 * t) = __floats2bfloat162_rn(y2, y3);
          }
        }
      }
      if (x16) {                                  // 64-bit fixed-point row statistics (integer adds: order-independent)
        long long a_lo = __float2ll_rn(ys_lo * 16777216.0f), b_lo = __float2ll_rn(yq_lo * 65536.0f);
        long long a_hi = __float2ll_rn(ys_hi * 16777216.0f), b_hi = __float2ll_rn(yq_hi * 65536.0f);
#pragma unroll
        for (int o2 = 1; o2 <= 2; o2 <<= 1) {
          a_lo += __shfl_xor_sync(0xffffffffu, a_lo, o2); b_lo += __shfl_xor_sync(0xffffffffu, b_lo, o2);
          a_hi += __shfl_xor_sync(0xffffffffu, a_hi, o2); b_hi += __shfl_xor_sync(0xffffffffu, b_hi, o2);
        }
        if (t == 0) {
          stats[2 * row_lo] = (unsigned long long)a_lo; stats[2 * row_lo + 1] = (unsigned long long)b_lo;
          stats[2 * row_hi] = (unsigned long long)a_hi; stats[2 * row_hi + 1] = (unsigned long long)b_hi;
        }
      }
    }
    __syncthreads();                              // the buffer read here is refilled by the next iteration's prefetch
  }
}

bool patch_embed4_tc_supported(int Cin, int S, int E) { return E == 192 && S % 64 == 0 && Cin >= 1 && Cin <= 4; }

cudaError_t launch_patch_embed4_tc(const float* img, const float* wq, const float* b, const float* gamma, const float* beta,
                                   float* out, int B, int Cin, int S, int E, cudaStream_t st, void* x16, int fp16, float* stats) {
  if (!patch_embed4_tc_supported(Cin, S, E)) return cudaErrorInvalidValue;
  const int G = S / 4;
  const size_t smem = ((size_t)2 * Cin * 4 * G * 4 + (size_t)Cin * 16 * (E + kPeTcWPitch) + 3 * (size_t)E) * sizeof(float);
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(patch_embed4_tf32_kernel<24>, smem, smem_state)) return e;
  const int groups = (G + kPeRows - 1) / kPeRows;
  const int threads = 32 * (G / 16 < 8 ? G / 16 : 8);
  return launch_k(patch_embed4_tf32_kernel<24>, dim3(B * groups), dim3(threads), smem, st, img, reinterpret_cast<const float4*>(wq), b, gamma,
                  beta, out, Cin, S, x16, fp16, reinterpret_cast<unsigned long long*>(stats));
}

cudaError_t launch_patch_embed(const float* img, const float* w, const float* b, const float* gamma,
                               const float* beta, float* out, int B, int Cin, int S, int P, int E,
                               cudaStream_t st) {
  if (E % 32 || E > 256 || S % P || S % 4) return cudaErrorInvalidValue;
  const size_t smem = ((size_t)Cin * P * S + (size_t)Cin * P * P * E + (size_t)Cin * P * P) * sizeof(float);
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(patch_embed_kernel, smem, smem_state)) return e;
  const int G = S / P, groups = (G + kPeRows - 1) / kPeRows;
  launch_k(patch_embed_kernel, dim3(B * groups), dim3(256), smem, st, img, w, b, gamma, beta, out, Cin, S, P, E);
  return cudaGetLastError();
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ x, T* __restrict__ y, long n) {
  pdl_wait();
  pdl_trigger();
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    store4<T>(y + i, v.x, v.y, v.z, v.w);
  } else {
    for (long j = i; j < n; ++j) y[j] = from_f32<T>(x[j]);
  }
}
template <typename T>
cudaError_t launch_cast(const float* x, T* y, long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  launch_k(cast_kernel<T>, dim3((unsigned)((n / 4 + 256) / 256)), dim3(256), 0, st, x, y, n);
  return cudaGetLastError();
}
template cudaError_t launch_cast<bf16>(const float*, bf16*, long, cudaStream_t);
template cudaError_t launch_cast<f16>(const float*, f16*, long, cudaStream_t);
template cudaError_t launch_cast<float>(const float*, float*, long, cudaStream_t);

// ------------------------------------------------------------------------------------------
// Static expansion weights (reference models/layers.py:55-92).
//   a_fw[b,e,n] = relu(z)[b,e,n]*[n < n_valid_b] / (sum_n ... + eps)          (row normalise)
//   a_bw[b,n,e] = relu(z)[b,e,n] / (sum_{e' in group(e)} relu(z)[b,e',n] + eps) (per-group)
// and the same for relu(-z).  Two launches: per-group column sums, then the scaled writes.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) se_group_sum_kernel(const float* __restrict__ z, const int* __restrict__ gstart,
                                                           float* __restrict__ gsum, int E, int N, int n_groups) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x, g = blockIdx.y;
  const int e0 = gstart[g], e1 = gstart[g + 1];
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float sa = 0.f, sb = 0.f;
    const float* zp = z + ((long)b * E + e0) * N + n;
    for (int e = e0; e < e1; ++e, zp += N) {
      const float v = *zp;
      sa += fmaxf(v, 0.f);
      sb += fmaxf(-v, 0.f);
    }
    gsum[(((long)b * n_groups + g) * 2 + 0) * N + n] = sa;
    gsum[(((long)b * n_groups + g) * 2 + 1) * N + n] = sb;
  }
}

template <typename WT>
__global__ void __launch_bounds__(256) se_weights_kernel(const float* __restrict__ z, const int* __restrict__ n_valid,
                                                         const int* __restrict__ gstart, int n_groups,
                                                         const float* __restrict__ gsum, WT* __restrict__ a_fw,
                                                         WT* __restrict__ b_fw, WT* __restrict__ a_bw,
                                                         WT* __restrict__ b_bw, int E, int N, int chunk) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float zt[];            // [chunk][N]
  const int b = blockIdx.x, e0 = blockIdx.y * chunk;
  const int nv = n_valid ? n_valid[b] : N;
  int g = 0;
  while (g + 1 < n_groups && gstart[g + 1] <= e0) ++g;
  const float* zp = z + ((long)b * E + e0) * N;
  for (int i = threadIdx.x; i < chunk * N; i += blockDim.x) zt[i] = zp[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int e = warp; e < chunk; e += nw) {
    float sa = 0.f, sb = 0.f;
    for (int n = lane; n < nv; n += 32) {
      const float v = zt[e * N + n];
      sa += fmaxf(v, 0.f);
      sb += fmaxf(-v, 0.f);
    }
    sa = warp_sum(sa) + kExpEps;
    sb = warp_sum(sb) + kExpEps;
    WT* ao = a_fw + ((long)b * E + e0 + e) * N;
    WT* bo = b_fw + ((long)b * E + e0 + e) * N;
    for (int n = lane; n < N; n += 32) {
      const float v = zt[e * N + n];
      const bool ok = n < nv;
      ao[n] = from_f32<WT>(ok ? fmaxf(v, 0.f) / sa : 0.f);
      bo[n] = from_f32<WT>(ok ? fmaxf(-v, 0.f) / sb : 0.f);
    }
  }
  const float* ga = gsum + (((long)b * n_groups + g) * 2 + 0) * N;
  const float* gb = ga + N;
  for (int i = threadIdx.x; i < chunk * N; i += blockDim.x) {
    const int e = i % chunk, n = i / chunk;
    const float v = zt[e * N + n];
    const long o = ((long)b * N + n) * E + e0 + e;
    a_bw[o] = from_f32<WT>(fmaxf(v, 0.f) / (ga[n] + kExpEps));
    b_bw[o] = from_f32<WT>(fmaxf(-v, 0.f) / (gb[n] + kExpEps));
  }
}

template <typename WT>
cudaError_t launch_static_exp_weights(const float* z, const int* n_valid, const int* group_start, int n_groups,
                                      WT* a_fw, WT* b_fw, WT* a_bw, WT* b_bw, float* gsum_scratch,
                                      int B, int E, int N, int chunk, cudaStream_t st) {
  // `chunk` rows of z per CTA; it must divide every group boundary (the engine picks the gcd, <= 32)
  if (chunk <= 0 || (E % chunk)) return cudaErrorInvalidValue;
  launch_k(se_group_sum_kernel, dim3(dim3(B, n_groups)), dim3(160), 0, st, z, group_start, gsum_scratch, E, N, n_groups);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  launch_k(se_weights_kernel<WT>, dim3(dim3(B, E / chunk)), dim3(256), (size_t)chunk * N * sizeof(float), st, 
      z, n_valid, group_start, n_groups, gsum_scratch, a_fw, b_fw, a_bw, b_bw, E, N, chunk);
  return cudaGetLastError();
}
template cudaError_t launch_static_exp_weights<float>(const float*, const int*, const int*, int, float*, float*, float*, float*, float*, int, int, int, int, cudaStream_t);
template cudaError_t launch_static_exp_weights<bf16>(const float*, const int*, const int*, int, bf16*, bf16*, bf16*, bf16*, float*, int, int, int, int, cudaStream_t);
template cudaError_t launch_static_exp_weights<f16>(const float*, const int*, const int*, int, f16*, f16*, f16*, f16*, float*, int, int, int, int, cudaStream_t);

// x_out = x_in + sigmoid(sel)*a + (1-sigmoid(sel))*b     (reference layers.py:98-102,118-120)
template <typename ST>
__global__ void selector_mix_kernel(const float* __restrict__ xi, long ldxi, const ST* __restrict__ sel, long lds,
                                    const float* __restrict__ a, const float* __restrict__ bb, long ldo,
                                    float* __restrict__ xo, long ldxo, long rows, int d) {
  pdl_wait();
  pdl_trigger();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * d) return;
  const long r = i / d;
  const int c = (int)(i % d);
  const float s = sigmoidf_(to_f32<ST>(sel[r * lds + c]));
  xo[r * ldxo + c] = xi[r * ldxi + c] + (s * a[r * ldo + c] + (1.0f - s) * bb[r * ldo + c]);
}
template <typename ST>
cudaError_t launch_selector_mix(const float* x_in, long ldxi, const ST* sel, long lds, const float* out_a,
                                const float* out_b, long ldo, float* x_out, long ldxo, long rows, int d,
                                cudaStream_t st) {
  const long n = rows * d;
  launch_k(selector_mix_kernel<ST>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, x_in, ldxi, sel, lds, out_a, out_b, ldo, x_out,
                                                                      ldxo, rows, d);
  return cudaGetLastError();
}
template cudaError_t launch_selector_mix<float>(const float*, long, const float*, long, const float*, const float*, long, float*, long, long, int, cudaStream_t);
template cudaError_t launch_selector_mix<bf16>(const float*, long, const bf16*, long, const float*, const float*, long, float*, long, long, int, cudaStream_t);
template cudaError_t launch_selector_mix<f16>(const float*, long, const f16*, long, const float*, const float*, long, float*, long, long, int, cudaStream_t);

// ---- LayerNorm folded into the consuming GEMM (kernels.h: TcGemmArgs::ln_stats): one warp per output row n
//   w16[n][k] = round16(gamma[k] W[n][k] - (1/K) sum_k' gamma[k'] W[n][k'])     (centred: the row mean of x drops out)
//   bias_out[n] = bias[n] + sum_k W[n][k] beta[k]
template <typename T>
__global__ void fold_ln_weight_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ bias, T* __restrict__ w16, float* __restrict__ bias_out, int N, int K) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float s = 0.f, bb = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = w[(long)n * K + k];
    s = fmaf(gamma[k], wv, s);
    bb = fmaf(wv, beta[k], bb);
  }
  s = warp_sum(s) / (float)K;
  bb = warp_sum(bb);
  for (int k = lane; k < K; k += 32) w16[(long)n * K + k] = from_f32<T>(gamma[k] * w[(long)n * K + k] - s);
  if (lane == 0) bias_out[n] = (bias ? bias[n] : 0.f) + bb;
}
cudaError_t launch_fold_ln_weight(const float* w, const float* gamma, const float* beta, const float* bias, void* w16,
                                  float* bias_out, int N, int K, int fp16, cudaStream_t st) {
  const int wpb = 8;
  if (fp16) fold_ln_weight_kernel<f16><<<(N + wpb - 1) / wpb, wpb * 32, 0, st>>>(w, gamma, beta, bias, reinterpret_cast<f16*>(w16), bias_out, N, K);
  else fold_ln_weight_kernel<bf16><<<(N + wpb - 1) / wpb, wpb * 32, 0, st>>>(w, gamma, beta, bias, reinterpret_cast<bf16*>(w16), bias_out, N, K);
  return cudaGetLastError();
}

}  // namespace xn
