// Static expansion (reference models/layers.py:20-102) around the tcgen05 contractions of the 16-bit modes.
//
// The scores come out of the linear-layer GEMM TRANSPOSED, zT[b][n][e] = key[b][n] . q_e / sqrt(d) (one plain
// X W^T launch over all images: M = B * N tokens, N = E expansion vectors), so the kernels here work on 16-token
// slabs of an image (16 x E fp32 = 62 KB in shared memory):
//
//   se_sums_t_kernel      per (image, slab): the per-group sums over e of relu(+-z) for its 16 tokens (the backward
//                         normalisers, layers.py:66-80) and the slab's partial sums over n of relu(+-z) per e (the
//                         forward normalisers, layers.py:55-58, tokens >= n_valid masked); partials are written per
//                         slab and added in slab order by the consumer -- no atomics, so results do not depend on
//                         scheduling
//   se_weights_t_kernel   per (image, slab): a_bw / b_bw [b][n][e] (coalesced along e) and a_fw / b_fw [b][e][n]
//                         (one 32-byte sector per (e, slab)), rounded to the operand type
//   transpose_ab_kernel   class_a / class_b projections [b n][c] -> [b][c][n]: the K-major left operand of
//                         class^T = A^T . fw^T
//   selector_mix_t_kernel x + sigmoid(sel) a + (1 - sigmoid(sel)) b with a, b given transposed ([b][c][n], the output
//                         of out^T = class^T . bw^T)
//
// Requirements (checked by static_exp_t_supported; the engine falls back to the (B,E,N)-layout kernels of
// elementwise.cu otherwise): N % 16 == 0, E % 32 == 0, E <= 1024, every group boundary a multiple of 32, <= 8 groups.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "kernels.h"
#include "common.cuh"

namespace xn {

namespace {
constexpr float kSeEps = 1e-9f;          // reference models/layers.py:42 (eps of the expansion normalisers)
constexpr int kSlab = 16;                // tokens per CTA
constexpr int kMaxJ = 32;                // E / 32 <= 32
constexpr int kMaxG = 8;

template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
}  // namespace

// gsum[((b * G + g) * 2 + s) * N + n]        s = 0: relu(z), 1: relu(-z)     (layout of se_group_sum_kernel)
// colpart[((b * nslab + slab) * 2 + s) * E + e]
__global__ void __launch_bounds__(256) se_sums_t_kernel(const float* __restrict__ zT, const int* __restrict__ n_valid,
                                                        const int* __restrict__ gstart, int G, float* __restrict__ gsum,
                                                        float* __restrict__ colpart, int E, int N) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[2][kMaxJ * 32];
  __shared__ int gs_s[kMaxG + 1];
  const int b = blockIdx.x, slab = blockIdx.y, n0 = slab * kSlab;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int J = E >> 5;
  if (tid <= G) gs_s[tid] = gstart[tid];
  for (int i = tid; i < 2 * kMaxJ * 32; i += 256) (&red[0][0])[i] = 0.f;
  __syncthreads();
  const int nv = n_valid ? n_valid[b] : N;
  float ca[kMaxJ], cb[kMaxJ];
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j) ca[j] = cb[j] = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int n = n0 + 2 * warp + t;
    const bool valid = n < nv;
    const float* zr = zT + ((long)b * N + n) * E + lane;
    float v[kMaxJ];
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j) v[j] = j < J ? zr[32 * j] : 0.f;      // all loads of the row in flight
    int g = 0;
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j) {
      if (j < J) {
        if (32 * j == gs_s[g + 1]) {                                          // warp-uniform
          sa = warp_sum(sa); sb = warp_sum(sb);
          if (lane == 0) {
            gsum[(((long)b * G + g) * 2 + 0) * N + n] = sa;
            gsum[(((long)b * G + g) * 2 + 1) * N + n] = sb;
          }
          ++g; sa = sb = 0.f;
        }
        const float p = fmaxf(v[j], 0.f), q = fmaxf(-v[j], 0.f);
        sa += p; sb += q;
        if (valid) { ca[j] += p; cb[j] += q; }
      }
    }
    sa = warp_sum(sa); sb = warp_sum(sb);
    if (lane == 0) {
      gsum[(((long)b * G + g) * 2 + 0) * N + n] = sa;
      gsum[(((long)b * G + g) * 2 + 1) * N + n] = sb;
    }
  }
  // column partials of the slab: warps add in warp order (fixed summation order)
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < kMaxJ; ++j)
        if (j < J) { red[0][32 * j + lane] += ca[j]; red[1][32 * j + lane] += cb[j]; }
    }
    __syncthreads();
  }
  const int nslab = gridDim.y;
  float* cp = colpart + ((long)b * nslab + slab) * 2 * E;
  for (int e = tid; e < E; e += 256) { cp[e] = red[0][e]; cp[E + e] = red[1][e]; }
}

template <typename T>
__global__ void __launch_bounds__(256) se_weights_t_kernel(const float* __restrict__ zT, const int* __restrict__ n_valid,
                                                           const int* __restrict__ gstart, int G,
                                                           const float* __restrict__ gsum, const float* __restrict__ colpart,
                                                           T* __restrict__ a_fw, T* __restrict__ b_fw, T* __restrict__ a_bw,
                                                           T* __restrict__ b_bw, int E, int N) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float se_sm[];
  float* zt = se_sm;                          // [16][E]
  float* inv = zt + kSlab * E;                // [2][E]   1 / (sum_n + eps)
  float* ginv = inv + 2 * E;                  // [2][G][16]   1 / (group sum + eps)
  __shared__ int gs_s[kMaxG + 1];
  const int b = blockIdx.x, slab = blockIdx.y, n0 = slab * kSlab, nslab = gridDim.y;
  const int tid = threadIdx.x;
  if (tid <= G) gs_s[tid] = gstart[tid];
  {
    const float4* src = reinterpret_cast<const float4*>(zT + ((long)b * N + n0) * E);
    float4* dst = reinterpret_cast<float4*>(zt);
    for (int i = tid; i < kSlab * E / 4; i += 256) dst[i] = src[i];
  }
  for (int i = tid; i < 2 * E; i += 256) {
    const int s = i / E, e = i - s * E;
    float acc = 0.f;
    for (int sl = 0; sl < nslab; ++sl) acc += colpart[(((long)b * nslab + sl) * 2 + s) * E + e];
    inv[i] = 1.0f / (acc + kSeEps);
  }
  for (int i = tid; i < 2 * G * kSlab; i += 256) {
    const int s = i / (G * kSlab), r = i - s * G * kSlab, g = r / kSlab, k = r - g * kSlab;
    ginv[i] = 1.0f / (gsum[(((long)b * G + g) * 2 + s) * N + n0 + k] + kSeEps);
  }
  __syncthreads();
  const int nv = n_valid ? n_valid[b] : N;
  // backward weights [b][n][e]: eight consecutive e per thread (one group: boundaries are multiples of 32)
  const int e8n = E >> 3;
  for (int idx = tid; idx < kSlab * e8n; idx += 256) {
    const int k = idx / e8n, e0 = (idx - k * e8n) * 8;
    int g = 0;
    while (g + 1 < G && gs_s[g + 1] <= e0) ++g;
    const float ia = ginv[(0 * G + g) * kSlab + k], ib = ginv[(1 * G + g) * kSlab + k];
    const float4 v0 = *reinterpret_cast<const float4*>(zt + k * E + e0), v1 = *reinterpret_cast<const float4*>(zt + k * E + e0 + 4);
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t pa[4], pb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pa[i] = pack2<T>(fmaxf(v[2 * i], 0.f) * ia, fmaxf(v[2 * i + 1], 0.f) * ia);
      pb[i] = pack2<T>(fmaxf(-v[2 * i], 0.f) * ib, fmaxf(-v[2 * i + 1], 0.f) * ib);
    }
    const long o = ((long)b * N + n0 + k) * E + e0;
    *reinterpret_cast<uint4*>(a_bw + o) = make_uint4(pa[0], pa[1], pa[2], pa[3]);
    *reinterpret_cast<uint4*>(b_bw + o) = make_uint4(pb[0], pb[1], pb[2], pb[3]);
  }
  // forward weights [b][e][n]: the slab's 16 tokens of one e per thread (32 bytes); consecutive threads take
  // consecutive e, so the shared-memory reads (row pitch E = 0 mod 32 words) are conflict-free
  for (int e = tid; e < E; e += 256) {
    const float ia = inv[e], ib = inv[E + e];
    uint32_t pa[8], pb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x0 = (n0 + 2 * i < nv) ? zt[(2 * i) * E + e] : 0.f, x1 = (n0 + 2 * i + 1 < nv) ? zt[(2 * i + 1) * E + e] : 0.f;
      pa[i] = pack2<T>(fmaxf(x0, 0.f) * ia, fmaxf(x1, 0.f) * ia);
      pb[i] = pack2<T>(fmaxf(-x0, 0.f) * ib, fmaxf(-x1, 0.f) * ib);
    }
    const long o = ((long)b * E + e) * N + n0;
    uint4* da = reinterpret_cast<uint4*>(a_fw + o);
    uint4* db = reinterpret_cast<uint4*>(b_fw + o);
    da[0] = make_uint4(pa[0], pa[1], pa[2], pa[3]); da[1] = make_uint4(pa[4], pa[5], pa[6], pa[7]);
    db[0] = make_uint4(pb[0], pb[1], pb[2], pb[3]); db[1] = make_uint4(pb[4], pb[5], pb[6], pb[7]);
  }
}

bool static_exp_t_supported(const int* group_start_host, int n_groups, int E, int N) {
  if (N % kSlab || E % 32 || E > 32 * kMaxJ || n_groups < 1 || n_groups > kMaxG || (N & 7)) return false;
  for (int g = 0; g <= n_groups; ++g)
    if (group_start_host[g] % 32) return false;
  return group_start_host[0] == 0 && group_start_host[n_groups] == E;
}

template <typename T>
cudaError_t launch_static_exp_weights_t(const float* zT, const int* n_valid, const int* group_start, int n_groups, T* a_fw,
                                        T* b_fw, T* a_bw, T* b_bw, float* gsum_scratch, float* colpart_scratch, int B, int E,
                                        int N, cudaStream_t st) {
  const dim3 grid(B, N / kSlab);
  launch_k(se_sums_t_kernel, grid, dim3(256), 0, st, zT, n_valid, group_start, n_groups, gsum_scratch, colpart_scratch, E, N);
  if (cudaError_t e = cudaGetLastError()) return e;
  const size_t smem = ((size_t)(kSlab + 2) * E + 2 * n_groups * kSlab) * sizeof(float);
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(se_weights_t_kernel<T>, smem, smem_state)) return e;
  launch_k(se_weights_t_kernel<T>, grid, dim3(256), smem, st, zT, n_valid, group_start, n_groups, (const float*)gsum_scratch,
           (const float*)colpart_scratch, a_fw, b_fw, a_bw, b_bw, E, N);
  return cudaGetLastError();
}
template cudaError_t launch_static_exp_weights_t<bf16>(const float*, const int*, const int*, int, bf16*, bf16*, bf16*, bf16*, float*, float*, int, int, int, cudaStream_t);
template cudaError_t launch_static_exp_weights_t<f16>(const float*, const int*, const int*, int, f16*, f16*, f16*, f16*, float*, float*, int, int, int, cudaStream_t);

// src [b * N + n][ld] columns c0 .. c0 + C (16-bit)  ->  dst [b][C][N]
template <typename T>
__global__ void __launch_bounds__(256) transpose_ab_kernel(const T* __restrict__ src, long ld, int c0, T* __restrict__ dst, int C, int N) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) uint32_t tr_sm[];     // [N][33] words = 64 (+2) 16-bit columns per token
  const int b = blockIdx.x, cb = blockIdx.y * 64, tid = threadIdx.x;
  for (int idx = tid; idx < N * 8; idx += 256) {
    const int n = idx >> 3, ch = idx & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(src + ((long)b * N + n) * ld + c0 + cb + ch * 8);
    uint32_t* d = tr_sm + n * 33 + ch * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const uint16_t* t16 = reinterpret_cast<const uint16_t*>(tr_sm);
  const int n8n = N >> 3;
  for (int idx = tid; idx < 64 * n8n; idx += 256) {
    const int c = idx & 63, n8 = (idx >> 6) * 8;
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      w[i] = (uint32_t)t16[(n8 + 2 * i) * 66 + c] | ((uint32_t)t16[(n8 + 2 * i + 1) * 66 + c] << 16);
    *reinterpret_cast<uint4*>(dst + ((long)b * C + cb + c) * N + n8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
template <typename T>
cudaError_t launch_transpose_ab(const T* src, long ld, int c0, T* dst, int B, int C, int N, cudaStream_t st) {
  if (C % 64 || N % 8 || (ld & 7) || (c0 & 7)) return cudaErrorInvalidValue;
  const size_t smem = (size_t)N * 33 * 4;
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(transpose_ab_kernel<T>, smem, smem_state)) return e;
  launch_k(transpose_ab_kernel<T>, dim3(B, C / 64), dim3(256), smem, st, src, ld, c0, dst, C, N);
  return cudaGetLastError();
}
template cudaError_t launch_transpose_ab<bf16>(const bf16*, long, int, bf16*, int, int, int, cudaStream_t);
template cudaError_t launch_transpose_ab<f16>(const f16*, long, int, f16*, int, int, int, cudaStream_t);

// x_out[b n][c] = x_in[b n][c] + s a^T[b][c][n] + (1 - s) b^T[b][c][n],  s = sigmoid(sel[b n][c])   (layers.py:98-102,118-120)
template <typename ST>
__global__ void __launch_bounds__(256) selector_mix_t_kernel(const float* __restrict__ xi, long ldxi, const ST* __restrict__ sel, long lds,
                                                             const float* __restrict__ at, const float* __restrict__ bt,
                                                             float* __restrict__ xo, long ldxo, int N, int d) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float mx_sm[];          // [2][32][N + 1]
  const int b = blockIdx.x, c0 = blockIdx.y * 32, tid = threadIdx.x;
  const int P = N + 1;
  for (int idx = tid; idx < 32 * N; idx += 256) {
    const int c = idx / N, n = idx - c * N;
    const long o = ((long)b * d + c0 + c) * N + n;
    mx_sm[c * P + n] = at[o];
    mx_sm[32 * P + c * P + n] = bt[o];
  }
  __syncthreads();
  for (int idx = tid; idx < 32 * N; idx += 256) {
    const int c = idx & 31, n = idx >> 5;
    const long r = (long)b * N + n;
    const float s = sigmoidf_(to_f32<ST>(sel[r * lds + c0 + c]));
    xo[r * ldxo + c0 + c] = xi[r * ldxi + c0 + c] + (s * mx_sm[c * P + n] + (1.0f - s) * mx_sm[32 * P + c * P + n]);
  }
}
template <typename ST>
cudaError_t launch_selector_mix_t(const float* x_in, long ldxi, const ST* sel, long lds, const float* out_a_t, const float* out_b_t,
                                  float* x_out, long ldxo, int B, int N, int d, cudaStream_t st) {
  if (d % 32) return cudaErrorInvalidValue;
  const size_t smem = (size_t)2 * 32 * (N + 1) * sizeof(float);
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(selector_mix_t_kernel<ST>, smem, smem_state)) return e;
  launch_k(selector_mix_t_kernel<ST>, dim3(B, d / 32), dim3(256), smem, st, x_in, ldxi, sel, lds, out_a_t, out_b_t, x_out, ldxo, N, d);
  return cudaGetLastError();
}
template cudaError_t launch_selector_mix_t<bf16>(const float*, long, const bf16*, long, const float*, const float*, float*, long, int, int, int, cudaStream_t);
template cudaError_t launch_selector_mix_t<f16>(const float*, long, const f16*, long, const float*, const float*, float*, long, int, int, int, cudaStream_t);

// out[c][r] = in[r][c]   (load-time transposes of small fp32 tables: bias_exp -> [d][E])
__global__ void transpose_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)R * C) return;
  const int c = (int)(i / R), r = (int)(i - (long)c * R);
  out[i] = in[(long)r * C + c];
}
cudaError_t launch_transpose_f32(const float* in, float* out, int R, int C, cudaStream_t st) {
  const long n = (long)R * C;
  transpose_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, R, C);
  return cudaGetLastError();
}

}  // namespace xn
