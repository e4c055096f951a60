// Decoder-step kernels: one new position per row per launch, with cached expansion state.
//
// The reference re-decodes the whole prefix every step (models/captioning_model.py:295-304).
// By causality everything position i contributes -- cond c_i, key K_i, the class
// projections A_i/B_i, the normalised forward weights of its row-block and q_e.K_i -- never
// changes once computed, so each step only adds the row-block and the column of position p.
// Beam reordering never copies that state: `anc[r][i]` names the slot (row) that holds position
// i of row r's history, and the kernels gather through it.
#include <algorithm>
#include "kernels.h"
#include "common.cuh"
#include "decode_rows.cuh"

namespace xn {


// y = E[tok] * sqrt(d) + P[p]   (reference layers.py:16-17, End_ExpansionNet_v2.py:171-189)
__global__ void embed_kernel(const int64_t* __restrict__ t64, const int* __restrict__ t32, long tok_stride, int p,
                             const float* __restrict__ emb, const float* __restrict__ pos, float* __restrict__ x,
                             long ldx, int R, int d) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x;
  const long tok = t64 ? (long)t64[r * tok_stride + p] : (long)t32[r * tok_stride + p];
  const float sc = sqrtf((float)d);
  for (int c = threadIdx.x; c < d; c += blockDim.x)
    x[(long)r * ldx + c] = emb[tok * d + c] * sc + pos[(long)p * d + c];
}
cudaError_t launch_embed(const int64_t* tokens64, const int* tokens32, long tok_stride, int p, const float* emb,
                         const float* pos, float* x, long ldx, int R, int d, cudaStream_t st) {
  launch_k(embed_kernel, dim3(R), dim3(128), 0, st, tokens64, tokens32, tok_stride, p, emb, pos, x, ldx, R, d);
  return cudaGetLastError();
}


// embedding + LayerNorm of the first decoder layer in one pass (d == 4 * blockDim.x)
template <typename T>
__global__ void __launch_bounds__(128) embed_ln_kernel(const int64_t* __restrict__ t64, const int* __restrict__ t32, long tok_stride,
                                                       int p, const float* __restrict__ emb, const float* __restrict__ pos,
                                                       float* __restrict__ x, long ldx, const float* __restrict__ g,
                                                       const float* __restrict__ be, T* __restrict__ xn, long ldn, int d) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  const int r = blockIdx.x, c = threadIdx.x * 4;
  const long tok = t64 ? (long)t64[r * tok_stride + p] : (long)t32[r * tok_stride + p];
  const float sc = sqrtf((float)d);
  const float4 e4 = *reinterpret_cast<const float4*>(emb + tok * d + c);
  const float4 p4 = *reinterpret_cast<const float4*>(pos + (long)p * d + c);
  float v[4] = {e4.x * sc + p4.x, e4.y * sc + p4.y, e4.z * sc + p4.z, e4.w * sc + p4.w};
  *reinterpret_cast<float4*>(x + (long)r * ldx + c) = make_float4(v[0], v[1], v[2], v[3]);
  float mean, rstd;
  block_ln_stats<4>(v, d, red, mean, rstd);
  const float4 g4 = *reinterpret_cast<const float4*>(g + c), b4 = *reinterpret_cast<const float4*>(be + c);
  T* o = xn + (long)r * ldn + c;
  o[0] = from_f32<T>((v[0] - mean) * rstd * g4.x + b4.x);
  o[1] = from_f32<T>((v[1] - mean) * rstd * g4.y + b4.y);
  o[2] = from_f32<T>((v[2] - mean) * rstd * g4.z + b4.z);
  o[3] = from_f32<T>((v[3] - mean) * rstd * g4.w + b4.w);
}
template <typename T>
cudaError_t launch_embed_ln(const int64_t* tokens64, const int* tokens32, long tok_stride, int p, const float* emb,
                            const float* pos, float* x, long ldx, const float* gamma, const float* beta, T* xn, long ldn, int R,
                            int d, cudaStream_t st) {
  if (d != 512 || (ldx & 3)) return cudaErrorInvalidValue;
  launch_k(embed_ln_kernel<T>, dim3(R), dim3(128), 0, st, tokens64, tokens32, tok_stride, p, emb, pos, x, ldx, gamma, beta, xn, ldn, d);
  return cudaGetLastError();
}
template cudaError_t launch_embed_ln<float>(const int64_t*, const int*, long, int, const float*, const float*, float*, long, const float*, const float*, float*, long, int, int, cudaStream_t);
template cudaError_t launch_embed_ln<bf16>(const int64_t*, const int*, long, int, const float*, const float*, float*, long, const float*, const float*, bf16*, long, int, int, cudaStream_t);
template cudaError_t launch_embed_ln<f16>(const int64_t*, const int*, long, int, const float*, const float*, float*, long, const float*, const float*, f16*, long, int, int, cudaStream_t);

template <typename T>
__global__ void __launch_bounds__(256) dyn_exp_step_kernel(DecState s, int layer, int p, const float* __restrict__ qexp,
                                                           const float* __restrict__ bexp, int n_exp,
                                                           const int* __restrict__ row_len,
                                                           const float* __restrict__ x_in, long ldxi,
                                                           float* __restrict__ x_out, long ldxo, int d,
                                                           const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                                           T* __restrict__ ln_out, long ldn) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float sm[];
  dyn_exp_row<T>(s, layer, p, qexp, bexp, n_exp, row_len, x_in, ldxi, x_out, ldxo, d, ln_g, ln_b, ln_out, ldn, blockIdx.x, sm);
}


template <typename T>
cudaError_t launch_dyn_exp_step(const DecState& s, int layer, int p, const float* qexp, const float* bexp, int n_exp,
                                const int* row_len, const float* x_in, long ldxi, float* x_out, long ldxo, int d,
                                const float* ln_g, const float* ln_b, T* ln_out, long ldn, cudaStream_t st) {
  if (s.P > 128 || (d & 3) || n_exp > 64) return cudaErrorInvalidValue;
  if (ln_out && d != 512) return cudaErrorInvalidValue;
  const size_t P4 = (s.P + 3) & ~3, E4 = (n_exp + 3) & ~3;
  const size_t smem = (P4 * 7 + E4 * 3 + 4 * P4 * E4 + 32 + std::max<size_t>(2 * (size_t)s.P * s.P, 1536)) * sizeof(float);
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(dyn_exp_step_kernel<T>, smem, smem_state)) return e;
  launch_k(dyn_exp_step_kernel<T>, dim3(s.R), dim3(256), smem, st, s, layer, p, qexp, bexp, n_exp, row_len, x_in, ldxi, x_out, ldxo, d, ln_g, ln_b,
                                                 ln_out, ldn);
  return cudaGetLastError();
}
template cudaError_t launch_dyn_exp_step<float>(const DecState&, int, int, const float*, const float*, int, const int*, const float*, long, float*, long, int, const float*, const float*, float*, long, cudaStream_t);
template cudaError_t launch_dyn_exp_step<bf16>(const DecState&, int, int, const float*, const float*, int, const int*, const float*, long, float*, long, int, const float*, const float*, bf16*, long, cudaStream_t);
template cudaError_t launch_dyn_exp_step<f16>(const DecState&, int, int, const float*, const float*, int, const int*, const float*, long, float*, long, int, const float*, const float*, f16*, long, cudaStream_t);

// ------------------------------------------------------------------------------------------
// Cross attention for one query position per row (reference models/layers.py:266-295).
// K/V of the encoder output are projected once per image and shared by its beams; one CTA
// per (image, head) stages that head's K and V in shared memory and serves all rows of the image.
// ------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void load4(const T* p, float (&o)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&o)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void load4<bf16>(const bf16* p, float (&o)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  o[0] = __bfloat162float(a.x); o[1] = __bfloat162float(a.y); o[2] = __bfloat162float(b.x); o[3] = __bfloat162float(b.y);
}
template <> __device__ __forceinline__ void load4<f16>(const f16* p, float (&o)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __half2 a = *reinterpret_cast<const __half2*>(&u.x), b = *reinterpret_cast<const __half2*>(&u.y);
  o[0] = __half2float(a.x); o[1] = __half2float(a.y); o[2] = __half2float(b.x); o[3] = __half2float(b.y);
}

constexpr int kMaxRpi = 8;      // rows (beams) per image handled together

template <typename KvT, typename OutT>
__global__ void __launch_bounds__(256) cross_attn_step_kernel(const float* __restrict__ q, long ldq,
                                                              const KvT* __restrict__ kv, long ldkv, int k_off,
                                                              int v_off, OutT* __restrict__ out, long ldo,
                                                              int rows_per_image, int n, int dk,
                                                              const int* __restrict__ n_valid,
                                                              const int* __restrict__ row_len, int p) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kt = n + 1;           // K is staged transposed [dk][n+1]: the score loop reads it conflict-free
  float* Kt = sm;                 // [dk][n+1]
  float* Vs = Kt + dk * kt;       // [n][dk]
  float* qs = Vs + n * dk;        // [rpi][dk]
  float* pr = qs + kMaxRpi * dk;  // [rpi][n]   scores, then probabilities
  float* po = pr + kMaxRpi * n;   // [parts][rpi][dk] partial outputs
  const int rpi = rows_per_image;
  const int dk4 = dk >> 2;
  for (int i = tid; i < rpi * dk; i += blockDim.x) {
    const int r = b * rpi + i / dk;
    qs[i] = q[(long)r * ldq + h * dk + (i % dk)];
  }
  for (int i = tid; i < n * dk4; i += blockDim.x) {
    const int j = i / dk4, c = (i % dk4) * 4;
    const KvT* row = kv + ((long)b * n + j) * ldkv + h * dk + c;
    float kk[4], vv[4];
    load4<KvT>(row + k_off, kk);       // one 8- or 16-byte load per operand (offsets are multiples of 4 elements)
    load4<KvT>(row + v_off, vv);
#pragma unroll
    for (int e = 0; e < 4; ++e) { Kt[(c + e) * kt + j] = kk[e]; Vs[j * dk + c + e] = vv[e]; }
  }
  const int nv = n_valid ? n_valid[b] : n;
  const float sq = sqrtf((float)dk);
  __syncthreads();
  // scores of all rows in one pass over K (reference layers.py:279-284)
  for (int j = tid; j < n; j += blockDim.x) {
    float a[kMaxRpi];
#pragma unroll
    for (int i = 0; i < kMaxRpi; ++i) a[i] = 0.f;
    for (int c = 0; c < dk; ++c) {
      const float kvv = Kt[c * kt + j];
#pragma unroll
      for (int i = 0; i < kMaxRpi; ++i)
        if (i < rpi) a[i] = fmaf(qs[i * dk + c], kvv, a[i]);
    }
#pragma unroll
    for (int i = 0; i < kMaxRpi; ++i)
      if (i < rpi) {
        float v = a[i] / sq;
        const bool row_padded = row_len && p >= row_len[b * rpi + i];
        if (row_padded || j >= nv) v = kCrossFill;
        pr[i * n + j] = v;
      }
  }
  __syncthreads();
  // softmax: one warp per row
  for (int i = warp; i < rpi; i += (blockDim.x >> 5)) {
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, pr[i * n + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) { const float e = expf(pr[i * n + j] - mx); pr[i * n + j] = e; sum += e; }
    sum = warp_sum(sum);
    for (int j = lane; j < n; j += 32) pr[i * n + j] = pr[i * n + j] / sum;
  }
  __syncthreads();
  // out[i][c] = sum_j p[i][j] V[j][c]: blockDim/dk partial sums per column, all rows in one pass over V
  const int parts = blockDim.x / dk;
  const int c = tid % dk, part = tid / dk;
  if (part < parts) {
    float o[kMaxRpi];
#pragma unroll
    for (int i = 0; i < kMaxRpi; ++i) o[i] = 0.f;
    for (int j = part; j < n; j += parts) {
      const float vv = Vs[j * dk + c];
#pragma unroll
      for (int i = 0; i < kMaxRpi; ++i)
        if (i < rpi) o[i] = fmaf(pr[i * n + j], vv, o[i]);
    }
#pragma unroll
    for (int i = 0; i < kMaxRpi; ++i)
      if (i < rpi) po[(part * kMaxRpi + i) * dk + c] = o[i];
  }
  __syncthreads();
  for (int i = tid; i < rpi * dk; i += blockDim.x) {
    const int ri = i / dk, cc = i % dk;
    float o = 0.f;
    for (int pp = 0; pp < parts; ++pp) o += po[(pp * kMaxRpi + ri) * dk + cc];   // fixed order: deterministic
    out[(long)(b * rpi + ri) * ldo + h * dk + cc] = from_f32<OutT>(o);
  }
}

template <typename T, int RPI>
__global__ void __launch_bounds__(256) cross_attn_step16_kernel(const float* __restrict__ q, long ldq,
                                                                const T* __restrict__ kv, long ldkv, int k_off, int v_off,
                                                                T* __restrict__ out, long ldo, int n,
                                                                const int* __restrict__ n_valid,
                                                                const int* __restrict__ row_len, int p) {
  pdl_wait();
  pdl_trigger();
  __shared__ float smf[cross16_smem_floats<RPI>()];
  cross_attn16_item<T, RPI>(q, ldq, kv, ldkv, k_off, v_off, out, ldo, n, n_valid, row_len, p, blockIdx.x, blockIdx.y,
                            blockIdx.x * RPI, smf);
}


template <typename T, int RPI>
static cudaError_t launch_ca16(const float* q, long ldq, const T* kv, long ldkv, int k_off, int v_off, T* out, long ldo, int R,
                               int n_keys, int heads, const int* n_valid, const int* row_len, int p, cudaStream_t st) {
  launch_k(cross_attn_step16_kernel<T, RPI>, dim3(dim3(R / RPI, heads)), dim3(256), 0, st, q, ldq, kv, ldkv, k_off, v_off, out, ldo, n_keys, n_valid, row_len, p);
  return cudaGetLastError();
}
template <typename T>
static bool try_ca16(cudaError_t* err, const float* q, long ldq, const T* kv, long ldkv, int k_off, int v_off, T* out, long ldo,
                     int R, int rpi, int n_keys, int heads, int dk, const int* n_valid, const int* row_len, int p, cudaStream_t st) {
  if (dk != 64 || n_keys > kCaMaxKeys || (ldkv & 7) || (k_off & 7) || (v_off & 7) || (ldq & 3) ||
      (reinterpret_cast<uintptr_t>(kv) & 15) || (reinterpret_cast<uintptr_t>(q) & 15))
    return false;
  switch (rpi) {
#define XN_CA16(N) case N: *err = launch_ca16<T, N>(q, ldq, kv, ldkv, k_off, v_off, out, ldo, R, n_keys, heads, n_valid, row_len, p, st); return true;
    XN_CA16(1) XN_CA16(2) XN_CA16(3) XN_CA16(4) XN_CA16(5) XN_CA16(6) XN_CA16(7) XN_CA16(8)
#undef XN_CA16
  }
  return false;
}
template <typename KvT, typename OutT> struct Ca16Dispatch {
  static bool run(cudaError_t*, const float*, long, const KvT*, long, int, int, OutT*, long, int, int, int, int, int, const int*,
                  const int*, int, cudaStream_t) { return false; }
};
template <> struct Ca16Dispatch<f16, f16> {
  static bool run(cudaError_t* e, const float* q, long ldq, const f16* kv, long ldkv, int k_off, int v_off, f16* out, long ldo, int R,
                  int rpi, int n, int heads, int dk, const int* nvp, const int* rl, int p, cudaStream_t st) {
    return try_ca16<f16>(e, q, ldq, kv, ldkv, k_off, v_off, out, ldo, R, rpi, n, heads, dk, nvp, rl, p, st);
  }
};
template <> struct Ca16Dispatch<bf16, bf16> {
  static bool run(cudaError_t* e, const float* q, long ldq, const bf16* kv, long ldkv, int k_off, int v_off, bf16* out, long ldo, int R,
                  int rpi, int n, int heads, int dk, const int* nvp, const int* rl, int p, cudaStream_t st) {
    return try_ca16<bf16>(e, q, ldq, kv, ldkv, k_off, v_off, out, ldo, R, rpi, n, heads, dk, nvp, rl, p, st);
  }
};

template <typename KvT, typename OutT>
cudaError_t launch_cross_attn_step(const float* q, long ldq, const KvT* kv, long ldkv, int k_off, int v_off,
                                   OutT* out, long ldo, int R, int rows_per_image, int n_keys, int heads, int dk,
                                   const int* n_valid, const int* row_len, int p, cudaStream_t st) {
  if (R % rows_per_image || (dk & 3) || dk > 256 || (256 % dk) || rows_per_image > kMaxRpi) return cudaErrorInvalidValue;
  {
    cudaError_t e16 = cudaSuccess;
    if (Ca16Dispatch<KvT, OutT>::run(&e16, q, ldq, kv, ldkv, k_off, v_off, out, ldo, R, rows_per_image, n_keys, heads, dk, n_valid,
                                     row_len, p, st))
      return e16;
  }
  const size_t smem = ((size_t)dk * (n_keys + 1) + (size_t)n_keys * dk + (size_t)kMaxRpi * dk + (size_t)kMaxRpi * n_keys +
                       (size_t)(256 / dk) * kMaxRpi * dk) * sizeof(float);
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(cross_attn_step_kernel<KvT, OutT>, smem, smem_state)) return e;
  launch_k(cross_attn_step_kernel<KvT, OutT>, dim3(dim3(R / rows_per_image, heads)), dim3(256), smem, st, 
      q, ldq, kv, ldkv, k_off, v_off, out, ldo, rows_per_image, n_keys, dk, n_valid, row_len, p);
  return cudaGetLastError();
}
template cudaError_t launch_cross_attn_step<float, float>(const float*, long, const float*, long, int, int, float*, long, int, int, int, int, int, const int*, const int*, int, cudaStream_t);
template cudaError_t launch_cross_attn_step<bf16, bf16>(const float*, long, const bf16*, long, int, int, bf16*, long, int, int, int, int, int, const int*, const int*, int, cudaStream_t);
template cudaError_t launch_cross_attn_step<f16, f16>(const float*, long, const f16*, long, int, int, f16*, long, int, int, int, int, int, const int*, const int*, int, cudaStream_t);

// ------------------------------------------------------------------------------------------
// log-softmax over the vocabulary + top-k (reference End_ExpansionNet_v2.py:204-207 and
// captioning_model.py:302-317).  One CTA per row; lp = (x - max) - log(sum exp(x - max)).
// Top-k order: descending value, ties to the lower index.
// ------------------------------------------------------------------------------------------
constexpr int kMaxTopK = 8;

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__global__ void __launch_bounds__(256) logsoftmax_topk_kernel(const float* __restrict__ logits, long ld, int V, int k,
                                                              float* __restrict__ top_val, int* __restrict__ top_idx,
                                                              float* __restrict__ logprob, long ldlp, int write_mode) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  __shared__ float cv[256];
  __shared__ int ci[256];
  __shared__ int win_tid;
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* x = logits + (long)r * ld;
  float lmax = -INFINITY;
  for (int i = tid; i < V; i += 256) lmax = fmaxf(lmax, x[i]);
  const float mx = block_max(lmax, red);
  float ls = 0.f;
  for (int i = tid; i < V; i += 256) ls += expf(x[i] - mx);
  const float lse = logf(block_sum(ls, red));
  __syncthreads();
  float tv[kMaxTopK];
  int ti[kMaxTopK];
#pragma unroll
  for (int j = 0; j < kMaxTopK; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
  for (int i = tid; i < V; i += 256) {
    const float v = (x[i] - mx) - lse;
    if (write_mode == 1) logprob[(long)r * ldlp + i] = v;
    if (k > 0 && better(v, i, tv[kMaxTopK - 1], ti[kMaxTopK - 1])) {
      tv[kMaxTopK - 1] = v; ti[kMaxTopK - 1] = i;
#pragma unroll
      for (int j = kMaxTopK - 1; j > 0; --j) {
        if (better(tv[j], ti[j], tv[j - 1], ti[j - 1])) {
          const float fv = tv[j]; tv[j] = tv[j - 1]; tv[j - 1] = fv;
          const int fi = ti[j]; ti[j] = ti[j - 1]; ti[j - 1] = fi;
        }
      }
    }
  }
  for (int round = 0; round < k; ++round) {
    cv[tid] = tv[0]; ci[tid] = ti[0];
    __syncthreads();
    if (tid < 32) {
      float bv = -INFINITY; int bi = 0x7fffffff, bt = 0;
      for (int t = tid; t < 256; t += 32)
        if (better(cv[t], ci[t], bv, bi)) { bv = cv[t]; bi = ci[t]; bt = t; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const int ot = __shfl_xor_sync(0xffffffffu, bt, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; bt = ot; }
      }
      if (tid == 0) {
        top_val[(long)r * k + round] = bv;
        top_idx[(long)r * k + round] = bi;
        win_tid = bt;
      }
    }
    __syncthreads();
    if (tid == win_tid) {
#pragma unroll
      for (int j = 0; j < kMaxTopK - 1; ++j) { tv[j] = tv[j + 1]; ti[j] = ti[j + 1]; }
      tv[kMaxTopK - 1] = -INFINITY; ti[kMaxTopK - 1] = 0x7fffffff;
    }
    __syncthreads();
  }
}

// Register-resident variant for V <= 256 * 4 * kLsVec: the row is read from memory exactly once (float4 loads, all in
// flight together), max / sum / top-k run on registers.  Same arithmetic as above: lp = (x - max) - log(sum exp(x - max)).
constexpr int kLsVec = 10;             // float4 per thread: 256 * 40 = 10240 logits

__global__ void __launch_bounds__(256) logsoftmax_topk_reg_kernel(const float* __restrict__ logits, long ld, int V, int k,
                                                                  float* __restrict__ top_val, int* __restrict__ top_idx,
                                                                  float* __restrict__ logprob, long ldlp, int write_mode) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  __shared__ float wv[8];
  __shared__ int wi[8];
  __shared__ int win_idx;
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + (long)r * ld;
  float v[kLsVec][4];
#pragma unroll
  for (int j = 0; j < kLsVec; ++j) {
    const int i = (j * 256 + tid) * 4;
    if (i + 3 < V) {
      const float4 f = *reinterpret_cast<const float4*>(x + i);
      v[j][0] = f.x; v[j][1] = f.y; v[j][2] = f.z; v[j][3] = f.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[j][e] = (i + e < V) ? x[i + e] : -INFINITY;
    }
  }
  float lmax = -INFINITY;
#pragma unroll
  for (int j = 0; j < kLsVec; ++j) lmax = fmaxf(lmax, fmaxf(fmaxf(v[j][0], v[j][1]), fmaxf(v[j][2], v[j][3])));
  const bool raw = (write_mode & 2) != 0;            // input already holds log-probabilities (ensemble path): top-k only
  float mx = 0.f, lse = 0.f;
  if (!raw) {
    mx = block_max(lmax, red);
    float ls = 0.f;
#pragma unroll
    for (int j = 0; j < kLsVec; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) ls += expf(v[j][e] - mx);            // exp(-inf) = 0 for the padding slots
    lse = logf(block_sum(ls, red));
  }
#pragma unroll
  for (int j = 0; j < kLsVec; ++j) {
#pragma unroll
    for (int e = 0; e < 4; ++e) if (!raw) v[j][e] = (v[j][e] - mx) - lse;
    const int i = (j * 256 + tid) * 4;
    if (write_mode & 1) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (i + e < V) logprob[(long)r * ldlp + i + e] = v[j][e];
    }
  }
  for (int round = 0; round < k; ++round) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < kLsVec; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = (j * 256 + tid) * 4 + e;
        if (i < V && better(v[j][e], i, bv, bi)) { bv = v[j][e]; bi = i; }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    __syncthreads();                       // previous round's win_idx has been consumed
    if (lane == 0) { wv[warp] = bv; wi[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float fv = wv[0];
      int fi = wi[0];
      for (int w = 1; w < 8; ++w)
        if (better(wv[w], wi[w], fv, fi)) { fv = wv[w]; fi = wi[w]; }
      top_val[(long)r * k + round] = fv;
      top_idx[(long)r * k + round] = fi;
      win_idx = fi;
    }
    __syncthreads();
    const int w = win_idx;                 // the owner retires the winner
#pragma unroll
    for (int j = 0; j < kLsVec; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if ((j * 256 + tid) * 4 + e == w) v[j][e] = -INFINITY;
  }
}

// ---- sampling: Gumbel-top-k over the row's log-probabilities (see kernels.h).  Philox4x32-10 keyed on the seed, counter
// (vocabulary block of 4, row, step): four uniforms per call, one per vocabulary entry.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float gumbel_from_bits(uint32_t r) {
  const float u = ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f);       // (0, 1) strictly, 24 bits
  return -logf(-logf(u));
}

__global__ void __launch_bounds__(256) gumbel_topk_kernel(const float* __restrict__ logits, long ld, int V, int k, uint2 seed, int step,
                                                          float* __restrict__ top_val, int* __restrict__ top_idx) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  __shared__ float cv[256];
  __shared__ int ci[256];
  __shared__ int win_tid;
  const int r = blockIdx.x, tid = threadIdx.x;
  const float* x = logits + (long)r * ld;
  float lmax = -INFINITY;
  for (int i = tid; i < V; i += 256) lmax = fmaxf(lmax, x[i]);
  const float mx = block_max(lmax, red);
  float ls = 0.f;
  for (int i = tid; i < V; i += 256) ls += expf(x[i] - mx);
  const float lse = logf(block_sum(ls, red));
  __syncthreads();
  float tv[kMaxTopK];
  int ti[kMaxTopK];
#pragma unroll
  for (int j = 0; j < kMaxTopK; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
  for (int i4 = tid; i4 * 4 < V; i4 += 256) {
    const uint4 rb = philox4x32_10(make_uint4((uint32_t)i4, (uint32_t)r, (uint32_t)step, 0x58433242u), seed);
    const uint32_t bits[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = i4 * 4 + e;
      if (i >= V) break;
      const float v = ((x[i] - mx) - lse) + gumbel_from_bits(bits[e]);
      if (better(v, i, tv[kMaxTopK - 1], ti[kMaxTopK - 1])) {
        tv[kMaxTopK - 1] = v; ti[kMaxTopK - 1] = i;
#pragma unroll
        for (int j = kMaxTopK - 1; j > 0; --j) {
          if (better(tv[j], ti[j], tv[j - 1], ti[j - 1])) {
            const float fv = tv[j]; tv[j] = tv[j - 1]; tv[j - 1] = fv;
            const int fi = ti[j]; ti[j] = ti[j - 1]; ti[j - 1] = fi;
          }
        }
      }
    }
  }
  for (int round = 0; round < k; ++round) {
    cv[tid] = tv[0]; ci[tid] = ti[0];
    __syncthreads();
    if (tid < 32) {
      float bv = -INFINITY; int bi = 0x7fffffff, bt = 0;
      for (int t = tid; t < 256; t += 32)
        if (better(cv[t], ci[t], bv, bi)) { bv = cv[t]; bi = ci[t]; bt = t; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const int ot = __shfl_xor_sync(0xffffffffu, bt, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; bt = ot; }
      }
      if (tid == 0) {
        top_val[(long)r * k + round] = (x[bi] - mx) - lse;       // the word's own log-probability, not the perturbed key
        top_idx[(long)r * k + round] = bi;
        win_tid = bt;
      }
    }
    __syncthreads();
    if (tid == win_tid) {
#pragma unroll
      for (int j = 0; j < kMaxTopK - 1; ++j) { tv[j] = tv[j + 1]; ti[j] = ti[j + 1]; }
      tv[kMaxTopK - 1] = -INFINITY; ti[kMaxTopK - 1] = 0x7fffffff;
    }
    __syncthreads();
  }
}
cudaError_t launch_gumbel_topk(const float* logits, long ld, int rows, int V, int k, uint64_t seed, int step, float* top_val,
                               int* top_idx, cudaStream_t st) {
  if (k < 1 || k > kMaxTopK || k > V) return cudaErrorInvalidValue;
  return launch_k(gumbel_topk_kernel, dim3(rows), dim3(256), 0, st, logits, ld, V, k,
                  make_uint2((uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)), step, top_val, top_idx);
}

// Ensemble step distribution (reference legacy_models/ensemble_captioning_model.py:55-84):
//   lp = log( mean_m softmax(logits_m) ),  softmax_m = exp(x - max_m) / sum_m,  mean = (p_0 + p_1 + ...) / n in model order.
// One CTA per row; each model's row is read twice from L2 (statistics, then the combination).
__global__ void __launch_bounds__(256) ensemble_logprob_kernel(EnsembleLogits in, long ld, int V, float* __restrict__ out, long ldo) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[32];
  __shared__ float s_mx[kMaxEnsemble], s_inv[kMaxEnsemble];
  const int r = blockIdx.x, tid = threadIdx.x;
  for (int m = 0; m < in.n; ++m) {
    const float* x = in.p[m] + (long)r * ld;
    float lmax = -INFINITY;
    for (int i = tid; i < V; i += 256) lmax = fmaxf(lmax, x[i]);
    const float mx = block_max(lmax, red);
    float ls = 0.f;
    for (int i = tid; i < V; i += 256) ls += expf(x[i] - mx);
    const float sum = block_sum(ls, red);
    if (tid == 0) { s_mx[m] = mx; s_inv[m] = sum; }
  }
  __syncthreads();
  const float nf = (float)in.n;
  for (int i = tid; i < V; i += 256) {
    float acc = 0.f;
    for (int m = 0; m < in.n; ++m) acc += expf(in.p[m][(long)r * ld + i] - s_mx[m]) / s_inv[m];
    out[(long)r * ldo + i] = logf(acc / nf);
  }
}
cudaError_t launch_ensemble_logprob(const EnsembleLogits& in, long ld, int rows, int V, float* out, long ldo, cudaStream_t st) {
  if (in.n < 1 || in.n > kMaxEnsemble) return cudaErrorInvalidValue;
  return launch_k(ensemble_logprob_kernel, dim3(rows), dim3(256), 0, st, in, ld, V, out, ldo);
}

cudaError_t launch_logsoftmax_topk(const float* logits, long ld, int rows, int V, int k, float* top_val, int* top_idx,
                                   float* logprob, long ldlp, int write_mode, cudaStream_t st) {
  if (k > kMaxTopK) return cudaErrorInvalidValue;
  if ((write_mode & 2) && !(V <= 256 * 4 * kLsVec && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0))
    return cudaErrorInvalidValue;            // top-k over ready log-probabilities is served by the register-resident kernel only
  if (V <= 256 * 4 * kLsVec && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0) {
    launch_k(logsoftmax_topk_reg_kernel, dim3(rows), dim3(256), 0, st, logits, ld, V, k, top_val, top_idx, logprob, ldlp, write_mode);
    return cudaGetLastError();
  }
  launch_k(logsoftmax_topk_kernel, dim3(rows), dim3(256), 0, st, logits, ld, V, k, top_val, top_idx, logprob, ldlp, write_mode);
  return cudaGetLastError();
}

}  // namespace xn
