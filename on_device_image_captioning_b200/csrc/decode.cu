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

namespace xn {

constexpr float kExpEps = 1e-9f;     // reference models/layers.py:208
constexpr float kCrossFill = -1e4f;  // reference models/layers.py:284

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

// Row LayerNorm tail shared by the fused decoder kernels: every thread holds NV values of the row (any column
// assignment), statistics over the whole CTA.  Same formula as layernorm_kernel (two-pass mean / variance, eps 1e-5).
template <int NV>
__device__ __forceinline__ void block_ln_stats(const float (&v)[NV], int d, float* red, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += v[i];
  mean = block_sum(s, red) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) { const float dd = v[i] - mean; q += dd * dd; }
  rstd = 1.0f / sqrtf(block_sum(q, red) / (float)d + 1e-5f);
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

// ------------------------------------------------------------------------------------------
// Dynamic expansion, incremental (reference models/layers.py:152-204, restated per position):
//   z[(i,e),j] = (q_e + c_i).K_j / sqrt(d)
//   forward  (row-block of p):  Af[(p,e),j] = relu(z)/(sum_{j<=p} relu(z) + eps)          -> cached
//   backward (output p):        ab[(i,e)]   = relu(z[(i,e),p]) / (sum_{i<=p,e} ... + eps)
//   out_a[p] = sum_{i,e} ab[(i,e)] * ( sum_{j<=i} Af[(i,e),j] A_j  + b_e + c_i )
//            = sum_j wA[j] A_j + sum_e sA[e] b_e + sum_i tA[i] c_i
//   with wA[j] = sum_{i>=j,e} ab[(i,e)] Af[(i,e),j],  sA[e] = sum_i ab[(i,e)],  tA[i] = sum_e ab[(i,e)]
// One CTA per row.
// ------------------------------------------------------------------------------------------
// When ln_out != nullptr (d == 2 * blockDim.x) the kernel also emits LayerNorm(x_out) in the operand type: the
// decoder layer's norm_2 (reference layers.py:238-241), fused because this CTA already holds the whole row.
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
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = s.P, np = p + 1;
  auto ln_tail = [&](float v0, float v1, float* red) {       // columns tid and tid + 256
    float v[2] = {v0, v1};
    float mean, rstd;
    block_ln_stats<2>(v, d, red, mean, rstd);
    ln_out[(long)r * ldn + tid] = from_f32<T>((v0 - mean) * rstd * ln_g[tid] + ln_b[tid]);
    ln_out[(long)r * ldn + tid + 256] = from_f32<T>((v1 - mean) * rstd * ln_g[tid + 256] + ln_b[tid + 256]);
  };
  if (row_len && p >= row_len[r]) {            // padded position: the block contributes 0 (all-zero mask rows)
    for (int c = tid; c < d; c += blockDim.x) x_out[(long)r * ldxo + c] = x_in[(long)r * ldxi + c];
    if (ln_out) ln_tail(x_in[(long)r * ldxi + tid], x_in[(long)r * ldxi + tid + 256], sm);
    return;
  }
  // carve-up in floats; P4 = P rounded up to 4 keeps the float4-read arrays (af, bf, part) 16-byte aligned
  const int P4 = (P + 3) & ~3, E4 = (n_exp + 3) & ~3;
  int* slot = reinterpret_cast<int*>(sm);      // [P4]
  float* ck_row = sm + P4;                     // [P4]  c_p . K_j
  float* ck_col = ck_row + P4;                 // [P4]  c_i . K_p
  float* qkp = ck_col + P4;                    // [E4]
  float* af = qkp + E4;                        // [P][n_exp] forward weights of the new row-block (A), key-major
  float* bf = af + P4 * E4;                    // [P][n_exp]
  float* ab = bf + P4 * E4;                    // [P][n_exp] backward weights (A)
  float* bb = ab + P4 * E4;                    // [P][n_exp]
  float* wA = bb + P4 * E4;                    // [P4]
  float* wB = wA + P4;
  float* tA = wB + P4;
  float* tB = tA + P4;
  float* sA = tB + P4;                         // [E4]
  float* sB = sA + E4;
  float* red = sB + E4;                        // [32]
  float* part = red + 32;                      // [2][P][P] partial forward-backward products; later >= 1536 floats of mix scratch

  for (int i = tid; i < np; i += blockDim.x) slot[i] = (i == p || !s.anc) ? r : s.anc[(long)r * P + i];
  __syncthreads();
  auto crow = [&](int i) { return s.cache + (((long)layer * P + i) * s.R + slot[i]) * s.cw; };
  const float* cp = crow(p);                   // [cond | key | A | B | sel] of the new position
  const float* Kp = cp + d;
  // The history rows this CTA will touch -- cond, key, A, B of every position -- are pulled towards L1 now, 128 bytes per
  // request, so the three phases below (each a dependent round of reads) hit L1 instead of paying an L2 trip apiece.
  if (s.cw == 5 * d) {
    const int lines_per_row = (4 * d * (int)sizeof(float)) / 128;
    for (int i = tid; i < np * lines_per_row; i += blockDim.x) {
      const float* a = crow(i / lines_per_row) + (i % lines_per_row) * 32;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
    }
  }

  // ---- phase A: the 2p+1+n_exp new dot products, 8 lanes each (4 concurrent per warp, 32 per CTA round)
  const int ntask = np + p + n_exp;
  const int sub = lane & 7, grp = tid >> 3, ngrp = blockDim.x >> 3;
  for (int t0 = 0; t0 < ntask; t0 += ngrp) {
    const int t = t0 + grp;
    float a = 0.f;
    if (t < ntask) {
      const float* u;
      const float* v;
      if (t < np) { u = cp; v = crow(t) + d; }                       // c_p . K_j
      else if (t < np + p) { u = crow(t - np); v = Kp; }             // c_i . K_p
      else { u = qexp + (long)(t - np - p) * d; v = Kp; }            // q_e . K_p
#pragma unroll 4
      for (int c = sub * 4; c < d; c += 32) {             // unrolled: 8 independent 16-byte loads in flight per lane
        const float4 x4 = *reinterpret_cast<const float4*>(u + c);
        const float4 y4 = *reinterpret_cast<const float4*>(v + c);
        a = fmaf(x4.x, y4.x, a); a = fmaf(x4.y, y4.y, a); a = fmaf(x4.z, y4.z, a); a = fmaf(x4.w, y4.w, a);
      }
    }
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    if (sub == 0 && t < ntask) {
      if (t < np) ck_row[t] = a;
      else if (t < np + p) ck_col[t - np] = a;
      else qkp[t - np - p] = a;
    }
  }
  __syncthreads();
  if (tid == 0) ck_col[p] = ck_row[p];
  float* qk_out = s.qk + (((long)layer * P + p) * s.R + r) * n_exp;
  for (int e = tid; e < n_exp; e += blockDim.x) qk_out[e] = qkp[e];
  __syncthreads();

  // ---- phase B: scalar work
  const float sq = sqrtf((float)d);
  // forward weights of the new row-block: one warp per expansion e
  float* fw_out = s.fw + (((long)layer * P + p) * s.R + r) * (2L * n_exp * P);
  for (int e = warp; e < n_exp; e += (blockDim.x >> 5)) {
    float za[4], sa = 0.f, sb = 0.f;               // P <= 128 -> up to 4 keys per lane
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = lane + 32 * c;
      float z = 0.f;
      if (j < np) {
        const float qk = (j == p) ? qkp[e] : s.qk[(((long)layer * P + j) * s.R + slot[j]) * n_exp + e];
        z = (qk + ck_row[j]) / sq;
        sa += fmaxf(z, 0.f);
        sb += fmaxf(-z, 0.f);
      }
      za[c] = z;
    }
    sa = warp_sum(sa) + kExpEps;
    sb = warp_sum(sb) + kExpEps;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = lane + 32 * c;
      if (j < np) {
        const float a = fmaxf(za[c], 0.f) / sa, b = fmaxf(-za[c], 0.f) / sb;
        af[j * n_exp + e] = a; bf[j * n_exp + e] = b;            // [key j][expansion e]: the 16 e of a key are contiguous
        fw_out[j * n_exp + e] = a; fw_out[(long)n_exp * P + j * n_exp + e] = b;
      }
    }
  }
  // backward weights for output position p (column p of z)
  float la = 0.f, lb = 0.f;
  for (int i = tid; i < np * n_exp; i += blockDim.x) {
    const int pi = i / n_exp, e = i % n_exp;
    const float z = (qkp[e] + ck_col[pi]) / sq;
    const float a = fmaxf(z, 0.f), b = fmaxf(-z, 0.f);
    ab[i] = a; bb[i] = b;
    la += a; lb += b;
  }
  const float ta = block_sum(la, red) + kExpEps;
  const float tb = block_sum(lb, red) + kExpEps;
  __syncthreads();
  for (int i = tid; i < np * n_exp; i += blockDim.x) { ab[i] = ab[i] / ta; bb[i] = bb[i] / tb; }
  __syncthreads();
  // wA[j] = sum_{i>=j} sum_e ab[(i,e)] * Af_i[e][j]: one task per (which, i, j<=i) pair, then a fixed-order sum over i
  for (int t = tid; t < 2 * np * np; t += blockDim.x) {
    const int which = t / (np * np), rem = t % (np * np), i = rem / np, j = rem % np;
    if (j > i) continue;
    const float* wsrc = which ? bb : ab;
    const float* f = (i == p) ? (which ? bf : af)
                              : s.fw + (((long)layer * P + i) * s.R + slot[i]) * (2L * n_exp * P) + (which ? (long)n_exp * P : 0);
    float acc = 0.f;
    const float* fj = f + j * n_exp;
    const float* wi = wsrc + i * n_exp;
    if ((n_exp & 3) == 0) {
      for (int e = 0; e < n_exp; e += 4) {
        const float4 f4 = *reinterpret_cast<const float4*>(fj + e);
        acc = fmaf(wi[e], f4.x, acc); acc = fmaf(wi[e + 1], f4.y, acc); acc = fmaf(wi[e + 2], f4.z, acc); acc = fmaf(wi[e + 3], f4.w, acc);
      }
    } else {
      for (int e = 0; e < n_exp; ++e) acc = fmaf(wi[e], fj[e], acc);
    }
    part[(which * P + i) * P + j] = acc;
  }
  __syncthreads();
  for (int t = tid; t < 2 * np; t += blockDim.x) {
    const int j = t >> 1, which = t & 1;
    float acc = 0.f;
    for (int i = j; i < np; ++i) acc += part[(which * P + i) * P + j];
    (which ? wB : wA)[j] = acc;
  }
  for (int t = tid; t < 2 * np; t += blockDim.x) {       // tA[i] = sum_e ab[(i,e)]
    const int i = t >> 1, which = t & 1;
    const float* wsrc = which ? bb : ab;
    float acc = 0.f;
    for (int e = 0; e < n_exp; ++e) acc += wsrc[i * n_exp + e];
    (which ? tB : tA)[i] = acc;
  }
  for (int t = tid; t < 2 * n_exp; t += blockDim.x) {    // sA[e] = sum_i ab[(i,e)]
    const int e = t >> 1, which = t & 1;
    const float* wsrc = which ? bb : ab;
    float acc = 0.f;
    for (int i = 0; i < np; ++i) acc += wsrc[i * n_exp + e];
    (which ? sB : sA)[e] = acc;
  }
  __syncthreads();

  // ---- phase C: the d-wide mixes.  Thread t owns the four columns 4*(t % 128) .. +3 and one half of the history
  // (t / 128): three independent 16-byte loads per position, unrolled so that a dozen are in flight; the two halves
  // meet in shared memory.  Needs d == 512 and 256 threads; other widths take the scalar loop.
  float keep[2] = {0.f, 0.f};
  if (d == 512 && blockDim.x == 256) {
    float4* mix = reinterpret_cast<float4*>(part);          // [2 (a|b)][128] partial sums of the upper half (part is free now)
    const int c4 = (tid & 127) * 4, half = tid >> 7;
    float4 oa = make_float4(0.f, 0.f, 0.f, 0.f), ob = oa;
    const int j0 = half ? (np + 1) / 2 : 0, j1 = half ? np : (np + 1) / 2;
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
      const float* cr = crow(j);
      const float4 cj = *reinterpret_cast<const float4*>(cr + c4);
      const float4 aj = *reinterpret_cast<const float4*>(cr + 2 * d + c4);
      const float4 bj = *reinterpret_cast<const float4*>(cr + 3 * d + c4);
      const float wa = wA[j], wb = wB[j], ta = tA[j], tb = tB[j];
      oa.x = fmaf(wa, aj.x, oa.x); oa.x = fmaf(ta, cj.x, oa.x); ob.x = fmaf(wb, bj.x, ob.x); ob.x = fmaf(tb, cj.x, ob.x);
      oa.y = fmaf(wa, aj.y, oa.y); oa.y = fmaf(ta, cj.y, oa.y); ob.y = fmaf(wb, bj.y, ob.y); ob.y = fmaf(tb, cj.y, ob.y);
      oa.z = fmaf(wa, aj.z, oa.z); oa.z = fmaf(ta, cj.z, oa.z); ob.z = fmaf(wb, bj.z, ob.z); ob.z = fmaf(tb, cj.z, ob.z);
      oa.w = fmaf(wa, aj.w, oa.w); oa.w = fmaf(ta, cj.w, oa.w); ob.w = fmaf(wb, bj.w, ob.w); ob.w = fmaf(tb, cj.w, ob.w);
    }
    if (half) {                                                // + the expansion-bias term, split the same way
      for (int e = 0; e < n_exp; ++e) {
        const float4 be = *reinterpret_cast<const float4*>(bexp + (long)e * d + c4);
        oa.x = fmaf(sA[e], be.x, oa.x); oa.y = fmaf(sA[e], be.y, oa.y); oa.z = fmaf(sA[e], be.z, oa.z); oa.w = fmaf(sA[e], be.w, oa.w);
        ob.x = fmaf(sB[e], be.x, ob.x); ob.y = fmaf(sB[e], be.y, ob.y); ob.z = fmaf(sB[e], be.z, ob.z); ob.w = fmaf(sB[e], be.w, ob.w);
      }
      mix[tid & 127] = oa;
      mix[128 + (tid & 127)] = ob;
    }
    __syncthreads();
    float* xo_s = reinterpret_cast<float*>(mix + 256);        // the row, for the LayerNorm tail's column assignment
    if (!half) {
      const float4 ua = mix[tid], ub = mix[128 + tid];
      oa.x += ua.x; oa.y += ua.y; oa.z += ua.z; oa.w += ua.w;
      ob.x += ub.x; ob.y += ub.y; ob.z += ub.z; ob.w += ub.w;
      const float4 sl4 = *reinterpret_cast<const float4*>(cp + 4 * d + c4);
      const float4 xi = *reinterpret_cast<const float4*>(x_in + (long)r * ldxi + c4);
      const float s0 = sigmoidf_(sl4.x), s1 = sigmoidf_(sl4.y), s2 = sigmoidf_(sl4.z), s3 = sigmoidf_(sl4.w);
      float4 xo;
      xo.x = xi.x + (s0 * oa.x + (1.0f - s0) * ob.x);
      xo.y = xi.y + (s1 * oa.y + (1.0f - s1) * ob.y);
      xo.z = xi.z + (s2 * oa.z + (1.0f - s2) * ob.z);
      xo.w = xi.w + (s3 * oa.w + (1.0f - s3) * ob.w);
      *reinterpret_cast<float4*>(x_out + (long)r * ldxo + c4) = xo;
      *reinterpret_cast<float4*>(xo_s + c4) = xo;
    }
    if (ln_out) {
      __syncthreads();
      keep[0] = xo_s[tid]; keep[1] = xo_s[tid + 256];
    }
  } else {
    for (int c = tid; c < d; c += blockDim.x) {
      float oa = 0.f, ob = 0.f;
      for (int j = 0; j < np; ++j) {
        const float* cr = crow(j);
        const float cj = cr[c];
        oa = fmaf(wA[j], cr[2 * d + c], oa); oa = fmaf(tA[j], cj, oa);
        ob = fmaf(wB[j], cr[3 * d + c], ob); ob = fmaf(tB[j], cj, ob);
      }
      for (int e = 0; e < n_exp; ++e) {
        const float be = bexp[(long)e * d + c];
        oa = fmaf(sA[e], be, oa);
        ob = fmaf(sB[e], be, ob);
      }
      const float sg = sigmoidf_(cp[4 * d + c]);
      const float xo = x_in[(long)r * ldxi + c] + (sg * oa + (1.0f - sg) * ob);
      x_out[(long)r * ldxo + c] = xo;
      if (c == tid) keep[0] = xo; else if (c == tid + 256) keep[1] = xo;
    }
  }
  if (ln_out) ln_tail(keep[0], keep[1], red);
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

// 16-bit K/V variant, dk == 64: no smem staging of K/V.  Eight lanes cover one key's 128-byte head slice with one 16-byte
// load each, a warp takes four keys per round and the CTA's eight warps 32; all K loads of a thread are issued before
// the first dot product and the V loads before the softmax, so the kernel pays roughly one L2 round trip per phase
// instead of a staged copy.  One CTA per (image, head) serves the image's RPI beam rows.
template <typename T> __device__ __forceinline__ void unpack8(const uint4& u, float (&o)[8]);
template <> __device__ __forceinline__ void unpack8<f16>(const uint4& u, float (&o)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void unpack8<bf16>(const uint4& u, float (&o)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}

constexpr int kCaRounds = 5;            // key rounds per warp: 8 warps x 4 keys x 5 rounds >= 160 keys
constexpr int kCaMaxKeys = 160;

template <typename T, int RPI>
__global__ void __launch_bounds__(256) cross_attn_step16_kernel(const float* __restrict__ q, long ldq,
                                                                const T* __restrict__ kv, long ldkv, int k_off, int v_off,
                                                                T* __restrict__ out, long ldo, int n,
                                                                const int* __restrict__ n_valid,
                                                                const int* __restrict__ row_len, int p) {
  pdl_wait();
  pdl_trigger();
  constexpr int dk = 64;
  __shared__ float pr[RPI][kCaMaxKeys];            // scores, then probabilities
  __shared__ float po[8][RPI][dk];                 // per-warp partial outputs
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & 7, kslot = lane >> 3;     // 8 dims per lane, 4 keys per warp round
  const T* base = kv + (long)b * n * ldkv + h * dk + sub * 8;
  uint4 kr[kCaRounds], vr[kCaRounds];
#pragma unroll
  for (int it = 0; it < kCaRounds; ++it) {
    const int j = (it * 8 + warp) * 4 + kslot;
    kr[it] = j < n ? *reinterpret_cast<const uint4*>(base + (long)j * ldkv + k_off) : make_uint4(0u, 0u, 0u, 0u);
  }
  float qv[RPI][8];
#pragma unroll
  for (int i = 0; i < RPI; ++i) {
    const float* qp = q + (long)(b * RPI + i) * ldq + h * dk + sub * 8;
    const float4 a = *reinterpret_cast<const float4*>(qp), c = *reinterpret_cast<const float4*>(qp + 4);
    qv[i][0] = a.x; qv[i][1] = a.y; qv[i][2] = a.z; qv[i][3] = a.w; qv[i][4] = c.x; qv[i][5] = c.y; qv[i][6] = c.z; qv[i][7] = c.w;
  }
  const int nv = n_valid ? n_valid[b] : n;
#pragma unroll
  for (int it = 0; it < kCaRounds; ++it) {
    const int j = (it * 8 + warp) * 4 + kslot;
    float kf[8];
    unpack8<T>(kr[it], kf);
    // V of the same key: in flight while the scores and the softmax are computed
    vr[it] = j < n ? *reinterpret_cast<const uint4*>(base + (long)j * ldkv + v_off) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < RPI; ++i) {
      float a = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) a = fmaf(qv[i][e], kf[e], a);
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      if (sub == 0 && j < n) {
        float v = a * 0.125f;                                        // / sqrt(64)
        const bool row_padded = row_len && p >= row_len[b * RPI + i];
        if (row_padded || j >= nv) v = kCrossFill;
        pr[i][j] = v;
      }
    }
  }
  __syncthreads();
  for (int i = warp; i < RPI; i += 8) {              // softmax: one warp per row
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, pr[i][j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) { const float e = expf(pr[i][j] - mx); pr[i][j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < n; j += 32) pr[i][j] *= inv;
  }
  __syncthreads();
  float acc[RPI][8];
#pragma unroll
  for (int i = 0; i < RPI; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;
#pragma unroll
  for (int it = 0; it < kCaRounds; ++it) {
    const int j = (it * 8 + warp) * 4 + kslot;
    if (j < n) {
      float vf[8];
      unpack8<T>(vr[it], vf);
#pragma unroll
      for (int i = 0; i < RPI; ++i) {
        const float w = pr[i][j];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[i][e] = fmaf(w, vf[e], acc[i][e]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RPI; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a = acc[i][e];
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      if (kslot == 0) po[warp][i][sub * 8 + e] = a;
    }
  __syncthreads();
  for (int i = tid; i < RPI * dk; i += 256) {
    const int ri = i / dk, cc = i % dk;
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) o += po[w][ri][cc];               // fixed order: deterministic
    out[(long)(b * RPI + ri) * ldo + h * dk + cc] = from_f32<T>(o);
  }
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
