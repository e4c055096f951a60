// Row-level bodies of the decoder-step kernels, shared by the one-kernel-per-operation path (decode.cu) and the
// persistent whole-step kernel (decode_mega.cu).  `r` / (image, head) are explicit arguments instead of blockIdx.
#pragma once
#include "kernels.h"
#include "common.cuh"

namespace xn {

constexpr float kExpEps = 1e-9f;     // reference models/layers.py:208
constexpr float kCrossFill = -1e4f;  // reference models/layers.py:284

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
// One row per call, executed by a whole 256-thread CTA; `sm` = dyn_exp_smem_floats(P, n_exp) floats of shared memory.
// None of the activation pointers is __restrict__/const-cached: inside the persistent kernel they were written by other
// CTAs earlier in the same launch.
template <typename T>
__device__ __forceinline__ void dyn_exp_row(const DecState& s, int layer, int p, const float* __restrict__ qexp,
                                            const float* __restrict__ bexp, int n_exp, const int* row_len,
                                            const float* x_in, long ldxi, float* x_out, long ldxo, int d,
                                            const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                            T* ln_out, long ldn, int r, float* sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

__host__ __device__ inline size_t dyn_exp_smem_floats(int P, int n_exp) {
  const size_t P4 = (P + 3) & ~3, E4 = (n_exp + 3) & ~3;
  const size_t part = 2 * (size_t)P * P > 1536 ? 2 * (size_t)P * P : 1536;
  return P4 * 7 + E4 * 3 + 4 * P4 * E4 + 32 + part;
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
// One (image b, head h) item for the RPI rows row0 .. row0 + RPI - 1 of that image, executed by a 256-thread CTA.
// `smf` = cross16_smem_floats(RPI) floats of shared memory.
template <int RPI> __host__ __device__ constexpr int cross16_smem_floats() { return RPI * kCaMaxKeys + 8 * RPI * 64; }
template <typename T, int RPI>
__device__ __forceinline__ void cross_attn16_item(const float* q, long ldq, const T* kv, long ldkv, int k_off, int v_off,
                                                  T* out, long ldo, int n, const int* n_valid, const int* row_len, int p,
                                                  int b, int h, int row0, float* smf) {
  constexpr int dk = 64;
  float (*pr)[kCaMaxKeys] = reinterpret_cast<float (*)[kCaMaxKeys]>(smf);                         // scores, then probabilities
  float (*po)[RPI][dk] = reinterpret_cast<float (*)[RPI][dk]>(smf + RPI * kCaMaxKeys);            // per-warp partial outputs
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    const float* qp = q + (long)(row0 + i) * ldq + h * dk + sub * 8;
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
        const bool row_padded = row_len && p >= row_len[row0 + i];
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
    out[(long)(row0 + ri) * ldo + h * dk + cc] = from_f32<T>(o);
  }
}

}  // namespace xn
