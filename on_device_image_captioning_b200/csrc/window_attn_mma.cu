// Swin (shifted-)window attention on tensor cores for the 16-bit modes (bf16 / fp16 operands,
// fp32 accumulate and fp32 softmax).  Same contract as window_attn.cu: reads only Q/K/V, writes
// only O; shift, partition, relative-position bias and the shift mask are index arithmetic
// (reference models/swin_transformer_mod.py:222-269, 397-437).
//
// One CTA per (window, group of kHeadsPerCta heads), 9 warps; warp w owns query rows [16w, 16w+16).
// The window's token addresses and key descriptors are computed once; Q/K/V of head h+1 stream
// into the second smem buffer with cp.async while head h is computed.  Per warp and head:
//   S = Q K^T        16 x 144 x 32   -> 18 n8-tiles of mma.sync.m16n8k16, 72 fp32 accumulators
//   S = S*scale + bias(yi-yj, xi-xj) + mask ;  row softmax in registers (quad shuffles)
//   O = P V          16 x 32 x 144   -> P re-used straight from the accumulator layout as the
//                                        A fragment (tiles 2j, 2j+1 -> k-step j), V^T via ldmatrix.trans
// The 144x144 score matrix never leaves the register file.  (tcgen05 is not used here: the
// 144-row window does not map onto 128-row UMMA tiles without 44% padding, and the kernel is
// bounded by its Q/K/V reads, not by the MMA rate.)
#include <cuda_fp16.h>
#include "kernels.h"
#include "common.cuh"

namespace xn {

constexpr int kMmaThreads = 288;
constexpr int kRowPad = 40;                      // 16-bit elements per smem row (80 B): conflict-free ldmatrix
constexpr int kBiasN = (2 * kWin - 1) * (2 * kWin - 1);   // 529

template <typename T> struct Mma16;
template <> struct Mma16<bf16> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};
template <> struct Mma16<__half> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// bias_t is the relative-position table transposed to (heads, 529) so a head's column is contiguous
// kHeadsPerCta: 3 for Swin-L (6/12/24/48 heads), 2 or 1 for other head counts
template <typename T, int kHeadsPerCta>
__global__ void __launch_bounds__(kMmaThreads, 2) window_attention_mma_kernel(const T* __restrict__ qkv,
                                                                              const float* __restrict__ bias_t,
                                                                              T* __restrict__ out, int H, int C, int heads,
                                                                              int shift) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: stage[2] x {Q,K,V} x [144][40] 16-bit | bias[kHeadsPerCta][532] f32 | jinfo[144] | tok[144]
  constexpr int kMatElems = kWinTok * kRowPad;
  T* stage = reinterpret_cast<T*>(smem_raw);
  float* bt_all = reinterpret_cast<float*>(smem_raw + 2 * 3 * kMatElems * sizeof(T));
  int* jinfo = reinterpret_cast<int*>(bt_all + kHeadsPerCta * 532);   // (yj*23 + xj) | region label << 16, per key token
  int* tok = jinfo + kWinTok;

  const int nWs = H / kWin;
  const int wid = blockIdx.x, head0 = blockIdx.y * kHeadsPerCta;
  const int b = wid / (nWs * nWs), wrem = wid % (nWs * nWs);
  const int wy = wrem / nWs, wx = wrem % nWs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  auto token_row = [&](int t) {              // global token row of window token t (roll(-shift): shifted[h] = x[h+shift])
    const int ty = t / kWin, tx = t % kWin;
    const int h = (wy * kWin + ty + shift) % H, w = (wx * kWin + tx + shift) % H;
    return (b * H + h) * H + w;
  };
  // each thread stages 6 (token, q/k/v, 16-byte chunk) slots per head; offsets are recomputed per call
  // (a few integer ops) rather than kept live across the MMA section
  const uint32_t stage_u32 = (uint32_t)__cvta_generic_to_shared(stage);
  auto issue_loads = [&](int hh, int bufi) {
    const uint32_t base = stage_u32 + (uint32_t)(bufi * 3 * kMatElems * sizeof(T));
    const T* src = qkv + (head0 + hh) * kHeadDim;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int i = tid + k * kMmaThreads;
      const int t = i / 12, r = i % 12, which = r >> 2, ch = r & 3;
      cp_async16(base + (uint32_t)((which * kMatElems + t * kRowPad + ch * 8) * sizeof(T)),
                 src + (long)token_row(t) * 3 * C + which * C + ch * 8);
    }
    cp_async_commit();
  };
  issue_loads(0, 0);

  if (tid < kWinTok) {
    const int ty = tid / kWin, tx = tid % kWin;
    const int hs = wy * kWin + ty, ws_ = wx * kWin + tx;
    int lh = 0, lw = 0;
    if (shift > 0) {
      lh = hs < H - kWin ? 0 : (hs < H - shift ? 1 : 2);
      lw = ws_ < H - kWin ? 0 : (ws_ < H - shift ? 1 : 2);
    }
    jinfo[tid] = (ty * (2 * kWin - 1) + tx) | ((lh * 3 + lw) << 16);
    tok[tid] = token_row(tid);
  }
  for (int i = tid; i < kHeadsPerCta * kBiasN; i += kMmaThreads) {
    const int hh = i / kBiasN, e = i % kBiasN;
    bt_all[hh * 532 + e] = bias_t[(long)(head0 + hh) * kBiasN + e];
  }

  const int m0 = warp * 16;
  const int r0 = m0 + (lane >> 2), r1 = r0 + 8;
  const float scale = 0.17677669529663687f;      // 32^-0.5

  for (int hh = 0; hh < kHeadsPerCta; ++hh) {
    const int bufi = hh & 1;
    cp_async_wait<0>();
    __syncthreads();                              // head hh landed; everyone is done with the other buffer
    if (hh + 1 < kHeadsPerCta) issue_loads(hh + 1, bufi ^ 1);

    const uint32_t q_base = stage_u32 + (uint32_t)(bufi * 3 * kMatElems * sizeof(T));
    const uint32_t k_base = q_base + (uint32_t)(kMatElems * sizeof(T));
    const uint32_t v_base = k_base + (uint32_t)(kMatElems * sizeof(T));
    T* Qs = stage + bufi * 3 * kMatElems;
    const float* bt = bt_all + hh * 532;

    // A fragments of Q for the two k16 steps (d 0-15, 16-31)
    uint32_t qa[2][4];
    {
      const int row = m0 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int col = (lane >> 4) * 8;
      ldsm_x4(qa[0], q_base + (row * kRowPad + col) * 2);
      ldsm_x4(qa[1], q_base + (row * kRowPad + col + 16) * 2);
    }
    float s[18][4];
#pragma unroll
    for (int nt = 0; nt < 18; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      uint32_t kb[4];     // {b0,b1} for d 0-15 and {b0,b1} for d 16-31 of keys nt*8 .. nt*8+7
      const int krow = nt * 8 + (lane & 7), kcol = (lane >> 3) * 8;
      ldsm_x4(kb, k_base + (krow * kRowPad + kcol) * 2);
      Mma16<T>::mma(s[nt], qa[0], kb[0], kb[1]);
      Mma16<T>::mma(s[nt], qa[1], kb[2], kb[3]);
    }

    // scale + relative-position bias + shift mask, then the row softmax (rows r0, r1 = r0 + 8)
    // bias index = (yi - yj + 11)*23 + (xi - xj + 11) = rowbase_i - (yj*23 + xj)
    const int i0 = jinfo[r0], i1 = jinfo[r1];
    const float* bt0 = bt + (i0 & 0xffff) + (kWin - 1) * (2 * kWin - 1) + (kWin - 1);
    const float* bt1 = bt + (i1 & 0xffff) + (kWin - 1) * (2 * kWin - 1) + (kWin - 1);
    const int l0 = i0 >> 16, l1 = i1 >> 16;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 18; ++nt) {
      const int2 jj = *reinterpret_cast<const int2*>(&jinfo[nt * 8 + (lane & 3) * 2]);
      const int ja = jj.x & 0xffff, jb = jj.y & 0xffff, la = jj.x >> 16, lb = jj.y >> 16;
      float a0 = fmaf(s[nt][0], scale, bt0[-ja]);
      float a1 = fmaf(s[nt][1], scale, bt0[-jb]);
      float c0 = fmaf(s[nt][2], scale, bt1[-ja]);
      float c1 = fmaf(s[nt][3], scale, bt1[-jb]);
      if (la != l0) a0 += -100.0f;
      if (lb != l0) a1 += -100.0f;
      if (la != l1) c0 += -100.0f;
      if (lb != l1) c1 += -100.0f;
      s[nt][0] = a0; s[nt][1] = a1; s[nt][2] = c0; s[nt][3] = c1;
      mx0 = fmaxf(mx0, fmaxf(a0, a1));
      mx1 = fmaxf(mx1, fmaxf(c0, c1));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t pa[18][2];   // packed probabilities: [nt][0] = row r0 (cols 2q,2q+1), [nt][1] = row r1
#pragma unroll
    for (int nt = 0; nt < 18; ++nt) {
      const float e0 = __expf(s[nt][0] - mx0), e1 = __expf(s[nt][1] - mx0);
      const float e2 = __expf(s[nt][2] - mx1), e3 = __expf(s[nt][3] - mx1);
      sum0 += e0 + e1; sum1 += e2 + e3;
      pa[nt][0] = Mma16<T>::pack(e0, e1);
      pa[nt][1] = Mma16<T>::pack(e2, e3);
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

    // O = P V : 9 k16 steps over the keys, 4 n8 tiles over head_dim
    float o[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 9; ++ks) {
      uint32_t a[4] = {pa[2 * ks][0], pa[2 * ks][1], pa[2 * ks + 1][0], pa[2 * ks + 1][1]};
      const int vrow = ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
      const int vcol = (lane >> 4) * 8;
      uint32_t vb0[4], vb1[4];                     // d 0-15 and d 16-31
      ldsm_x4_trans(vb0, v_base + (vrow * kRowPad + vcol) * 2);
      ldsm_x4_trans(vb1, v_base + (vrow * kRowPad + vcol + 16) * 2);
      Mma16<T>::mma(o[0], a, vb0[0], vb0[1]);
      Mma16<T>::mma(o[1], a, vb0[2], vb0[3]);
      Mma16<T>::mma(o[2], a, vb1[0], vb1[1]);
      Mma16<T>::mma(o[3], a, vb1[2], vb1[3]);
    }
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;

    // stage the warp's 16 x 32 output tile in its own (now dead) Q rows, then 64-byte row stores
    __syncwarp();
    uint32_t* qw = reinterpret_cast<uint32_t*>(Qs);
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int col = n * 8 + (lane & 3) * 2;
      qw[(r0 * kRowPad + col) >> 1] = Mma16<T>::pack(o[n][0] * inv0, o[n][1] * inv0);
      qw[(r1 * kRowPad + col) >> 1] = Mma16<T>::pack(o[n][2] * inv1, o[n][3] * inv1);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int row = m0 + it * 8 + (lane >> 2), ch = lane & 3;
      const uint4 v = *reinterpret_cast<const uint4*>(Qs + row * kRowPad + ch * 8);
      *reinterpret_cast<uint4*>(out + (long)tok[row] * C + (head0 + hh) * kHeadDim + ch * 8) = v;
    }
  }
}

template <typename T, int HPC>
static cudaError_t launch_wa_hpc(const T* qkv, const float* bias_t, T* out, int B, int H, int C, int heads, int shift,
                                 cudaStream_t st) {
  const size_t smem = 2 * 3 * kWinTok * kRowPad * sizeof(T) + HPC * 532 * sizeof(float) + 2 * kWinTok * sizeof(int);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attention_mma_kernel<T, HPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int nW = (H / kWin) * (H / kWin);
  window_attention_mma_kernel<T, HPC><<<dim3(B * nW, heads / HPC), kMmaThreads, smem, st>>>(qkv, bias_t, out, H, C, heads, shift);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_window_attention_mma(const T* qkv, const float* bias_t, T* out, int B, int H, int C, int heads,
                                        int shift, cudaStream_t st) {
  if (H % kWin || C != heads * kHeadDim) return cudaErrorInvalidValue;
  if (heads % 3 == 0) return launch_wa_hpc<T, 3>(qkv, bias_t, out, B, H, C, heads, shift, st);
  if (heads % 2 == 0) return launch_wa_hpc<T, 2>(qkv, bias_t, out, B, H, C, heads, shift, st);
  return launch_wa_hpc<T, 1>(qkv, bias_t, out, B, H, C, heads, shift, st);
}
template cudaError_t launch_window_attention_mma<bf16>(const bf16*, const float*, bf16*, int, int, int, int, int, cudaStream_t);
template cudaError_t launch_window_attention_mma<__half>(const __half*, const float*, __half*, int, int, int, int, int, cudaStream_t);

}  // namespace xn

namespace xn {
// (529, heads) relative-position table -> (heads, 529), once per weight load
__global__ void transpose_bias_kernel(const float* __restrict__ t, float* __restrict__ o, int heads) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kBiasN * heads) { const int e = i / heads, h = i % heads; o[h * kBiasN + e] = t[i]; }
}
cudaError_t launch_transpose_bias(const float* table, float* out, int heads, cudaStream_t st) {
  transpose_bias_kernel<<<(kBiasN * heads + 255) / 256, 256, 0, st>>>(table, out, heads);
  return cudaGetLastError();
}
}  // namespace xn
