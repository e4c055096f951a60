// Swin (shifted-)window multi-head self-attention, one CTA per (window, head).
//
// Replaces reference models/swin_transformer_mod.py:397-437 (roll, window_partition,
// window_reverse, roll back) and :222-269 (scale, q.k^T, relative-position bias gather,
// shift mask, softmax, .v).  Nothing but Q/K/V is read and nothing but O is written: the
// cyclic shift and the window partition are folded into the token addressing, the bias is
// table[(yi-yj+11)*23 + (xi-xj+11)][head] computed from coordinates, the shift mask is
// "-100 when the 3x3 region labels of the shifted frame differ" (swin:366-391).
//
// This file holds the fp32-accurate CUDA-core version (used by the fp32 parity mode, and by
// the bf16 mode until the tensor-core variant in window_attn_mma.cu takes over).
#include "kernels.h"
#include "common.cuh"

namespace xn {

constexpr int kWaThreads = 288;                 // 9 warps x 16 query rows = 144
constexpr int kKs = kHeadDim + 1;               // padded row stride (bank-conflict free)

template <typename T>
__global__ void __launch_bounds__(kWaThreads) window_attention_kernel(const T* __restrict__ qkv,
                                                                      const float* __restrict__ bias_table,
                                                                      T* __restrict__ out, int H, int C, int heads,
                                                                      int shift) {
  extern __shared__ float smem[];
  float* Qs = smem;                              // [144][33], pre-scaled by head_dim^-0.5
  float* Ks = Qs + kWinTok * kKs;
  float* Vs = Ks + kWinTok * kKs;
  float* Ps = Vs + kWinTok * kKs;                // [9][144] probabilities of the warp's current row
  float* bt = Ps + 9 * kWinTok;                  // [529] bias column of this head
  int* tok = reinterpret_cast<int*>(bt + 532);   // [144] global token row of every window token
  int* lab = tok + kWinTok;                      // [144] region label (shift mask)

  const int nWs = H / kWin;                      // windows per side
  const int wid = blockIdx.x;                    // b*nW + wy*nWs + wx
  const int head = blockIdx.y;
  const int b = wid / (nWs * nWs), wrem = wid % (nWs * nWs);
  const int wy = wrem / nWs, wx = wrem % nWs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid < kWinTok) {
    const int ty = tid / kWin, tx = tid % kWin;
    const int hs = wy * kWin + ty, ws_ = wx * kWin + tx;          // shifted-frame coordinates
    const int h = (hs + shift) % H, w = (ws_ + shift) % H;        // roll(-shift): shifted[h] = x[h+shift]
    tok[tid] = (b * H + h) * H + w;
    int lh = 0, lw = 0;
    if (shift > 0) {
      lh = hs < H - kWin ? 0 : (hs < H - shift ? 1 : 2);
      lw = ws_ < H - kWin ? 0 : (ws_ < H - shift ? 1 : 2);
    }
    lab[tid] = lh * 3 + lw;
  }
  for (int i = tid; i < (2 * kWin - 1) * (2 * kWin - 1); i += kWaThreads) bt[i] = bias_table[(long)i * heads + head];
  __syncthreads();

  const float scale = 0.17677669529663687f;      // 32^-0.5 (qk_scale=None at every call site)
  for (int i = tid; i < kWinTok * kHeadDim; i += kWaThreads) {
    const int t = i >> 5, d = i & 31;
    const T* base = qkv + (long)tok[t] * 3 * C + head * kHeadDim + d;
    Qs[t * kKs + d] = to_f32<T>(base[0]) * scale;
    Ks[t * kKs + d] = to_f32<T>(base[C]);
    Vs[t * kKs + d] = to_f32<T>(base[2 * C]);
  }
  __syncthreads();

  float* P = Ps + warp * kWinTok;
  for (int r = 0; r < 16; ++r) {
    const int i = warp * 16 + r;
    const int yi = i / kWin, xi = i % kWin, li = lab[i];
    float s[5];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const int j = lane + 32 * c;
      float a = -INFINITY;
      if (j < kWinTok) {
        a = 0.f;
#pragma unroll
        for (int d = 0; d < kHeadDim; ++d) a = fmaf(Qs[i * kKs + d], Ks[j * kKs + d], a);
        const int yj = j / kWin, xj = j % kWin;
        a += bt[(yi - yj + kWin - 1) * (2 * kWin - 1) + (xi - xj + kWin - 1)];
        if (lab[j] != li) a += -100.0f;
      }
      s[c] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const int j = lane + 32 * c;
      const float e = (j < kWinTok) ? expf(s[c] - mx) : 0.f;
      s[c] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const int j = lane + 32 * c;
      if (j < kWinTok) P[j] = s[c] / sum;
    }
    __syncwarp();
    float o = 0.f;
#pragma unroll 8
    for (int j = 0; j < kWinTok; ++j) o = fmaf(P[j], Vs[j * kKs + lane], o);
    out[(long)tok[i] * C + head * kHeadDim + lane] = from_f32<T>(o);
  }
}

template <typename T>
cudaError_t launch_window_attention(const T* qkv, const float* bias_table, T* out, int B, int H, int C, int heads,
                                    int shift, cudaStream_t st) {
  if (H % kWin || C != heads * kHeadDim) return cudaErrorInvalidValue;
  const size_t smem = (3 * kWinTok * kKs + 9 * kWinTok + 532) * sizeof(float) + 2 * kWinTok * sizeof(int);
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(window_attention_kernel<T>, smem, smem_state)) return e;
  const int nW = (H / kWin) * (H / kWin);
  window_attention_kernel<T><<<dim3(B * nW, heads), kWaThreads, smem, st>>>(qkv, bias_table, out, H, C, heads, shift);
  return cudaGetLastError();
}
template cudaError_t launch_window_attention<float>(const float*, const float*, float*, int, int, int, int, int, cudaStream_t);
template cudaError_t launch_window_attention<bf16>(const bf16*, const float*, bf16*, int, int, int, int, int, cudaStream_t);

}  // namespace xn
