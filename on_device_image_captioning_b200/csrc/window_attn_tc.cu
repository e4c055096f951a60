// Swin (shifted-)window attention on the 5th-generation tensor cores (tcgen05.mma, scores and probabilities in TMEM)
// for the 16-bit modes.  Same contract as window_attn.cu / window_attn_mma.cu: reads only Q/K/V, writes only O; cyclic
// shift, window partition / reverse, relative-position bias and the shift mask are index arithmetic
// (reference models/swin_transformer_mod.py:222-269, 397-437).
//
// Work item = (window, PAIR of heads): the pair's Q, K, V columns are 64 contiguous 16-bit values = one 128-byte row per
// token, i.e. exactly a 128-byte-swizzled UMMA operand row.  One persistent CTA per SM (12 warps) loops over items:
//
//   warp 11      producer   token addresses of the window, cp.async gather of the three 144 x 128 B tiles (+ the pair's
//                           bias tables, the shift labels) into one of two smem stages
//   warp 10      issuer     one thread: S_h = Q_h K_h^T (M 128 x N 144 x K 32, two UMMA k-steps inside the swizzle atom) for
//                           both heads into TMEM, later O_h = P_h V_h (M 128 x N 32 x K 144; A = P straight from TMEM, B = V
//                           as an MN-major operand -- the tile is stored token-major, no transpose)
//   warps 0-3    softmax of head A, warps 4-7 of head B: THREAD PER ROW (TMEM lane = query row 0..127): tcgen05.ld the
//                           fp32 scores in 16-column slices, add the relative-position bias (one LDS per element, the
//                           table index is row constant - compile-time key constant) and the shift mask, row max and row
//                           sum are plain per-thread reductions (no shuffles), probabilities go back into the same TMEM
//                           columns as packed 16-bit pairs; after the P.V MMA the thread reads its 32 outputs, normalises
//                           and stores 64 contiguous bytes
//   warps 8, 9   the 16 query rows 128..143 of head A / head B that do not fit the 128-row UMMA tile: mma.sync m16n8k16 on
//                           the same smem tiles (exactly one m16 row block), scores in registers
//
// TMEM: S_A 144 | S_B 144 | O_A 32 | O_B 32 columns.  Per element of the 144 x 144 score matrix the softmax threads issue
// LDS + FADD + FMNMX (pass 1) and FFMA + MUFU.EX2 + FADD + half a CVT (pass 2); the 72-register score tile, the quad
// shuffles, the ldmatrix traffic and the HMMA issue slots of the mma.sync kernel are gone.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <type_traits>
#include "kernels.h"
#include "common.cuh"
#include "mma16_frag.cuh"
#include "tcgen05_ptx.cuh"

namespace xn {
using namespace tc5;

namespace wtc {
constexpr int kThreads = 384;
constexpr uint32_t kRowBytes = 128;
constexpr uint32_t kTileBytes = kWinTok * kRowBytes;            // 18432 = 18 x 1024
constexpr uint32_t kStageData = 3 * kTileBytes;                 // Q | K | V of a head pair
constexpr int kBiasP = 532;                                     // floats per head of the derived table
constexpr int kLabP = 24;                                       // 16-bit elements per label row (48 B)
constexpr int kBiasN = (2 * kWin - 1) * (2 * kWin - 1);
constexpr uint32_t kMetaTok = 0;                                // int[144]
constexpr uint32_t kMetaBias = 576;                             // float[2][532]
constexpr uint32_t kMetaLab = kMetaBias + 2 * kBiasP * 4;       // 16-bit [144][24]
constexpr uint32_t kMetaFlag = kMetaLab + kWinTok * kLabP * 2;  // int[8]: masked, seam_y, seam_x, head0
constexpr uint32_t kMetaBytes = kMetaFlag + 32;
constexpr uint32_t kSmemMeta = 2 * kStageData;
constexpr uint32_t kSmemJneg = kSmemMeta + 2 * kMetaBytes;      // int[144]
constexpr uint32_t kSmemBars = kSmemJneg + kWinTok * 4;
constexpr uint32_t kSmemTotal = kSmemBars + 128;
static_assert(kMetaBytes % 16 == 0 && kMetaLab % 16 == 0 && kSmemBars % 8 == 0, "alignment");
constexpr uint32_t kColS = 0, kColO = 288, kTmemCols = 512;     // S_h at 144 h, O_h at 288 + 32 h
constexpr float kLabelVal = 24.0f;                              // 24 * 24 = 576 per matching axis (x scale = 101.8 > 100)
constexpr float kMaskStep = 576.0f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kScale = 0.17677669529663687f;                  // 32^-0.5
}  // namespace wtc

// K-major operand descriptor (known-good form of gemm_tcgen05.cu: LBO unused)
__device__ __forceinline__ uint64_t wtc_desc_k(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <typename T>
__global__ void __launch_bounds__(wtc::kThreads, 1)
window_attention_tc_kernel(const T* __restrict__ qkv, const float* __restrict__ bias_l2, T* __restrict__ out, int H, int C, int heads,
                           int shift, int n_items) {
  using namespace wtc;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  if (sbase & 1023u) __trap();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bars = sbase + kSmemBars;
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (2 + s); };
  auto bar_sfull = [&](int h) { return bars + 8u * (4 + h); };
  auto bar_pfull = [&](int h) { return bars + 8u * (6 + h); };
  auto bar_ofull = [&](int h) { return bars + 8u * (8 + h); };
  auto bar_ofree = [&](int h) { return bars + 8u * (10 + h); };
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + kSmemBars + 96);
  int* jneg = reinterpret_cast<int*>(smem + kSmemJneg);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(bar_full(s), 32); mbar_init(bar_empty(s), 11); }
    for (int h = 0; h < 2; ++h) { mbar_init(bar_sfull(h), 1); mbar_init(bar_pfull(h), 4); mbar_init(bar_ofull(h), 1); mbar_init(bar_ofree(h), 4); }
    fence_mbar_init();
  }
  if (warp == 10) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // one-time: label tiles zeroed (columns 4..23 stay zero for ever), key offsets of the bias index
  for (int s = 0; s < 2; ++s)
    for (int i = tid; i < kWinTok * kLabP / 2; i += kThreads) reinterpret_cast<uint32_t*>(smem + kSmemMeta + s * kMetaBytes + kMetaLab)[i] = 0u;
  if (tid < kWinTok) jneg[tid] = -4 * ((tid / kWin) * (2 * kWin - 1) + (tid % kWin));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  pdl_trigger();

  const int n_mine = blockIdx.x < n_items ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int nWs = H / kWin, nW = nWs * nWs, npairs = heads / 2;

  if (warp == 11) {
    // ================================================================ producer
    for (int n = 0; n < n_mine; ++n) {
      const int item = blockIdx.x + n * gridDim.x, s = n & 1;
      mbar_wait(bar_empty(s), (uint32_t)(((n >> 1) & 1) ^ 1));
      const int g = item % npairs, wid = item / npairs;
      const int b = wid / nW, wrem = wid - b * nW, wy = wrem / nWs, wx = wrem - wy * nWs;
      const bool ey = shift > 0 && wy == nWs - 1, ex = shift > 0 && wx == nWs - 1;
      unsigned char* meta = smem + kSmemMeta + s * kMetaBytes;
      int* tok = reinterpret_cast<int*>(meta + kMetaTok);
      for (int t = lane; t < kWinTok; t += 32) {
        const int ty = t / kWin, tx = t % kWin;
        int hh = wy * kWin + ty + shift, ww = wx * kWin + tx + shift;           // roll(-shift): shifted[h] = x[h + shift]
        if (hh >= H) hh -= H;
        if (ww >= H) ww -= H;
        tok[t] = (b * H + hh) * H + ww;
        const float ya = ey ? (ty < kWin - shift ? kLabelVal : 0.f) : kLabelVal, yb = ey ? (ty < kWin - shift ? 0.f : kLabelVal) : 0.f;
        const float xa = ex ? (tx < kWin - shift ? kLabelVal : 0.f) : kLabelVal, xb = ex ? (tx < kWin - shift ? 0.f : kLabelVal) : 0.f;
        uint2 v;
        v.x = Mma16<T>::pack(ya, yb);
        v.y = Mma16<T>::pack(xa, xb);
        *reinterpret_cast<uint2*>(meta + kMetaLab + t * kLabP * 2) = v;
      }
      if (lane == 0) {
        int* fl = reinterpret_cast<int*>(meta + kMetaFlag);
        fl[0] = (ey || ex) ? 1 : 0; fl[1] = ey ? 1 : 0; fl[2] = ex ? 1 : 0; fl[3] = 2 * g;
      }
      __syncwarp();
      const uint32_t data = sbase + s * kStageData;
      const int c = lane & 7, r_in = lane >> 3;
      const T* src0 = qkv + g * 64 + c * 8;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
#pragma unroll 4
        for (int k = 0; k < 36; ++k) {
          const int r = 4 * k + r_in;
          cp_async16(data + t * kTileBytes + r * kRowBytes + ((uint32_t)(c ^ (r & 7)) << 4), src0 + (long)tok[r] * 3 * C + t * C);
        }
      }
      const float* bsrc = bias_l2 + (long)(2 * g) * kBiasP;
      const uint32_t bdst = smem_u32(meta + kMetaBias);
      for (int i = lane; i < 2 * kBiasP / 4; i += 32) cp_async16(bdst + i * 16, bsrc + i * 4);
      cp_async_commit();
      cp_async_wait<0>();
      fence_async_smem();                     // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(bar_full(s));
    }
  } else if (warp == 10) {
    // ================================================================ MMA issuer (one thread)
    if (lane == 0) {
      constexpr int fmt = std::is_same<T, __half>::value ? 0 : 1;
      constexpr uint32_t idesc_qk = make_idesc(128, kWinTok, fmt, 0);
      constexpr uint32_t idesc_pv = make_idesc(128, kHeadDim, fmt, 1);
      for (int n = 0; n < n_mine; ++n) {
        const int s = n & 1;
        const uint32_t q_base = sbase + s * kStageData, k_base = q_base + kTileBytes, v_base = k_base + kTileBytes;
        mbar_wait(bar_full(s), (uint32_t)((n >> 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)           // +32 bytes per k-step, +64 bytes for the second head, inside the swizzle atom
            umma_ss(tmem_base + kColS + h * kWinTok, wtc_desc_k(q_base) + (uint64_t)(h * 4 + ks * 2), wtc_desc_k(k_base) + (uint64_t)(h * 4 + ks * 2),
                    idesc_qk, (uint32_t)(ks != 0));
          umma_commit(bar_sfull(h));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(bar_pfull(h), (uint32_t)(n & 1));               // P_h is in TMEM
          mbar_wait(bar_ofree(h), (uint32_t)((n & 1) ^ 1));         // O_h of the previous item has been read
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 9; ++ks)           // 16 keys per k-step: 8 TMEM columns of packed pairs, 16 token rows of V
            umma_ts(tmem_base + kColO + h * kHeadDim, tmem_base + kColS + h * kWinTok + ks * 8,
                    make_sw128_desc(v_base + ks * 2048) + (uint64_t)(h * 4), idesc_pv, (uint32_t)(ks != 0));
          umma_commit(bar_ofull(h));
        }
        umma_commit(bar_empty(s));                 // every MMA that reads this stage has retired
      }
    }
  } else if (warp >= 8) {
    // ================================================================ rows 128..143 of head hsel on mma.sync
    const int hsel = warp - 8;
    const int r0 = 128 + (lane >> 2), r1 = r0 + 8;
    const int rowoff0 = 4 * ((r0 / kWin) * (2 * kWin - 1) + (r0 % kWin) + (kWin - 1) * (2 * kWin - 1) + (kWin - 1));
    const int rowoff1 = 4 * ((r1 / kWin) * (2 * kWin - 1) + (r1 % kWin) + (kWin - 1) * (2 * kWin - 1) + (kWin - 1));
    const float sl = kScale * kLog2e;
    const uint32_t ones = Mma16<T>::pack(1.0f, 1.0f);
    const int l7 = lane & 7;
    for (int n = 0; n < n_mine; ++n) {
      const int s = n & 1;
      mbar_wait(bar_full(s), (uint32_t)((n >> 1) & 1));
      const unsigned char* meta = smem + kSmemMeta + s * kMetaBytes;
      const int* tok = reinterpret_cast<const int*>(meta + kMetaTok);
      const int* fl = reinterpret_cast<const int*>(meta + kMetaFlag);
      const bool masked = fl[0] != 0;
      const int head = fl[3] + hsel;
      const uint32_t q_base = sbase + s * kStageData, k_base = q_base + kTileBytes, v_base = k_base + kTileBytes;
      const unsigned char* bt = meta + kMetaBias + hsel * kBiasP * 4;
      const unsigned char* bt0 = bt + rowoff0;
      const unsigned char* bt1 = bt + rowoff1;
      const int* jq = jneg + (lane & 3) * 2;
      const float off = masked ? -2.0f * kMaskStep : 0.f;
      float sc[18][4];
#pragma unroll
      for (int nt = 0; nt < 18; ++nt) {
        const int jo = jq[nt * 8];                               // key 2q; key 2q+1 sits 4 bytes lower (same window row)
        sc[nt][0] = *reinterpret_cast<const float*>(bt0 + jo) + off;
        sc[nt][1] = *reinterpret_cast<const float*>(bt0 + jo - 4) + off;
        sc[nt][2] = *reinterpret_cast<const float*>(bt1 + jo) + off;
        sc[nt][3] = *reinterpret_cast<const float*>(bt1 + jo - 4) + off;
      }
      {
        const int row = 128 + l7 + ((lane >> 3) & 1) * 8;         // row & 7 == lane & 7
        if (masked) {
          const uint32_t l_base = smem_u32(meta + kMetaLab);
          uint32_t la[4];
          ldsm_x4(la, l_base + (row * kLabP + (lane >> 4) * 8) * 2);
          const uint32_t lb_base = l_base + (uint32_t)((l7 * kLabP + ((lane >> 3) & 1) * 8) * 2);
          static_for<18>([&](auto nt_c) {
            constexpr int nt = decltype(nt_c)::value;
            uint32_t lb[2];
            ldsm_x2_o<nt * 8 * kLabP * 2>(lb, lb_base);
            Mma16<T>::mma(sc[nt], la, lb[0], lb[1]);
          });
        }
        // A fragments of Q: k-step ks covers the head's 16-byte chunks 2 ks, 2 ks + 1
        uint32_t qa[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          ldsm_x4(qa[ks], q_base + row * kRowBytes + ((uint32_t)((hsel * 4 + ks * 2 + (lane >> 4)) ^ l7) << 4));
        // K fragments: key rows 8 nt + (lane & 7): the swizzle term depends on the lane only, tiles are 1024 bytes apart
        const uint32_t kaddr0 = k_base + l7 * kRowBytes + ((uint32_t)((hsel * 4 + 0 + ((lane >> 3) & 1)) ^ l7) << 4);
        const uint32_t kaddr1 = k_base + l7 * kRowBytes + ((uint32_t)((hsel * 4 + 2 + ((lane >> 3) & 1)) ^ l7) << 4);
        static_for<18>([&](auto nt_c) {
          constexpr int nt = decltype(nt_c)::value;
          uint32_t kb[2];
          ldsm_x2_o<nt * 1024>(kb, kaddr0);
          Mma16<T>::mma(sc[nt], qa[0], kb[0], kb[1]);
        });
        static_for<18>([&](auto nt_c) {
          constexpr int nt = decltype(nt_c)::value;
          uint32_t kb[2];
          ldsm_x2_o<nt * 1024>(kb, kaddr1);
          Mma16<T>::mma(sc[nt], qa[1], kb[0], kb[1]);
        });
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 18; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float nm0 = -mx0 * sl, nm1 = -mx1 * sl;
      uint32_t pa[18][2];
#pragma unroll
      for (int nt = 0; nt < 18; ++nt) {
        pa[nt][0] = Mma16<T>::pack(ex2_ftz(fmaf(sc[nt][0], sl, nm0)), ex2_ftz(fmaf(sc[nt][1], sl, nm0)));
        pa[nt][1] = Mma16<T>::pack(ex2_ftz(fmaf(sc[nt][2], sl, nm1)), ex2_ftz(fmaf(sc[nt][3], sl, nm1)));
      }
      float o[5][4];
#pragma unroll
      for (int i = 0; i < 5; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
      {
        // V^T fragments: key = 16 ks + 8 ((lane >> 3) & 1) + (lane & 7), dims chunk (lane >> 4) [+ 2 for d 16-31]
        const uint32_t vrow = (uint32_t)((((lane >> 3) & 1) * 8 + l7) * kRowBytes);
        const uint32_t vaddr0 = v_base + vrow + ((uint32_t)((hsel * 4 + (lane >> 4)) ^ l7) << 4);
        const uint32_t vaddr1 = v_base + vrow + ((uint32_t)((hsel * 4 + 2 + (lane >> 4)) ^ l7) << 4);
        static_for<9>([&](auto ks_c) {
          constexpr int ks = decltype(ks_c)::value;
          uint32_t a[4] = {pa[2 * ks][0], pa[2 * ks][1], pa[2 * ks + 1][0], pa[2 * ks + 1][1]};
          uint32_t vb0[4], vb1[4];
          ldsm_x4_trans_o<ks * 16 * 128>(vb0, vaddr0);
          ldsm_x4_trans_o<ks * 16 * 128>(vb1, vaddr1);
          Mma16<T>::mma(o[0], a, vb0[0], vb0[1]);
          Mma16<T>::mma(o[1], a, vb0[2], vb0[3]);
          Mma16<T>::mma(o[2], a, vb1[0], vb1[1]);
          Mma16<T>::mma(o[3], a, vb1[2], vb1[3]);
          Mma16<T>::mma(o[4], a, ones, ones);
        });
      }
      const float inv0 = 1.0f / o[4][0], inv1 = 1.0f / o[4][2];
      uint32_t* o0 = reinterpret_cast<uint32_t*>(out + (long)tok[r0] * C + head * kHeadDim + (lane & 3) * 2);
      uint32_t* o1 = reinterpret_cast<uint32_t*>(out + (long)tok[r1] * C + head * kHeadDim + (lane & 3) * 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o0[i * 4] = Mma16<T>::pack(o[i][0] * inv0, o[i][1] * inv0);      // columns 8 i + 2 (lane & 3), +1
        o1[i * 4] = Mma16<T>::pack(o[i][2] * inv1, o[i][3] * inv1);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty(s));
    }
  } else {
    // ================================================================ softmax, thread per row (TMEM lane = query row)
    const int hsel = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int yi = row / kWin, xi = row % kWin;
    const int rowc4 = 4 * ((yi + kWin - 1) * (2 * kWin - 1) + xi + kWin - 1);
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t tS = t_lane + kColS + hsel * kWinTok, tO = t_lane + kColO + hsel * kHeadDim;
    const float c2 = kScale * kLog2e;
    for (int n = 0; n < n_mine; ++n) {
      const int s = n & 1;
      mbar_wait(bar_full(s), (uint32_t)((n >> 1) & 1));
      const unsigned char* meta = smem + kSmemMeta + s * kMetaBytes;
      const int tokrow = reinterpret_cast<const int*>(meta + kMetaTok)[row];
      const int* fl = reinterpret_cast<const int*>(meta + kMetaFlag);
      const bool masked = fl[0] != 0;
      const int head = fl[3] + hsel;
      const unsigned char* bt = meta + kMetaBias + hsel * kBiasP * 4 + rowc4;
      // shift mask: -576 (raw score units; x scale = -101.8) per axis on which the key's region differs from the row's
      float mq[4] = {0.f, 0.f, 0.f, 0.f};
      if (masked) {
        const bool ey = fl[1] != 0, ex = fl[2] != 0;
        const bool ry = yi >= kWin - shift, rx = xi >= kWin - shift;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool ky = (q >> 1) != 0, kx = (q & 1) != 0;
          mq[q] = -kMaskStep * (float)((ey && ky != ry) + (ex && kx != rx));
        }
      }
      mbar_wait(bar_sfull(hsel), (uint32_t)(n & 1));
      tc_fence_after();
      // ---- pass 1: s + bias (+ mask) written back in place, row max
      float mx = -INFINITY;
      uint32_t v[2][16];
      tmem_ld16(tS, v[0]);
      static_for<9>([&](auto c_c) {
        constexpr int c = decltype(c_c)::value;
        tmem_ld_wait();
        if (c + 1 < 9) tmem_ld16(tS + (c + 1) * 16, v[(c + 1) & 1]);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int j = c * 16 + e, yj = j / kWin, xj = j % kWin;              // compile-time after unrolling
          float f = __uint_as_float(v[c & 1][e]) + *reinterpret_cast<const float*>(bt - 4 * (yj * (2 * kWin - 1) + xj));
          if (masked) f += mq[(yj >= kWin / 2 ? 2 : 0) + (xj >= kWin / 2 ? 1 : 0)];
          mx = fmaxf(mx, f);
          v[c & 1][e] = __float_as_uint(f);
        }
        tmem_st16(tS + c * 16, v[c & 1]);
      });
      tmem_st_wait();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty(s));          // bias table / token rows of this stage are no longer read
      // ---- pass 2: p = 2^((s - max) * scale * log2 e), packed pairs into the low half of the same columns
      const float negm = -mx * c2;
      float sum = 0.f;
      tmem_ld16(tS, v[0]);
      static_for<9>([&](auto c_c) {
        constexpr int c = decltype(c_c)::value;
        tmem_ld_wait();
        if (c + 1 < 9) tmem_ld16(tS + (c + 1) * 16, v[(c + 1) & 1]);
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float p0 = ex2_ftz(fmaf(__uint_as_float(v[c & 1][2 * e]), c2, negm));
          const float p1 = ex2_ftz(fmaf(__uint_as_float(v[c & 1][2 * e + 1]), c2, negm));
          sum += p0 + p1;
          pk[e] = Mma16<T>::pack(p0, p1);
        }
        tmem_st8(tS + c * 8, pk);
      });
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pfull(hsel));
      // ---- O = P V from TMEM, normalise, 64 contiguous bytes per row
      mbar_wait(bar_ofull(hsel), (uint32_t)(n & 1));
      tc_fence_after();
      uint32_t ov[32];
      tmem_ld32(tO, ov);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ofree(hsel));
      const float inv = 1.0f / sum;
      uint4* dst = reinterpret_cast<uint4*>(out + (long)tokrow * C + head * kHeadDim);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 w;
        w.x = Mma16<T>::pack(__uint_as_float(ov[8 * i + 0]) * inv, __uint_as_float(ov[8 * i + 1]) * inv);
        w.y = Mma16<T>::pack(__uint_as_float(ov[8 * i + 2]) * inv, __uint_as_float(ov[8 * i + 3]) * inv);
        w.z = Mma16<T>::pack(__uint_as_float(ov[8 * i + 4]) * inv, __uint_as_float(ov[8 * i + 5]) * inv);
        w.w = Mma16<T>::pack(__uint_as_float(ov[8 * i + 6]) * inv, __uint_as_float(ov[8 * i + 7]) * inv);
        dst[i] = w;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

int g_attn_tc = 1;

static int wtc_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

bool window_attention_tc_supported(int B, int H, int C, int heads, int shift) {
  return H % kWin == 0 && C == heads * kHeadDim && heads % 2 == 0 && (shift == 0 || shift == kWin / 2) && C % 8 == 0;
}

template <typename T>
cudaError_t launch_window_attention_tc(const T* qkv, const float* bias_l2, T* out, int B, int H, int C, int heads, int shift,
                                       cudaStream_t st) {
  if (!window_attention_tc_supported(B, H, C, heads, shift)) return cudaErrorInvalidValue;
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(window_attention_tc_kernel<T>, wtc::kSmemTotal, smem_state)) return e;
  const int nW = (H / kWin) * (H / kWin);
  const int n_items = B * nW * (heads / 2);
  const int grid = std::min(n_items, wtc_sm_count());
  return launch_k(window_attention_tc_kernel<T>, dim3(grid), dim3(wtc::kThreads), wtc::kSmemTotal, st, qkv, bias_l2, out, H, C, heads, shift,
                  n_items);
}
template cudaError_t launch_window_attention_tc<bf16>(const bf16*, const float*, bf16*, int, int, int, int, int, cudaStream_t);
template cudaError_t launch_window_attention_tc<__half>(const __half*, const float*, __half*, int, int, int, int, int, cudaStream_t);

}  // namespace xn
