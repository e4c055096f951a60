// Swin (shifted-)window attention on the 5th-generation tensor cores (tcgen05.mma, scores and probabilities in TMEM)
// for the 16-bit modes.  Same contract as window_attn.cu / window_attn_mma.cu: reads only Q/K/V, writes only O; cyclic
// shift, window partition / reverse, relative-position bias and the shift mask are index arithmetic
// (reference models/swin_transformer_mod.py:222-269, 397-437).
//
// Work item = (window, PAIR of heads): the pair's Q, K, V columns are 64 contiguous 16-bit values = one 128-byte row per
// token, i.e. exactly a 128-byte-swizzled UMMA operand row.  One persistent CTA per SM (24 warps) loops over items:
//
//   warps 21-23  producers  token addresses of the window, cp.async gather of the three 144 x 128 B tiles (+ the pair's
//                           bias tables, the shift labels) into one of two smem stages
//   warp 20      issuer     one thread: S_h = Q_h K_h^T (M 128 x N 144 x K 32, two UMMA k-steps inside the swizzle atom) for
//                           both heads into TMEM, later O_h = P_h V_h (M 128 x N 32 x K 144; A = P straight from TMEM, B = V
//                           as an MN-major operand -- the tile is stored token-major, no transpose)
//   warps 0-15   softmax: warps 0-3 / 8-11 head A, 4-7 / 12-15 head B.  TWO THREADS PER ROW (TMEM lane = query row 0..127,
//                           the warps m and m + 8 share a lane quarter and take 72 of the 144 columns each; row max and
//                           row sum are combined through shared memory and a named barrier of the pair): tcgen05.ld the
//                           fp32 scores in 16-column slices, add the relative-position bias (one LDS per element, the
//                           table index is row constant - compile-time key constant) and the shift mask, row max and row
//                           sum are plain per-thread reductions (no shuffles), probabilities go back into the same TMEM
//                           columns as packed 16-bit pairs; after the P.V MMA the thread reads its 32 outputs, normalises
//                           and stores 64 contiguous bytes
//   warps 16-19  the 16 query rows 128..143 of head A / head B that do not fit the 128-row UMMA tile: mma.sync m16n8k16 on
//                           the same smem tiles (exactly one m16 row block), scores in registers
//
// TMEM: S_A 144 | S_B 144 | P_A 72 | P_B 72 | O_A 32 | O_B 32 columns.  Per element of the 144 x 144 score matrix the softmax threads issue
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
// 16 warps = 4 per scheduler, 128 registers each: 8 softmax (one thread per row), 2 tail, issuer, one spare, 4 producers.
// Variants with 16 softmax warps (two threads per row, 24 warps, setmaxnreg), with four tail warps, and with the whole
// row held in 144 registers were all measured SLOWER (93.5 us per stage-3 launch here; 114.8, 121.2 and 108.8 us):
// profiles/README.md, "window attention on tcgen05".
constexpr int kThreads = 512;
constexpr int kMainWarps = 8, kTailWarp0 = 8, kIssuerWarp = 10, kProducerWarp0 = 12, kProducers = 4;
constexpr uint32_t kRowBytes = 128;
constexpr uint32_t kTileBytes = kWinTok * kRowBytes;            // 18432 = 18 x 1024
constexpr uint32_t kStageData = 3 * kTileBytes;                 // Q | K | V of a head pair
// three stages: the gather of item n+2 is in flight while item n is computed (with two, the softmax warps waited for
// the producers 2100 cycles per item: profiles/r2_ncu_wattn_tc_v3_summary.txt)
constexpr int kStages = 3;
// Relative-position bias table of a head as the softmax threads read it: entry (dy + 11) * 44 + (dx + 11).  A row pitch of
// 44 = 12 (mod 32) makes the word address of a thread's bias  const(key) + 44 yi + xi = const + row (mod 32): the 32 lanes
// of a warp (32 consecutive rows) hit 32 different banks, where the natural pitch 23 gave two-way conflicts on every load
// (the first version of this kernel was bound by exactly that: profiles/r2_ncu_wattn_tc_v2_summary.txt).
constexpr int kBiasRow = 44;
constexpr int kBiasP = 1024;                                    // floats per head (23 * 44 = 1012, padded)
constexpr int kLabP = 24;                                       // 16-bit elements per label row (48 B)
constexpr int kBiasN = (2 * kWin - 1) * (2 * kWin - 1);
constexpr uint32_t kMetaTok = 0;                                // int[144]
constexpr uint32_t kMetaBias = 576;                             // float[2][532]
constexpr uint32_t kMetaLab = kMetaBias + 2 * kBiasP * 4;       // 16-bit [144][24]
constexpr uint32_t kMetaFlag = kMetaLab + kWinTok * kLabP * 2;  // int[8]: masked, seam_y, seam_x, head0
constexpr uint32_t kMetaBytes = kMetaFlag + 32;
constexpr uint32_t kSmemMeta = kStages * kStageData;
constexpr uint32_t kSmemJneg = kSmemMeta + kStages * kMetaBytes; // int[144]
constexpr uint32_t kSmemBars = kSmemJneg + kWinTok * 4;
constexpr uint32_t kSmemTotal = kSmemBars + 160;
static_assert(kMetaBytes % 16 == 0 && kMetaLab % 16 == 0 && kSmemBars % 8 == 0, "alignment");
static_assert(kSmemTotal <= 232448, "exceeds the 227 KB dynamic shared memory limit");
// TMEM columns: scores S_h at 144 h, probabilities P_h (packed 16-bit pairs) at 288 + 72 h, outputs O_h at 432 + 32 h
constexpr uint32_t kColS = 0, kColP = 288, kColO = 432, kTmemCols = 512;
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
window_attention_tc_kernel(const T* __restrict__ qkv, const float* __restrict__ bias_tc, T* __restrict__ out, int H, int C, int heads,
                           int shift, int n_items, int dbg) {
  using namespace wtc;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  if (sbase & 1023u) __trap();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bars = sbase + kSmemBars;
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (kStages + s); };
  auto bar_sfull = [&](int h) { return bars + 8u * (2 * kStages + h); };
  auto bar_pfull = [&](int h) { return bars + 8u * (2 * kStages + 2 + h); };
  auto bar_ofull = [&](int h) { return bars + 8u * (2 * kStages + 4 + h); };
  auto bar_ofree = [&](int h) { return bars + 8u * (2 * kStages + 6 + h); };
  auto bar_sfree = [&](int h) { return bars + 8u * (2 * kStages + 8 + h); };
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + kSmemBars + 152);
  int* jneg = reinterpret_cast<int*>(smem + kSmemJneg);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_full(s), 32 * kProducers); mbar_init(bar_empty(s), 1 + kMainWarps + 2); }
    for (int h = 0; h < 2; ++h) { mbar_init(bar_sfull(h), 1); mbar_init(bar_pfull(h), 4); mbar_init(bar_ofull(h), 1); mbar_init(bar_ofree(h), 4); }
    for (int h = 0; h < 2; ++h) mbar_init(bar_sfree(h), 4);
    fence_mbar_init();
  }
  if (warp == kIssuerWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // one-time: label tiles zeroed (columns 4..23 stay zero for ever), key offsets of the bias index
  for (int s = 0; s < kStages; ++s)
    for (int i = tid; i < kWinTok * kLabP / 2; i += kThreads) reinterpret_cast<uint32_t*>(smem + kSmemMeta + s * kMetaBytes + kMetaLab)[i] = 0u;
  if (tid < kWinTok) jneg[tid] = -4 * ((tid / kWin) * kBiasRow + (tid % kWin));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  pdl_trigger();

  const int n_mine = blockIdx.x < n_items ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int nWs = H / kWin, nW = nWs * nWs, npairs = heads / 2;

  if (warp >= kProducerWarp0) {
    // ================================================================ producers: four warps share the gather of an item
    // (one warp issuing all 3456 16-byte copies was the bottleneck of the first version: 7.5 us per item, the softmax
    // warps idle on the stage barrier -- profiles/r2_ncu_wattn_tc_v1_summary.txt)
    const int pw = warp - kProducerWarp0;
    for (int n = 0; n < n_mine; ++n) {
      const int item = blockIdx.x + n * gridDim.x, s = n % kStages;
      mbar_wait(bar_empty(s), (uint32_t)(((n / kStages) & 1) ^ 1));
      const int g = item % npairs, wid = item / npairs;
      const int b = wid / nW, wrem = wid - b * nW, wy = wrem / nWs, wx = wrem - wy * nWs;
      const bool ey = shift > 0 && wy == nWs - 1, ex = shift > 0 && wx == nWs - 1;
      unsigned char* meta = smem + kSmemMeta + s * kMetaBytes;
      int* tok = reinterpret_cast<int*>(meta + kMetaTok);
      // every producer warp writes the whole (identical) token table, so each reads back its own writes after a warp sync
      for (int t = lane; t < kWinTok; t += 32) {
        const int ty = t / kWin, tx = t % kWin;
        int hh = wy * kWin + ty + shift, ww = wx * kWin + tx + shift;           // roll(-shift): shifted[h] = x[h + shift]
        if (hh >= H) hh -= H;
        if (ww >= H) ww -= H;
        tok[t] = (b * H + hh) * H + ww;
        if (pw == 0) {
          const float ya = ey ? (ty < kWin - shift ? kLabelVal : 0.f) : kLabelVal, yb = ey ? (ty < kWin - shift ? 0.f : kLabelVal) : 0.f;
          const float xa = ex ? (tx < kWin - shift ? kLabelVal : 0.f) : kLabelVal, xb = ex ? (tx < kWin - shift ? 0.f : kLabelVal) : 0.f;
          uint2 v;
          v.x = Mma16<T>::pack(ya, yb);
          v.y = Mma16<T>::pack(xa, xb);
          *reinterpret_cast<uint2*>(meta + kMetaLab + t * kLabP * 2) = v;
        }
      }
      if (pw == 0 && lane == 0) {
        int* fl = reinterpret_cast<int*>(meta + kMetaFlag);
        fl[0] = (ey || ex) ? 1 : 0; fl[1] = ey ? 1 : 0; fl[2] = ex ? 1 : 0; fl[3] = 2 * g;
      }
      __syncwarp();
      const uint32_t data = sbase + s * kStageData;
      const int c = lane & 7, r_in = lane >> 3;
      const T* src0 = qkv + g * 64 + c * 8;
#pragma unroll
      for (int k = 0; k < 36 / kProducers; ++k) {            // rows 4 (4 k + pw) + r_in of the three tiles
        const int r = 4 * (kProducers * k + pw) + r_in;
        const T* src = src0 + (long)tok[r] * 3 * C;
        const uint32_t dst = data + r * kRowBytes + ((uint32_t)(c ^ (r & 7)) << 4);
        cp_async16(dst, src);
        cp_async16(dst + kTileBytes, src + C);
        cp_async16(dst + 2 * kTileBytes, src + 2 * C);
      }
      const float* bsrc = bias_tc + (long)(2 * g) * kBiasP;
      const uint32_t bdst = smem_u32(meta + kMetaBias);
      for (int i = pw * 32 + lane; i < 2 * kBiasP / 4; i += 32 * kProducers) cp_async16(bdst + i * 16, bsrc + i * 4);
      cp_async_commit();
      // Two items of loads in flight: the stage of item n-1 is announced once ITS copies have landed, after item n's have
      // been issued.  (Waiting for every item's copies before touching the next one put a full DRAM round trip into
      // each item: 69 of the 96 us of a stage-3 launch were this skeleton, profiles/r2_wattn_tc_decomposition.txt.)
      if (n > 0) {
        cp_async_wait<1>();
        fence_async_smem();                   // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(bar_full((n - 1) % kStages));
      }
    }
    if (n_mine > 0) {
      cp_async_wait<0>();
      fence_async_smem();
      mbar_arrive(bar_full((n_mine - 1) % kStages));
    }
  } else if (warp == kIssuerWarp) {
    // ================================================================ MMA issuer (one thread)
    if (lane == 0) {
      constexpr int fmt = std::is_same<T, __half>::value ? 0 : 1;
      constexpr uint32_t idesc_qk = make_idesc(128, kWinTok, fmt, 0);
      constexpr uint32_t idesc_pv = make_idesc(128, kHeadDim, fmt, 1);
      // S_h = Q_h K_h^T of item m (its stage must have landed)
      auto issue_qk = [&](int m, int h) {
        const uint32_t q_base = sbase + (m % kStages) * kStageData, k_base = q_base + kTileBytes;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)           // +32 bytes per k-step, +64 bytes for the second head, inside the swizzle atom
          umma_ss(tmem_base + kColS + h * kWinTok, wtc_desc_k(q_base) + (uint64_t)(h * 4 + ks * 2), wtc_desc_k(k_base) + (uint64_t)(h * 4 + ks * 2),
                  idesc_qk, (uint32_t)(ks != 0));
        umma_commit(bar_sfull(h));
      };
      if (n_mine > 0) {
        mbar_wait(bar_full(0), 0u);
        tc_fence_after();
        issue_qk(0, 0);
        // Head B starts late (once head A's warps have finished their first pass 1) and stays out of phase, so that the
        // two softmax warps of a scheduler are rarely in the MUFU-bound exp pass at the same time.
        mbar_wait(bar_sfree(0), 0u);
        issue_qk(0, 1);
      }
      for (int n = 0; n < n_mine; ++n) {
        const int s = n % kStages;
        const uint32_t v_base = sbase + s * kStageData + 2 * kTileBytes;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(bar_pfull(h), (uint32_t)(n & 1));               // P_h(n) is in TMEM, S_h(n) has been consumed
          tc_fence_after();
          if (n + 1 < n_mine) {                                     // the next item's scores first: the softmax warps wait for them
            if (h == 0) { mbar_wait(bar_full((n + 1) % kStages), (uint32_t)(((n + 1) / kStages) & 1)); tc_fence_after(); }
            issue_qk(n + 1, h);
          }
          mbar_wait(bar_ofree(h), (uint32_t)((n & 1) ^ 1));         // O_h of the previous item has been read
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 9; ++ks)           // 16 keys per k-step: 8 TMEM columns of packed pairs, 16 token rows of V
            umma_ts(tmem_base + kColO + h * kHeadDim, tmem_base + kColP + h * (kWinTok / 2) + ks * 8,
                    make_sw128_desc(v_base + ks * 2048) + (uint64_t)(h * 4), idesc_pv, (uint32_t)(ks != 0));
          umma_commit(bar_ofull(h));
        }
        umma_commit(bar_empty(s));                 // every MMA that reads this stage has retired
      }
    }
  } else if (warp == kTailWarp0 || warp == kTailWarp0 + 1) {
    // ================================================================ rows 128..143 of head hsel on mma.sync
    const int hsel = warp - kTailWarp0;
    const int r0 = 128 + (lane >> 2), r1 = r0 + 8;
    const int rowoff0 = 4 * ((r0 / kWin) * kBiasRow + (r0 % kWin) + (kWin - 1) * kBiasRow + (kWin - 1));
    const int rowoff1 = 4 * ((r1 / kWin) * kBiasRow + (r1 % kWin) + (kWin - 1) * kBiasRow + (kWin - 1));
    const float sl = kScale * kLog2e;
    const uint32_t ones = Mma16<T>::pack(1.0f, 1.0f);
    const int l7 = lane & 7;
    for (int n = 0; n < n_mine; ++n) {
      const int s = n % kStages;
      mbar_wait(bar_full(s), (uint32_t)((n / kStages) & 1));
      const unsigned char* meta = smem + kSmemMeta + s * kMetaBytes;
      const int* tok = reinterpret_cast<const int*>(meta + kMetaTok);
      const int* fl = reinterpret_cast<const int*>(meta + kMetaFlag);
      const bool masked = fl[0] != 0;
      const int head = fl[3] + hsel;
      const uint32_t q_base = sbase + s * kStageData, k_base = q_base + kTileBytes, v_base = k_base + kTileBytes;
      const unsigned char* bt = meta + kMetaBias + hsel * kBiasP * 4;
      const unsigned char* bt0 = bt + rowoff0;
      const unsigned char* bt1 = bt + rowoff1;
      if (dbg & 1) { __syncwarp(); if (lane == 0) mbar_arrive(bar_empty(s)); continue; }     // timing experiment: no tail work
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
  } else if (warp < kMainWarps) {
    // ================================================================ softmax, thread per row (TMEM lane = query row)
    // Software pipeline over this CTA's items: pass 1 of item n, then the output of item n-1 (its P.V MMA ran meanwhile),
    // then pass 2 of item n; the issuer puts the scores of item n+1 into the S columns as soon as pass 2 has consumed them.
    // The row is streamed through registers in slices of 32 columns (the last one 16), the next slice's tcgen05.ld in
    // flight during the current slice's arithmetic.
    const int hsel = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int yi = row / kWin, xi = row % kWin;
    const int rowc4 = 4 * ((yi + kWin - 1) * kBiasRow + xi + kWin - 1);
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t tS = t_lane + kColS + hsel * kWinTok, tP = t_lane + kColP + hsel * (kWinTok / 2), tO = t_lane + kColO + hsel * kHeadDim;
    const float c2 = kScale * kLog2e;
    float sum_prev = 1.f;
    long out_prev = 0;
    // O = P V of an earlier item from TMEM, normalise, 64 contiguous bytes per row
    auto write_out = [&](int m, float sum, long off) {
      mbar_wait(bar_ofull(hsel), (uint32_t)(m & 1));
      tc_fence_after();
      uint32_t ov[32];
      tmem_ld32_nc(tO, ov);
      tmem_ld_wait32(ov);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ofree(hsel));
      const float inv = 1.0f / sum;
      uint4* dst = reinterpret_cast<uint4*>(out + off);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 w;
        w.x = Mma16<T>::pack(__uint_as_float(ov[8 * i + 0]) * inv, __uint_as_float(ov[8 * i + 1]) * inv);
        w.y = Mma16<T>::pack(__uint_as_float(ov[8 * i + 2]) * inv, __uint_as_float(ov[8 * i + 3]) * inv);
        w.z = Mma16<T>::pack(__uint_as_float(ov[8 * i + 4]) * inv, __uint_as_float(ov[8 * i + 5]) * inv);
        w.w = Mma16<T>::pack(__uint_as_float(ov[8 * i + 6]) * inv, __uint_as_float(ov[8 * i + 7]) * inv);
        dst[i] = w;
      }
    };
    for (int n = 0; n < n_mine; ++n) {
      const int s = n % kStages;
      mbar_wait(bar_full(s), (uint32_t)((n / kStages) & 1));
      const unsigned char* meta = smem + kSmemMeta + s * kMetaBytes;
      const int tokrow = reinterpret_cast<const int*>(meta + kMetaTok)[row];
      const int* fl = reinterpret_cast<const int*>(meta + kMetaFlag);
      const bool masked = fl[0] != 0;
      const int head = fl[3] + hsel;
      const unsigned char* bt = meta + kMetaBias + hsel * kBiasP * 4 + rowc4;
      mbar_wait(bar_sfull(hsel), (uint32_t)(n & 1));
      tc_fence_after();
      // ---- pass 1: s + bias (+ mask) written back in place, row max (four partial maxima: no serial chain)
      float mxs[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      uint32_t v[2][32];
      auto pass1 = [&](auto masked_c) {
        constexpr bool kMasked = decltype(masked_c)::value;
        // shift mask: -576 (raw score units; x scale = -101.8) per axis on which the key's region differs from the row's
        float mq[4] = {0.f, 0.f, 0.f, 0.f};
        if (kMasked) {
          const bool ey = fl[1] != 0, ex = fl[2] != 0;
          const bool ry = yi >= kWin - shift, rx = xi >= kWin - shift;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool ky = (q >> 1) != 0, kx = (q & 1) != 0;
            mq[q] = -kMaskStep * (float)((ey && ky != ry) + (ex && kx != rx));
          }
        }
        tmem_ld32_nc(tS, v[0]);
        static_for<5>([&](auto c_c) {
          constexpr int c = decltype(c_c)::value;
          constexpr int w = c < 4 ? 32 : 16;                       // columns of this slice
          if (w == 32) tmem_ld_wait32(v[c & 1]);
          else tmem_ld_wait16(v[c & 1]);
          if (c + 1 < 4) tmem_ld32_nc(tS + (c + 1) * 32, v[(c + 1) & 1]);
          else if (c + 1 == 4) tmem_ld16_nc(tS + 128, v[0]);
#pragma unroll
          for (int e = 0; e < w; ++e) {
            const int j = c * 32 + e, yj = j / kWin, xj = j % kWin;              // compile-time after unrolling
            float f = __uint_as_float(v[c & 1][e]) + *reinterpret_cast<const float*>(bt - 4 * (yj * kBiasRow + xj));
            if (kMasked) f += mq[(yj >= kWin / 2 ? 2 : 0) + (xj >= kWin / 2 ? 1 : 0)];
            mxs[e & 3] = fmaxf(mxs[e & 3], f);
            v[c & 1][e] = __float_as_uint(f);
          }
          if (w == 32) tmem_st32_nc(tS + c * 32, v[c & 1]);
          else tmem_st16_nc(tS + c * 32, v[c & 1]);
        });
      };
      if (!(dbg & 2)) {
        if (masked) pass1(std::true_type{});
        else pass1(std::false_type{});
      }
      const float mx = fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3]));
      tmem_st_wait();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_empty(s));                       // bias table / token rows of this stage are no longer read
        if (n == 0) mbar_arrive(bar_sfree(hsel));        // (one-time start signal for the issuer's stagger)
      }
      // ---- the previous item's output: its P.V MMA ran during pass 1
      if (n > 0) write_out(n - 1, sum_prev, out_prev);
      // ---- pass 2: p = 2^((s - max) * scale * log2 e), packed 16-bit pairs into the P columns
      const float negm = -mx * c2;
      float sums[2] = {0.f, 0.f};
      if (!(dbg & 4)) {
        tmem_ld32_nc(tS, v[0]);
        static_for<5>([&](auto c_c) {
          constexpr int c = decltype(c_c)::value;
          constexpr int w = c < 4 ? 32 : 16;
          if (w == 32) tmem_ld_wait32(v[c & 1]);
          else tmem_ld_wait16(v[c & 1]);
          if (c + 1 < 4) tmem_ld32_nc(tS + (c + 1) * 32, v[(c + 1) & 1]);
          else if (c + 1 == 4) tmem_ld16_nc(tS + 128, v[0]);
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < w / 2; ++e) {
            const float p0 = ex2_ftz(fmaf(__uint_as_float(v[c & 1][2 * e]), c2, negm));
            const float p1 = ex2_ftz(fmaf(__uint_as_float(v[c & 1][2 * e + 1]), c2, negm));
            pk[e] = Mma16<T>::pack(p0, p1);
            if (std::is_same<T, __half>::value) sums[e & 1] += p0 + p1;
            else      // bf16 keeps 8 significand bits: normalise by the sum of the ROUNDED probabilities (weights sum to 1 exactly)
              sums[e & 1] += __uint_as_float(pk[e] << 16) + __uint_as_float(pk[e] & 0xffff0000u);
          }
          if (w == 32) tmem_st16_nc(tP + c * 16, pk);
          else tmem_st8_nc(tP + c * 16, pk);
        });
      }
      const float sum = sums[0] + sums[1] + ((dbg & 4) ? 1.f : 0.f);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pfull(hsel));
      sum_prev = sum;
      out_prev = (long)tokrow * C + head * kHeadDim;
    }
    if (n_mine > 0) write_out(n_mine - 1, sum_prev, out_prev);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kIssuerWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

int g_attn_tc = 1;
int g_attn_tc_dbg = 0;      // timing experiments only (bit 0: no tail rows, 1: no pass 1, 2: no pass 2): wrong results

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
cudaError_t launch_window_attention_tc(const T* qkv, const float* bias_derived, T* out, int B, int H, int C, int heads, int shift,
                                       cudaStream_t st) {
  if (!window_attention_tc_supported(B, H, C, heads, shift)) return cudaErrorInvalidValue;
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(window_attention_tc_kernel<T>, wtc::kSmemTotal, smem_state)) return e;
  const int nW = (H / kWin) * (H / kWin);
  const int n_items = B * nW * (heads / 2);
  const int grid = std::min(n_items, wtc_sm_count());
  return launch_k(window_attention_tc_kernel<T>, dim3(grid), dim3(wtc::kThreads), wtc::kSmemTotal, st, qkv, bias_derived + (size_t)2 * 532 * heads, out, H, C, heads, shift,
                  n_items, g_attn_tc_dbg);
}
template cudaError_t launch_window_attention_tc<bf16>(const bf16*, const float*, bf16*, int, int, int, int, int, cudaStream_t);
template cudaError_t launch_window_attention_tc<__half>(const __half*, const float*, __half*, int, int, int, int, int, cudaStream_t);

}  // namespace xn
