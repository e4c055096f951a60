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
#include <algorithm>
#include <type_traits>
#include "kernels.h"
#include "common.cuh"
#include "mma16_frag.cuh"

namespace xn {

constexpr int kMmaThreads = 288;
constexpr int kRowPad = 40;                      // 16-bit elements per smem row (80 B): conflict-free ldmatrix
constexpr int kBiasN = (2 * kWin - 1) * (2 * kWin - 1);   // 529

constexpr int kBiasPitch = 532;                  // floats per head in the derived table (16-byte multiple)
constexpr int kLPitch = 24;                      // 16-bit elements per row of the label tile (48 B: conflict-free ldmatrix)
constexpr float kLabelVal = 24.0f;               // label one-hot magnitude: 24*24 = 576 -> mask = -576 * scale = -101.8 per mismatching axis
constexpr float kLog2e = 1.4426950408889634f;

// bias_l2 is the relative-position table transposed to (heads, 532) and divided by the score scale (derived once per
// weight load): the score accumulators are initialised with it, so s = q.k + bias/scale comes straight out of the MMAs
// and the softmax is p = 2^((s - max) * scale * log2 e)  (ex2.f16x2 was tried: it issues two MUFU ops plus a PRMT, no gain).
//
// Persistent CTAs loop over items = (window, group of kHeadsPerCta heads), head group fastest so that CTAs running
// side by side share the 128-byte lines of a token's QKV row.  Q/K/V of the next head -- or of the next item's first
// head -- stream into the other smem buffer with cp.async while the current head is computed; the per-item metadata
// (token rows, label tile, bias tables) is double-buffered by item parity.
//
// The shift mask costs no ALU work: every token carries a 4-wide "label" row  24 * [y<6, y>=6, x<6, x>=6]  (only on
// the axes where the window straddles the cyclic seam; [1,0] otherwise), the score accumulators start at bias - 2*576 (second copy of the table) and
// one extra k16 MMA step adds 576 per matching axis: same region -> 0, different region -> <= -576 (x scale = -101.8,
// where the reference adds -100; both vanish in the softmax).  The row sum comes out of the P.V MMA through an
// all-ones B fragment, i.e. it is the sum of the rounded probabilities that are actually multiplied with V.
template <typename T, int kHeadsPerCta>
__global__ void __launch_bounds__(kMmaThreads, 2) window_attention_mma_kernel(const T* __restrict__ qkv,
                                                                              const float* __restrict__ bias_l2,
                                                                              T* __restrict__ out, int H, int C, int heads,
                                                                              int shift, int n_items) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: stage[2] x {Q,K,V} x [144][40] 16-bit | bias[2][kHeadsPerCta][532] f32 | label[2][144][24] 16-bit | tok[2][144] | jneg[144]
  constexpr int kMatElems = kWinTok * kRowPad;
  T* stage = reinterpret_cast<T*>(smem_raw);
  float* bt_all = reinterpret_cast<float*>(smem_raw + 2 * 3 * kMatElems * sizeof(T));
  T* lab_all = reinterpret_cast<T*>(bt_all + 2 * kHeadsPerCta * kBiasPitch);
  int* tok_all = reinterpret_cast<int*>(lab_all + 2 * kWinTok * kLPitch);
  int* jneg = tok_all + 2 * kWinTok;             // byte offset -4 * (yj*23 + xj) per key token

  const int nWs = H / kWin, nW = nWs * nWs, ngroups = heads / kHeadsPerCta;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int my_ty = tid / kWin, my_tx = tid % kWin;          // token handled in the metadata phase (tid < 144)
  const uint32_t stage_u32 = (uint32_t)__cvta_generic_to_shared(stage);
  const uint32_t lab_u32 = (uint32_t)__cvta_generic_to_shared(lab_all);
  const uint32_t bt_u32 = (uint32_t)__cvta_generic_to_shared(bt_all);

  // item -> (head group g, batch b, window wy/wx): decoded once per item (runtime divisions)
  struct ItemPos { int g, b, wy, wx; bool masked; };
  auto decode_item = [&](int item) {
    ItemPos p;
    p.g = item % ngroups;
    const int wid = item / ngroups;
    p.b = wid / nW;
    const int wrem = wid - p.b * nW;
    p.wy = wrem / nWs;
    p.wx = wrem - p.wy * nWs;
    p.masked = shift > 0 && (p.wy == nWs - 1 || p.wx == nWs - 1);
    return p;
  };
  // per-item metadata (token rows, label rows, bias tables) into buffer mb
  auto compute_meta = [&](const ItemPos& ip, int mb) {
    if (tid < kWinTok) {
      int h = ip.wy * kWin + my_ty + shift, w = ip.wx * kWin + my_tx + shift;   // roll(-shift): shifted[h] = x[h + shift]
      if (h >= H) h -= H;
      if (w >= H) w -= H;
      tok_all[mb * kWinTok + tid] = (ip.b * H + h) * H + w;
      const bool ey = shift > 0 && ip.wy == nWs - 1, ex = shift > 0 && ip.wx == nWs - 1;
      const float ya = ey ? (my_ty < kWin - shift ? kLabelVal : 0.f) : kLabelVal, yb = ey ? (my_ty < kWin - shift ? 0.f : kLabelVal) : 0.f;
      const float xa = ex ? (my_tx < kWin - shift ? kLabelVal : 0.f) : kLabelVal, xb = ex ? (my_tx < kWin - shift ? 0.f : kLabelVal) : 0.f;
      uint2 v;
      v.x = Mma16<T>::pack(ya, yb);
      v.y = Mma16<T>::pack(xa, xb);
      *reinterpret_cast<uint2*>(lab_all + (mb * kWinTok + tid) * kLPitch) = v;
    }
    // masked windows take the second copy of the table, which already carries the -2 * 576 accumulator offset
    const float* src = bias_l2 + ((long)(ip.masked ? heads : 0) + (long)ip.g * kHeadsPerCta) * kBiasPitch;
    for (int i = tid; i < kHeadsPerCta * (kBiasPitch / 4); i += kMmaThreads)
      cp_async16(bt_u32 + (uint32_t)((mb * kHeadsPerCta * kBiasPitch + i * 4) * sizeof(float)), src + i * 4);
  };
  // this thread's six (token, q/k/v, 16-byte chunk) slots of a head: slot i = tid + 288 k  ->  token tid/12 + 24 k and
  // the same (q/k/v, chunk) = tid % 12 for every k, so two registers describe all six
  const int ld_t0 = tid / 12, ld_r = tid % 12;
  const int ld_col = (ld_r >> 2) * C + (ld_r & 3) * 8;
  const uint32_t ld_dst0 = (uint32_t)(((ld_r >> 2) * kMatElems + ld_t0 * kRowPad + (ld_r & 3) * 8) * sizeof(T));
  auto issue_loads = [&](int head, int mb, int bufi) {
    const uint32_t base = stage_u32 + (uint32_t)(bufi * 3 * kMatElems * sizeof(T)) + ld_dst0;
    const T* src = qkv + head * kHeadDim + ld_col;
    const int* tk = tok_all + mb * kWinTok + ld_t0;
#pragma unroll
    for (int k = 0; k < 6; ++k)
      cp_async16(base + (uint32_t)(k * 24 * kRowPad * sizeof(T)), src + (long)tk[k * 24] * 3 * C);
    cp_async_commit();
  };

  // one-time: zero the label tiles (columns 4..23 stay zero for ever) and the key offsets
  for (int i = tid; i < 2 * kWinTok * kLPitch / 2; i += kMmaThreads) reinterpret_cast<uint32_t*>(lab_all)[i] = 0u;
  if (tid < kWinTok) jneg[tid] = -4 * (my_ty * (2 * kWin - 1) + my_tx);
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  int item = blockIdx.x;
  if (item >= n_items) return;
  int cur_g;
  bool cur_masked;
  {
    const ItemPos first = decode_item(item);
    cur_g = first.g;
    cur_masked = first.masked;
    compute_meta(first, 0);
  }
  __syncthreads();
  issue_loads(cur_g * kHeadsPerCta, 0, 0);

  const int m0 = warp * 16;
  const int r0 = m0 + (lane >> 2), r1 = r0 + 8;
  const float sl = 0.17677669529663687f * kLog2e;      // 32^-0.5 * log2(e)
  // byte offset of (row token) + centre of the 23 x 23 table; the key offset jneg is added per column
  const int rowoff0 = 4 * ((r0 / kWin) * (2 * kWin - 1) + (r0 % kWin) + (kWin - 1) * (2 * kWin - 1) + (kWin - 1));
  const int rowoff1 = 4 * ((r1 / kWin) * (2 * kWin - 1) + (r1 % kWin) + (kWin - 1) * (2 * kWin - 1) + (kWin - 1));
  const uint32_t ones = Mma16<T>::pack(1.0f, 1.0f);

  int unit = 0;
  for (int k = 0; item < n_items; item += gridDim.x, ++k) {
    const int mb = k & 1;
    const int next_item = item + gridDim.x;
    const bool has_next = next_item < n_items;
    const bool masked = cur_masked;
    const int* tok = tok_all + mb * kWinTok;
    const int head0 = cur_g * kHeadsPerCta;
    int nxt_g = 0;
    bool nxt_masked = false;
#pragma unroll 1
    for (int hh = 0; hh < kHeadsPerCta; ++hh, ++unit) {
      const int bufi = unit & 1;
      cp_async_wait<0>();
      __syncthreads();                              // this unit landed; everyone is done with the other buffer
      if (hh == 0 && has_next) {
        const ItemPos nxt = decode_item(next_item);  // only the head group and the mask flag stay live across the heads
        nxt_g = nxt.g;
        nxt_masked = nxt.masked;
        compute_meta(nxt, mb ^ 1);                  // its cp.async traffic joins the next commit group
        if (kHeadsPerCta == 1) __syncthreads();
      }
      if (hh + 1 < kHeadsPerCta) issue_loads(head0 + hh + 1, mb, bufi ^ 1);
      else if (has_next) issue_loads(nxt_g * kHeadsPerCta, mb ^ 1, bufi ^ 1);

      const uint32_t q_base = stage_u32 + (uint32_t)(bufi * 3 * kMatElems * sizeof(T));
      const uint32_t k_base = q_base + (uint32_t)(kMatElems * sizeof(T));
      const uint32_t v_base = k_base + (uint32_t)(kMatElems * sizeof(T));
      T* Qs = stage + bufi * 3 * kMatElems;

      // The accumulators start at bias / scale (the derived table), so after the MMAs s = q.k + bias/scale and
      // p = 2^((s - max s) * scale * log2 e): no separate bias pass, no zero fill.
      // bias index = (yi - yj + 11)*23 + (xi - xj + 11) = rowoff_i + jneg_j
      const unsigned char* bt = reinterpret_cast<const unsigned char*>(bt_all + (mb * kHeadsPerCta + hh) * kBiasPitch);
      const unsigned char* bt0 = bt + rowoff0;
      const unsigned char* bt1 = bt + rowoff1;
      const int* jq = jneg + (lane & 3) * 2;
      float s[18][4];
#pragma unroll
      for (int nt = 0; nt < 18; ++nt) {
        const int jo = jq[nt * 8];                             // key 2q; key 2q+1 sits 4 bytes lower (same window row)
        s[nt][0] = *reinterpret_cast<const float*>(bt0 + jo);
        s[nt][1] = *reinterpret_cast<const float*>(bt0 + jo - 4);
        s[nt][2] = *reinterpret_cast<const float*>(bt1 + jo);
        s[nt][3] = *reinterpret_cast<const float*>(bt1 + jo - 4);
      }
      {
        const int row = m0 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = (lane >> 4) * 8;
        if (masked) {
          const uint32_t l_base = lab_u32 + (uint32_t)(mb * kWinTok * kLPitch * sizeof(T));
          uint32_t la[4];
          ldsm_x4(la, l_base + (row * kLPitch + col) * 2);
          const uint32_t lb_base = l_base + (uint32_t)(((lane & 7) * kLPitch + ((lane >> 3) & 1) * 8) * 2);
          static_for<18>([&](auto nt_c) {
            constexpr int nt = decltype(nt_c)::value;
            uint32_t lb[2];
            ldsm_x2_o<nt * 8 * kLPitch * 2>(lb, lb_base);
            Mma16<T>::mma(s[nt], la, lb[0], lb[1]);
          });
        }
        // A fragments of Q for the two k16 steps (d 0-15, 16-31)
        uint32_t qa[2][4];
        const uint32_t qaddr = q_base + (uint32_t)((row * kRowPad + col) * 2);
        ldsm_x4_o<0>(qa[0], qaddr);
        ldsm_x4_o<32>(qa[1], qaddr);
        // two passes over the 18 key tiles (d 0-15, then d 16-31): the two MMAs that accumulate into the same score tile
        // are 18 instructions apart instead of back to back, so neither waits for the other's result
        const uint32_t kaddr = k_base + (uint32_t)(((lane & 7) * kRowPad + ((lane >> 3) & 1) * 8) * 2);
        static_for<18>([&](auto nt_c) {
          constexpr int nt = decltype(nt_c)::value;
          uint32_t kb[2];
          ldsm_x2_o<nt * 8 * kRowPad * 2>(kb, kaddr);
          Mma16<T>::mma(s[nt], qa[0], kb[0], kb[1]);
        });
        static_for<18>([&](auto nt_c) {
          constexpr int nt = decltype(nt_c)::value;
          uint32_t kb[2];
          ldsm_x2_o<nt * 8 * kRowPad * 2 + 32>(kb, kaddr);
          Mma16<T>::mma(s[nt], qa[1], kb[0], kb[1]);
        });
      }

      // row softmax (rows r0, r1 = r0 + 8) straight into packed 16-bit probabilities
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 18; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float nm0 = -mx0 * sl, nm1 = -mx1 * sl;
      uint32_t pa[18][2];   // packed probabilities: [nt][0] = row r0 (cols 2q,2q+1), [nt][1] = row r1
#pragma unroll
      for (int nt = 0; nt < 18; ++nt) {
        pa[nt][0] = Mma16<T>::pack(ex2_ftz(fmaf(s[nt][0], sl, nm0)), ex2_ftz(fmaf(s[nt][1], sl, nm0)));
        pa[nt][1] = Mma16<T>::pack(ex2_ftz(fmaf(s[nt][2], sl, nm1)), ex2_ftz(fmaf(s[nt][3], sl, nm1)));
      }

      // O = P V : 9 k16 steps over the keys, 4 n8 tiles over head_dim + one all-ones tile for the row sums
      float o[5][4];
#pragma unroll
      for (int n = 0; n < 5; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
      {
        const uint32_t vaddr = v_base + (uint32_t)(((((lane >> 3) & 1) * 8 + (lane & 7)) * kRowPad + (lane >> 4) * 8) * 2);
        static_for<9>([&](auto ks_c) {
          constexpr int ks = decltype(ks_c)::value;
          uint32_t a[4] = {pa[2 * ks][0], pa[2 * ks][1], pa[2 * ks + 1][0], pa[2 * ks + 1][1]};
          uint32_t vb0[4], vb1[4];                     // d 0-15 and d 16-31
          ldsm_x4_trans_o<ks * 16 * kRowPad * 2>(vb0, vaddr);
          ldsm_x4_trans_o<ks * 16 * kRowPad * 2 + 32>(vb1, vaddr);
          Mma16<T>::mma(o[0], a, vb0[0], vb0[1]);
          Mma16<T>::mma(o[1], a, vb0[2], vb0[3]);
          Mma16<T>::mma(o[2], a, vb1[0], vb1[1]);
          Mma16<T>::mma(o[3], a, vb1[2], vb1[3]);
          Mma16<T>::mma(o[4], a, ones, ones);
        });
      }
      const float inv0 = 1.0f / o[4][0], inv1 = 1.0f / o[4][2];

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
    cur_g = nxt_g;
    cur_masked = nxt_masked;
  }
}

static int wa_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T, int HPC>
static cudaError_t launch_wa_hpc(const T* qkv, const float* bias_l2, T* out, int B, int H, int C, int heads, int shift,
                                 cudaStream_t st) {
  const size_t smem = 2 * 3 * kWinTok * kRowPad * sizeof(T) + 2 * HPC * kBiasPitch * sizeof(float) +
                      2 * kWinTok * kLPitch * sizeof(T) + 3 * kWinTok * sizeof(int);
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(window_attention_mma_kernel<T, HPC>, smem, smem_state)) return e;
  const int nW = (H / kWin) * (H / kWin);
  const int n_items = B * nW * (heads / HPC);
  const int grid = std::min(n_items, 2 * wa_sm_count());
  launch_k(window_attention_mma_kernel<T, HPC>, dim3(grid), dim3(kMmaThreads), smem, st, qkv, bias_l2, out, H, C, heads, shift, n_items);
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch_window_attention_mma(const T* qkv, const float* bias_l2, T* out, int B, int H, int C, int heads,
                                        int shift, cudaStream_t st) {
  if (H % kWin || C != heads * kHeadDim || (shift != 0 && shift != kWin / 2)) return cudaErrorInvalidValue;
  const int nW = (H / kWin) * (H / kWin);
  // three heads per item amortise the per-window metadata; small problems (batch-1 latency) take one head per item
  // so that there are enough items to fill the machine
  const bool wide = (long)B * nW * (heads / 3) >= 2L * wa_sm_count();
  if (heads % 3 == 0 && wide) return launch_wa_hpc<T, 3>(qkv, bias_l2, out, B, H, C, heads, shift, st);
  if (heads % 2 == 0 && (long)B * nW * (heads / 2) >= 2L * wa_sm_count()) return launch_wa_hpc<T, 2>(qkv, bias_l2, out, B, H, C, heads, shift, st);
  return launch_wa_hpc<T, 1>(qkv, bias_l2, out, B, H, C, heads, shift, st);
}
template cudaError_t launch_window_attention_mma<bf16>(const bf16*, const float*, bf16*, int, int, int, int, int, cudaStream_t);
template cudaError_t launch_window_attention_mma<__half>(const __half*, const float*, __half*, int, int, int, int, int, cudaStream_t);

}  // namespace xn

namespace xn {
// (529, heads) relative-position table -> derived tables, once per weight load (kBiasDerivedFloats(heads) floats):
//   [0, 532 heads)            (heads, 532) table / scale (scale = 32^-0.5)                       -- mma.sync kernel
//   [532 heads, 1064 heads)   the same shifted by the mask offset the label MMA cancels           -- mma.sync kernel, seam windows
//   [1064 heads, ...)         (heads, 1024): entry (dy + 11) * 44 + (dx + 11), / scale            -- tcgen05 kernel (window_attn_tc.cu:
//                             the row pitch 44 = 12 mod 32 makes the softmax threads' loads bank-conflict free)
__global__ void transpose_bias_kernel(const float* __restrict__ t, float* __restrict__ o, int heads) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kBiasPitch * heads) {
    const int h = i / kBiasPitch, e = i % kBiasPitch;
    const float v = e < kBiasN ? t[e * heads + h] * 5.656854249492380f : 0.f;     // / 32^-0.5
    o[i] = v;
    o[i + kBiasPitch * heads] = v - 2.0f * kLabelVal * kLabelVal;                   // copy for windows on the cyclic seam
  }
  if (i < 1024 * heads) {
    const int h = i / 1024, e = i % 1024, dy = e / 44, dx = e % 44;
    o[2 * kBiasPitch * heads + i] = (dy < 2 * kWin - 1 && dx < 2 * kWin - 1) ? t[(dy * (2 * kWin - 1) + dx) * heads + h] * 5.656854249492380f : 0.f;
  }
}
cudaError_t launch_transpose_bias(const float* table, float* out, int heads, cudaStream_t st) {
  transpose_bias_kernel<<<(1024 * heads + 255) / 256, 256, 0, st>>>(table, out, heads);
  return cudaGetLastError();
}
}  // namespace xn
