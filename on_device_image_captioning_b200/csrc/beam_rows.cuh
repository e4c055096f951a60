// Beam-search bookkeeping bodies shared by the stand-alone kernels (beam.cu) and the persistent whole-search kernel
// (decode_mega.cu).  Reference: legacy_models/captioning_model.py:111-241 ('max' branch) == models/captioning_model.py:220-427.
#pragma once
#include "kernels.h"
#include "common.cuh"

namespace xn {

constexpr int kMaxBeam = 8;

// step 0 (:242-271): all beams of an image hold [SOS]; beam k takes the k-th best first word of row (b,0).  One thread per row r.
__device__ __forceinline__ void beam_first_row(const BeamBufs& bb, const float* top_val, const int* top_idx, int beam, int L, int eos, int r) {
  const int b = r / beam, k = r % beam;
  const long src = (long)(b * beam) * beam + k;             // row (b,0), candidate k
  bb.tokens[0][(long)r * L + 1] = top_idx[src];
  bb.lps[0][(long)r * L + 1] = top_val[src];
  bb.len[0][r] = 2;
  bb.cum[0][r] = 0.f + top_val[src];                         // running history sum, same order as history.sum(-1)
  bb.eos[0][r] = top_idx[src] == eos;
  bb.anc[0][(long)r * L + 0] = r;                            // every slot computed identical position-0 state
  bb.anc[0][(long)r * L + 1] = r;
  if (r == 0) *bb.final_src = 0;
}

// One loop iteration for time_step t (tokens 0..t-1 known, choosing token t)  (:295-397).  One warp per image: the
// beam^2 candidates live in lanes (two per lane for beam > 5), the sorted top-k is `beam` rounds of a warp arg-max
// (descending, ties to the lower flat index, as torch.topk on the flattened (beam, beam) candidates), the histories are
// copied lane-parallel.  cum / eos are the running history sum and "prefix contains EOS" flag of every beam: the running
// sum performs exactly the additions of history.sum(-1) in the same order.
__device__ __forceinline__ void beam_step_image(const BeamBufs& bb, int src, const float* top_val, const int* top_idx, int beam,
                                                int L, int t, int eos, int b, int lane) {
  const int dst = src ^ 1, nc = beam * beam;
  const int* tk = bb.tokens[src] + (long)b * beam * L;
  const float* lp = bb.lps[src] + (long)b * beam * L;
  const int* an = bb.anc[src] + (long)b * beam * L;
  float cand[2], wlp[2];
  bool used[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = lane + 32 * h;
    cand[h] = -INFINITY; wlp[h] = 0.f; used[h] = c >= nc;
    if (c < nc) {
      const int k = c / beam, w = c % beam;
      float v = top_val[((long)(b * beam + k)) * beam + w];
      if (bb.eos[src][b * beam + k]) v = (w == 0) ? 0.0f : -999.0f;      // (:322-335)
      wlp[h] = v;
      cand[h] = bb.cum[src][b * beam + k] + v;                            // cumul = history.sum(-1)  (:381)
    }
  }
  int mypick = 0;
  for (int j = 0; j < beam; ++j) {                           // top-k of beam^2, sorted
    float bv = 0.f;
    int bi = 0x7fffffff;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (!used[h] && (bi == 0x7fffffff || cand[h] > bv)) { bv = cand[h]; bi = lane + 32 * h; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if ((bi & 31) == lane) used[bi >> 5] = true;
    if (lane == j) mypick = bi;
  }
  int* tko = bb.tokens[dst] + (long)b * beam * L;
  float* lpo = bb.lps[dst] + (long)b * beam * L;
  int* ano = bb.anc[dst] + (long)b * beam * L;
  for (int j = 0; j < beam; ++j) {
    const int pick = __shfl_sync(0xffffffffu, mypick, j);
    const int parent = pick / beam;
    const float w0 = __shfl_sync(0xffffffffu, wlp[0], pick & 31), w1 = __shfl_sync(0xffffffffu, wlp[1], pick & 31);
    const float c0 = __shfl_sync(0xffffffffu, cand[0], pick & 31), c1 = __shfl_sync(0xffffffffu, cand[1], pick & 31);
    for (int i = lane; i < t; i += 32) {
      tko[j * L + i] = tk[parent * L + i];
      lpo[j * L + i] = lp[parent * L + i];
      ano[j * L + i] = an[parent * L + i];
    }
    if (lane == 0) {
      const int w = pick % beam;
      const int tokn = top_idx[((long)(b * beam + parent)) * beam + w];
      const bool pe = bb.eos[src][b * beam + parent] != 0;
      tko[j * L + t] = tokn;
      lpo[j * L + t] = pick < 32 ? w0 : w1;
      if (t < L) ano[j * L + t] = b * beam + j;               // the next step writes position t into slot j
      const int nl = bb.len[src][b * beam + parent] + (pe ? 0 : 1);   // (:384-395)
      bb.len[dst][b * beam + j] = nl;
      bb.cum[dst][b * beam + j] = pick < 32 ? c0 : c1;       // = cum[parent] + word log-prob
      bb.eos[dst][b * beam + j] = pe || tokn == eos;
      if (nl == t + 1) bb.grew[t] = 1;                       // some beam is still growing: the search goes on (:397)
      if (b == 0 && j == 0) *bb.final_src = dst;
    }
  }
}

// (:401-425)  score = cumul / len, best `how_many` beams, tokens [:len], log-probs zero padded.
__device__ __forceinline__ void beam_finalize_image(const BeamBufs& bb, int beam, int L, int t_final, int how_many,
                                                    int* out_tokens, int* out_len, float* out_lp, int b) {
  const int src = *bb.final_src;                    // the state after the last executed step (the host's `src` assumes none was skipped)
  const int* tk = bb.tokens[src] + (long)b * beam * L;
  const float* lp = bb.lps[src] + (long)b * beam * L;
  const int* ln = bb.len[src] + b * beam;
  float score[kMaxBeam];
  for (int k = 0; k < beam; ++k) {
    float cum = 0.f;
    for (int i = 0; i < t_final; ++i) cum += lp[k * L + i];
    score[k] = cum / (float)ln[k];
  }
  unsigned used = 0u;
  for (int j = 0; j < how_many; ++j) {
    int best = -1;
    float bv = 0.f;
    for (int k = 0; k < beam; ++k) {
      if ((used >> k) & 1u) continue;
      if (best < 0 || score[k] > bv) { best = k; bv = score[k]; }
    }
    used |= 1u << best;
    const int n = ln[best];
    out_len[b * how_many + j] = n;
    for (int i = 0; i < L; ++i) {
      out_tokens[((long)b * how_many + j) * L + i] = (i < n) ? tk[best * L + i] : -1;
      out_lp[((long)b * how_many + j) * L + i] = (i < n) ? lp[best * L + i] : 0.f;
    }
  }
}

}  // namespace xn
