// Beam-search bookkeeping on the device (reference legacy_models/captioning_model.py:111-241,
// 'max' branch == models/captioning_model.py:220-427).  The candidate sets are tiny
// (beam <= 8, beam^2 <= 64), so one thread owns one image and runs the reference's rules
// serially and deterministically: EOS override (0.0 / -999), beam^2 merge top-k (descending,
// ties to the lower flat index), parent/word split, history append, length update.  The
// expansion-state "gather by parent beam" is the `anc` table: per row and position, the slot
// holding that position's cached state; reordering beams only permutes these small tables.
#include <algorithm>
#include "kernels.h"
#include "common.cuh"

namespace xn {

constexpr int kMaxBeam = 8;

__global__ void beam_init_kernel(BeamBufs bb, int B, int beam, int L, int sos) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;      // row = b*beam + k
  if (r < L) bb.grew[r] = 0;
  if (r == 0) { *bb.all_done = 0; *bb.final_src = 0; }
  if (r >= B * beam) return;
  for (int s = 0; s < 2; ++s) {
    for (int i = 0; i < L; ++i) {
      bb.tokens[s][(long)r * L + i] = (i == 0) ? sos : 0;
      bb.lps[s][(long)r * L + i] = 0.f;
      bb.anc[s][(long)r * L + i] = r;
    }
    bb.len[s][r] = 1;
    bb.cum[s][r] = 0.f;
    bb.eos[s][r] = 0;
  }
}

// step 0 (:242-271): all beams of an image hold [SOS]; beam k takes the k-th best first word of row (b,0).
__global__ void beam_first_kernel(BeamBufs bb, const float* __restrict__ top_val, const int* __restrict__ top_idx,
                                  int B, int beam, int L, int eos) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B * beam) return;
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
__global__ void __launch_bounds__(128) beam_step_kernel(BeamBufs bb, int src, const float* __restrict__ top_val,
                                                        const int* __restrict__ top_idx, int B, int beam, int L, int t, int eos) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
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
__global__ void beam_finalize_kernel(BeamBufs bb, int src, int B, int beam, int L, int t_final, int how_many,
                                     int* __restrict__ out_tokens, int* __restrict__ out_len,
                                     float* __restrict__ out_lp) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  src = *bb.final_src;                    // the state after the last executed step (the host's `src` assumes none was skipped)
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

// ---- mode='sampling' (legacy_models/captioning_model.py:60-109): every row is an independent sample path.  bb.*[0]
// holds the histories; len = where_is_eos + 1 (position of the first sampled EOS, inclusive) once EOS has been drawn,
// otherwise the number of tokens so far.  The reference keeps sampling finished rows and cuts afterwards (:96-107); so
// does this kernel -- the tokens after EOS are stored and ignored by the finaliser.
__global__ void sample_append_kernel(BeamBufs bb, const float* __restrict__ top_val, const int* __restrict__ top_idx, int R, int L,
                                     int t, int eos) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int tok = top_idx[r];
  bb.tokens[0][(long)r * L + t] = tok;
  bb.lps[0][(long)r * L + t] = top_val[r];
  bb.anc[0][(long)r * L + t] = r;
  if (!bb.eos[0][r]) {
    bb.len[0][r] = t + 1;
    if (tok == eos) bb.eos[0][r] = 1;
  }
}

// tokens [:where_is_eos + 1], log-probs zeroed after the first EOS (:99-107); -1 / 0 padded to L
__global__ void sample_finalize_kernel(BeamBufs bb, int R, int L, int t_final, int* __restrict__ out_tokens,
                                       int* __restrict__ out_len, float* __restrict__ out_lp) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int n = min(bb.len[0][r], t_final);
  out_len[r] = n;
  for (int i = 0; i < L; ++i) {
    out_tokens[(long)r * L + i] = i < n ? bb.tokens[0][(long)r * L + i] : -1;
    out_lp[(long)r * L + i] = i < n ? bb.lps[0][(long)r * L + i] : 0.f;
  }
}

cudaError_t launch_sample_append(const BeamBufs& bb, const float* top_val, const int* top_idx, int R, int L, int t, int eos,
                                 cudaStream_t st) {
  return launch_k(sample_append_kernel, dim3((R + 127) / 128), dim3(128), 0, st, bb, top_val, top_idx, R, L, t, eos);
}
cudaError_t launch_sample_finalize(const BeamBufs& bb, int R, int L, int t_final, int* out_tokens, int* out_len, float* out_lp,
                                   cudaStream_t st) {
  return launch_k(sample_finalize_kernel, dim3((R + 127) / 128), dim3(128), 0, st, bb, R, L, t_final, out_tokens, out_len, out_lp);
}

__global__ void beam_set_condition_kernel(cudaGraphConditionalHandle handle, const int* __restrict__ grew_t) {
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) cudaGraphSetConditional(handle, *grew_t != 0 ? 1u : 0u);
}
cudaError_t launch_beam_set_condition(unsigned long long cond_handle, const int* grew_t, cudaStream_t st) {
  return launch_k(beam_set_condition_kernel, dim3(1), dim3(32), 0, st, (cudaGraphConditionalHandle)cond_handle, grew_t);
}

cudaError_t launch_beam_init(const BeamBufs& bb, int B, int beam, int L, int sos, cudaStream_t st) {
  if (beam > kMaxBeam) return cudaErrorInvalidValue;
  launch_k(beam_init_kernel, dim3((std::max(B * beam, L) + 127) / 128), dim3(128), 0, st, bb, B, beam, L, sos);
  return cudaGetLastError();
}
cudaError_t launch_beam_first(const BeamBufs& bb, const float* top_val, const int* top_idx, int B, int beam, int L,
                              int eos, cudaStream_t st) {
  launch_k(beam_first_kernel, dim3((B * beam + 127) / 128), dim3(128), 0, st, bb, top_val, top_idx, B, beam, L, eos);
  return cudaGetLastError();
}
cudaError_t launch_beam_step(const BeamBufs& bb, int src, const float* top_val, const int* top_idx, int B, int beam,
                             int L, int t, int eos, cudaStream_t st) {
  launch_k(beam_step_kernel, dim3((B + 3) / 4), dim3(128), 0, st, bb, src, top_val, top_idx, B, beam, L, t, eos);
  return cudaGetLastError();
}
cudaError_t launch_beam_finalize(const BeamBufs& bb, int src, int B, int beam, int L, int t_final, int how_many,
                                 int* out_tokens, int* out_len, float* out_lp, cudaStream_t st) {
  launch_k(beam_finalize_kernel, dim3((B + 63) / 64), dim3(64), 0, st, bb, src, B, beam, L, t_final, how_many, out_tokens, out_len,
                                                     out_lp);
  return cudaGetLastError();
}

}  // namespace xn
