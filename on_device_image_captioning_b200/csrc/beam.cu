// Beam-search bookkeeping on the device (reference legacy_models/captioning_model.py:111-241,
// 'max' branch == models/captioning_model.py:220-427).  The candidate sets are tiny
// (beam <= 8, beam^2 <= 64), so one thread owns one image and runs the reference's rules
// serially and deterministically: EOS override (0.0 / -999), beam^2 merge top-k (descending,
// ties to the lower flat index), parent/word split, history append, length update.  The
// expansion-state "gather by parent beam" is the `anc` table: per row and position, the slot
// holding that position's cached state; reordering beams only permutes these small tables.
#include "kernels.h"
#include "common.cuh"

namespace xn {

constexpr int kMaxBeam = 8;

__global__ void beam_init_kernel(BeamBufs bb, int B, int beam, int L, int sos) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;      // row = b*beam + k
  if (r >= B * beam) return;
  for (int s = 0; s < 2; ++s) {
    for (int i = 0; i < L; ++i) {
      bb.tokens[s][(long)r * L + i] = (i == 0) ? sos : 0;
      bb.lps[s][(long)r * L + i] = 0.f;
      bb.anc[s][(long)r * L + i] = r;
    }
    bb.len[s][r] = 1;
  }
  if (r == 0) *bb.all_done = 0;
}

// step 0 (:242-271): all beams of an image hold [SOS]; beam k takes the k-th best first word of row (b,0).
__global__ void beam_first_kernel(BeamBufs bb, const float* __restrict__ top_val, const int* __restrict__ top_idx,
                                  int B, int beam, int L) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B * beam) return;
  const int b = r / beam, k = r % beam;
  const long src = (long)(b * beam) * beam + k;             // row (b,0), candidate k
  bb.tokens[0][(long)r * L + 1] = top_idx[src];
  bb.lps[0][(long)r * L + 1] = top_val[src];
  bb.len[0][r] = 2;
  bb.anc[0][(long)r * L + 0] = r;                            // every slot computed identical position-0 state
  bb.anc[0][(long)r * L + 1] = r;
}

// One loop iteration for time_step t (tokens 0..t-1 known, choosing token t)  (:295-397)
__global__ void beam_step_kernel(BeamBufs bb, int src, const float* __restrict__ top_val,
                                 const int* __restrict__ top_idx, int B, int beam, int L, int t, int eos) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int dst = src ^ 1;
  const int* tk = bb.tokens[src] + (long)b * beam * L;
  const float* lp = bb.lps[src] + (long)b * beam * L;
  const int* ln = bb.len[src] + b * beam;
  const int* an = bb.anc[src] + (long)b * beam * L;
  float cand[kMaxBeam * kMaxBeam];
  float word_lp[kMaxBeam * kMaxBeam];
  bool has_eos[kMaxBeam];
  for (int k = 0; k < beam; ++k) {
    bool e = false;
    float cum = 0.f;
    for (int i = 0; i < t; ++i) {
      e = e || (tk[k * L + i] == eos);
      cum += lp[k * L + i];                                  // cumul = history.sum(-1)  (:381)
    }
    has_eos[k] = e;
    for (int w = 0; w < beam; ++w) {
      float v = top_val[((long)(b * beam + k)) * beam + w];
      if (e) v = (w == 0) ? 0.0f : -999.0f;                  // (:322-335)
      word_lp[k * beam + w] = v;
      cand[k * beam + w] = cum + v;
    }
  }
  int pick[kMaxBeam];
  unsigned long long used = 0ull;
  for (int j = 0; j < beam; ++j) {                           // top-k of beam^2, sorted
    int best = -1;
    float bv = 0.f;
    for (int c = 0; c < beam * beam; ++c) {
      if ((used >> c) & 1ull) continue;
      if (best < 0 || cand[c] > bv) { best = c; bv = cand[c]; }
    }
    used |= 1ull << best;
    pick[j] = best;
  }
  int* tko = bb.tokens[dst] + (long)b * beam * L;
  float* lpo = bb.lps[dst] + (long)b * beam * L;
  int* lno = bb.len[dst] + b * beam;
  int* ano = bb.anc[dst] + (long)b * beam * L;
  bool any_grew = false;
  for (int j = 0; j < beam; ++j) {
    const int parent = pick[j] / beam, w = pick[j] % beam;
    for (int i = 0; i < t; ++i) {
      tko[j * L + i] = tk[parent * L + i];
      lpo[j * L + i] = lp[parent * L + i];
      ano[j * L + i] = an[parent * L + i];
    }
    tko[j * L + t] = top_idx[((long)(b * beam + parent)) * beam + w];
    lpo[j * L + t] = word_lp[parent * beam + w];
    if (t < L) ano[j * L + t] = b * beam + j;               // the next step writes position t into slot j
    const int nl = ln[parent] + (has_eos[parent] ? 0 : 1);  // (:384-395)
    lno[j] = nl;
    any_grew = any_grew || (nl == t + 1);
  }
  if (any_grew) atomicExch(bb.all_done, 0);                  // informational; the host does not poll it
}

// (:401-425)  score = cumul / len, best `how_many` beams, tokens [:len], log-probs zero padded.
__global__ void beam_finalize_kernel(BeamBufs bb, int src, int B, int beam, int L, int t_final, int how_many,
                                     int* __restrict__ out_tokens, int* __restrict__ out_len,
                                     float* __restrict__ out_lp) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
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

cudaError_t launch_beam_init(const BeamBufs& bb, int B, int beam, int L, int sos, cudaStream_t st) {
  if (beam > kMaxBeam) return cudaErrorInvalidValue;
  beam_init_kernel<<<(B * beam + 127) / 128, 128, 0, st>>>(bb, B, beam, L, sos);
  return cudaGetLastError();
}
cudaError_t launch_beam_first(const BeamBufs& bb, const float* top_val, const int* top_idx, int B, int beam, int L,
                              cudaStream_t st) {
  beam_first_kernel<<<(B * beam + 127) / 128, 128, 0, st>>>(bb, top_val, top_idx, B, beam, L);
  return cudaGetLastError();
}
cudaError_t launch_beam_step(const BeamBufs& bb, int src, const float* top_val, const int* top_idx, int B, int beam,
                             int L, int t, int eos, cudaStream_t st) {
  beam_step_kernel<<<(B + 63) / 64, 64, 0, st>>>(bb, src, top_val, top_idx, B, beam, L, t, eos);
  return cudaGetLastError();
}
cudaError_t launch_beam_finalize(const BeamBufs& bb, int src, int B, int beam, int L, int t_final, int how_many,
                                 int* out_tokens, int* out_len, float* out_lp, cudaStream_t st) {
  beam_finalize_kernel<<<(B + 63) / 64, 64, 0, st>>>(bb, src, B, beam, L, t_final, how_many, out_tokens, out_len,
                                                     out_lp);
  return cudaGetLastError();
}

}  // namespace xn
