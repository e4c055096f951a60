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
#include "beam_rows.cuh"

namespace xn {

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

__global__ void beam_first_kernel(BeamBufs bb, const float* __restrict__ top_val, const int* __restrict__ top_idx,
                                  int B, int beam, int L, int eos) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B * beam) return;
  beam_first_row(bb, top_val, top_idx, beam, L, eos, r);
}

__global__ void __launch_bounds__(128) beam_step_kernel(BeamBufs bb, int src, const float* __restrict__ top_val,
                                                        const int* __restrict__ top_idx, int B, int beam, int L, int t, int eos) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  beam_step_image(bb, src, top_val, top_idx, beam, L, t, eos, b, lane);
}

__global__ void beam_finalize_kernel(BeamBufs bb, int src, int B, int beam, int L, int t_final, int how_many,
                                     int* __restrict__ out_tokens, int* __restrict__ out_len,
                                     float* __restrict__ out_lp) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  beam_finalize_image(bb, beam, L, t_final, how_many, out_tokens, out_len, out_lp, b);
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
