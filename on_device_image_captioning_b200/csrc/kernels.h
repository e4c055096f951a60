// Host-side launch prototypes of every xnv2_b200 kernel family.  Each launcher returns the
// CUDA error of the launch and counts as one "gpu launch" of this library.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <vector>

namespace xn {

extern int g_pdl_enabled;   // programmatic dependent launch for the kernels of the 16-bit path (common.cuh)

typedef __nv_bfloat16 bf16;
typedef __half f16;

// ---------------------------------------------------------------- GEMM (fp32 FFMA path)
struct GemmArgs {
  const float* A; long lda, sA;      // (M x K), K contiguous; sA = batch stride (elements)
  const float* W; long ldw, sW;      // (N x K) [w_kn=0] or (K x N) [w_kn=1]
  float* C; long ldc, sC;
  const float* bias;                 // [N] or nullptr
  const float* res; long ldr, sR;    // residual added after the activation, or nullptr
  int M, N, K, batch;
  float div;                         // != 0: accumulator is divided by this first
  int act;                           // 0 none, 1 GELU(erf), 2 ReLU
  int w_kn;
};
cudaError_t launch_gemm_f32(const GemmArgs& p, cudaStream_t st);

// ---------------------------------------------------------------- GEMM (bf16 tcgen05 path)
// C = act(A x W^T + bias) + res, A (M x K) and W (N x K) 16-bit (bf16, or fp16 when fp16 != 0),
// K-contiguous, fp32 accumulation in TMEM.  Output fp32 (Cf) or 16-bit (Cb, same format as the
// operands), exactly one non-null.
struct TcGemmArgs {
  const void* A; long lda;
  const void* W; long ldw;
  float* Cf; void* Cb; long ldc;
  int fp16;
  const float* bias;
  const float* res; long ldr;        // fp32 residual
  int M, N, K;
  float div;
  int act;
  int w_static;                      // W was written before the previous kernel started (model weights): its tiles may
                                     // be fetched ahead of the programmatic-dependent-launch wait
  // LayerNorm-on-load (decoder-step linears, K == 512, one tile per CTA): A is produced inside the kernel from these fp32
  // rows -- LayerNorm(gamma, beta) (or a plain conversion when ln_g == nullptr), rounded to the operand type and written
  // into shared memory in the UMMA layout by the epilogue warps before their epilogue role.  `A` is then unused.
  const float* a32; long lda32;
  const float *ln_g, *ln_b;
  // LayerNorm folded ALGEBRAICALLY around the GEMM (Swin blocks).  With W' = gamma (.) W centred along K
  // (W'_nk -= mean_k' W'_nk', so that sum_k W'_nk = 0):   LN(x) W^T + b = rstd (x W'^T) + (b + W beta)
  // -- the row mean drops out of the contraction.  The producer of the residual stream (proj / fc2 / patch-merging
  // reduction, fp32 output) also writes the raw rows rounded to 16 bits (x16_out) and adds their sum and sum of squares into
  // stats_out (M x 2 64-bit fixed-point integers -- order-independent, hence deterministic -- zeroed by the caller); the consumer (qkv / fc1) takes those raw rows as A, the folded weights as W,
  // b + W beta as `bias`, and scales every accumulator row by rstd from ln_stats (ln_k = LayerNorm width).
  float* stats_out; void* x16_out; long ldx16;
  float* stats_zero;                 // fp32-output launches: rows of this (other) statistics buffer are cleared by the first column tile
  const float* ln_stats; int ln_k;
  // Batched mode (batch > 0): `batch` independent problems of the same M x N x K, operands stacked with the element
  // pitches sA / sW (3-D tensor maps: the M / N / K edges are clipped per problem), outputs with sC, the fp32 residual
  // with sR (0 = shared).  Single-CTA tiles and the generic epilogue only; no LayerNorm hooks.
  // Split-K is the same mode: the problems are K-slices of one product (sA = sW = K / batch, K = the slice length) that
  // leave fp32 partial products at the pitch sC for launch_layernorm_sum.
  int batch; long sA, sW, sC, sR;
};
bool tc_gemm_ln_supported(int M, int N, int K);
// W (N x K) fp32, LayerNorm affine (gamma, beta: K) -> w16 = round16(gamma (.) W - row mean), bias_out_n = bias_n + sum_k W_nk beta_k
cudaError_t launch_fold_ln_weight(const float* w, const float* gamma, const float* beta, const float* bias, void* w16,
                                  float* bias_out, int N, int K, int fp16, cudaStream_t st);
cudaError_t launch_gemm_tc(const TcGemmArgs& p, cudaStream_t st);
bool tc_gemm_supported(int M, int N, int K);
void set_tc_debug(int v);
void set_tc_pair(int v);    // 0: never use CTA-pair (cta_group::2) tiles   // timing experiments (see TcEpilogue::dbg)

// ---------------------------------------------------------------- batched 16-bit GEMM (mma.sync path)
// C[b] = scale * A[b] . op(B[b]) + res[b];  A (M x K) K-contiguous;  B (N x K) K-contiguous [b_kn=0] or (K x N) [b_kn=1]
struct Mma16Args {
  const void* A; long lda, sA;
  const void* B; long ldb, sB;
  void* C; long ldc, sC;
  const float* res; long ldr, sR;
  int M, N, K, batch;
  float scale;
  int b_kn;
};
template <typename T, typename OutT>
cudaError_t launch_gemm_mma16(const Mma16Args& p, cudaStream_t st);

// ---------------------------------------------------------------- skinny 16-bit GEMM (decoder-step linears)
// C = act(LN?(A) . W^T + bias) + res for M <= 512 rows: small tiles, operands resident in smem, optional LayerNorm /
// fp32->16-bit conversion of A on load, split-K over a thread-block cluster (gemm_skinny.cu).
struct SkinnyArgs {
  const void* A16;                   // (M x K) 16-bit operand, or
  const float* A32;                  // (M x K) fp32 rows, converted on load (exactly one of A16 / A32)
  long lda;
  const float *ln_g, *ln_b;          // LayerNorm over the K columns of A32 (needs K <= 1024), or null
  const void* W; long ldw;           // (N x K) 16-bit
  const float* bias;
  const float* res; long ldr;        // fp32 residual (may alias Cf)
  float* Cf; void* Cb; long ldc;     // fp32 or 16-bit output, exactly one non-null
  int M, N, K;
  int act;                           // 0 none, 1 GELU(erf), 2 ReLU
};
bool skinny_gemm_supported(const SkinnyArgs& p);
template <typename T>
cudaError_t launch_gemm_skinny(const SkinnyArgs& p, cudaStream_t st);

// ---------------------------------------------------------------- normalisation / embedding
// split-K consumer: x = res + bias + sum of `nparts` fp32 partial products (slice order), written to xout (may alias
// res); gamma != nullptr: followed by LayerNorm into y.  C % 128 == 0, C <= 1024.
template <typename OutT>
cudaError_t launch_layernorm_sum(const float* part, int nparts, long pstride, long ldp, const float* bias, const float* res, long ldr,
                                 float* xout, long ldxo, const float* gamma, const float* beta, OutT* y, long ldy, long rows, int C,
                                 cudaStream_t st);
template <typename OutT>
cudaError_t launch_layernorm(const float* x, long ldx, const float* gamma, const float* beta, OutT* y, long ldy,
                             long rows, int C, cudaStream_t st);
// PatchMerging front half: gather the 2x2 neighbourhood (order (0,0),(1,0),(0,1),(1,1)) and LayerNorm(4C).
template <typename OutT>
cudaError_t launch_merge_layernorm(const float* x, const float* gamma, const float* beta, OutT* y,
                                   int B, int H, int C, cudaStream_t st);
// PatchEmbed: conv(k = stride = P) as a K = Cin*P*P dot product + bias + LayerNorm(E).
cudaError_t launch_patch_embed(const float* img, const float* w, const float* b, const float* gamma,
                               const float* beta, float* out, int B, int Cin, int S, int P, int E,
                               cudaStream_t st);
// patch width 4: filter bank pre-packed by launch_patch_filter_pack4 into float4-over-dx rows
cudaError_t launch_patch_filter_pack4(const float* w, float* wq, int Cin, int E, cudaStream_t st);
// x16 / stats != nullptr: also the raw rows in 16 bits and their 64-bit fixed-point sum / sum of squares (producer of the
// first block's folded norm1, see TcGemmArgs::stats_out)
cudaError_t launch_patch_embed4(const float* img, const float* wq, const float* b, const float* gamma, const float* beta,
                                float* out, int B, int Cin, int S, int E, cudaStream_t st, void* x16 = nullptr, int fp16 = 0,
                                float* stats = nullptr);
// the same on the tensor cores (mma.sync TF32; 16-bit modes): E == 192, S % 64 == 0
bool patch_embed4_tc_supported(int Cin, int S, int E);
cudaError_t launch_patch_embed4_tc(const float* img, const float* wq, const float* b, const float* gamma, const float* beta,
                                   float* out, int B, int Cin, int S, int E, cudaStream_t st, void* x16, int fp16, float* stats);
template <typename T>
cudaError_t launch_cast(const float* x, T* y, long n, cudaStream_t st);

// ---------------------------------------------------------------- Swin window attention
// qkv (B*H*H, 3C) token-major in the unshifted frame -> out (B*H*H, C); cyclic shift, window
// partition/reverse, relative-position bias and the shift mask are index arithmetic.
template <typename T>
cudaError_t launch_window_attention(const T* qkv, const float* bias_table, T* out, int B, int H, int C,
                                    int heads, int shift, cudaStream_t st);
// tensor-core (mma.sync) variant for the 16-bit modes; takes the derived bias table (heads, 532) / scale built by launch_transpose_bias
template <typename T>
cudaError_t launch_window_attention_mma(const T* qkv, const float* bias_t, T* out, int B, int H, int C,
                                        int heads, int shift, cudaStream_t st);
cudaError_t launch_transpose_bias(const float* table, float* out, int heads, cudaStream_t st);
inline size_t bias_derived_floats(int heads) { return (size_t)(2 * 532 + 1024) * heads; }     // layout: window_attn_mma.cu
// tcgen05 variant (window_attn_tc.cu): scores and probabilities in TMEM, thread-per-row softmax; needs an even head count.
// Same derived bias table as the mma.sync variant.
bool window_attention_tc_supported(int B, int H, int C, int heads, int shift);
template <typename T>
cudaError_t launch_window_attention_tc(const T* qkv, const float* bias_t, T* out, int B, int H, int C, int heads, int shift,
                                       cudaStream_t st);
extern int g_attn_tc_dbg;
extern int g_attn_tc;      // 1: use the tcgen05 kernel where supported (default); 0: mma.sync kernel

// ---------------------------------------------------------------- static expansion (encoder)
// z (B, E, N) raw scores (already / sqrt(d)).  Produces forward weights (B,E,N) normalised over
// the N keys (keys >= n_valid[b] masked) and backward weights (B,N,E) normalised per group.
template <typename WT>
cudaError_t launch_static_exp_weights(const float* z, const int* n_valid, const int* group_start, int n_groups,
                                      WT* a_fw, WT* b_fw, WT* a_bw, WT* b_bw, float* gsum_scratch,
                                      int B, int E, int N, int chunk, cudaStream_t st);
// x_out = x_in + sigmoid(sel) * out_a + (1 - sigmoid(sel)) * out_b
template <typename ST>
cudaError_t launch_selector_mix(const float* x_in, long ldxi, const ST* sel, long lds, const float* out_a,
                                const float* out_b, long ldo, float* x_out, long ldxo, long rows, int d,
                                cudaStream_t st);

// Transposed-score variants for the tcgen05 path (static_exp.cu): zT (B, N, E) from the linear-layer GEMM; same outputs
// as launch_static_exp_weights; colpart_scratch holds B * (N / 16) * 2 * E floats.
bool static_exp_t_supported(const int* group_start_host, int n_groups, int E, int N);
template <typename T>
cudaError_t launch_static_exp_weights_t(const float* zT, const int* n_valid, const int* group_start, int n_groups, T* a_fw,
                                        T* b_fw, T* a_bw, T* b_bw, float* gsum_scratch, float* colpart_scratch, int B, int E,
                                        int N, cudaStream_t st);
// src [b * N + n][ld], columns c0 .. c0 + C (16-bit)  ->  dst [b][C][N]
template <typename T>
cudaError_t launch_transpose_ab(const T* src, long ld, int c0, T* dst, int B, int C, int N, cudaStream_t st);
cudaError_t launch_transpose_f32(const float* in, float* out, int R, int C, cudaStream_t st);   // out[c][r] = in[r][c]
// selector mix with out_a / out_b given transposed: [b][d][N]
template <typename ST>
cudaError_t launch_selector_mix_t(const float* x_in, long ldxi, const ST* sel, long lds, const float* out_a_t, const float* out_b_t,
                                  float* x_out, long ldxo, int B, int N, int d, cudaStream_t st);

// ---------------------------------------------------------------- decoder step
struct DecState {
  // per decoder layer l, position p, row r:  cache[((l*P + p)*R + r)*cw ...] = [cond | key | A | B | sel]
  float* cache; int cw;        // cw = 5*d
  float* fw;                   // forward weights of the row-block of position p: ((l*P+p)*R + r)*2*n_exp*P
  float* qk;                   // q_e . K_p   : ((l*P+p)*R + r)*n_exp
  const int* anc;              // (R, P): slot (row index) holding position i of row r's history
  int P;                       // max positions
  int R;
};
cudaError_t launch_embed(const int64_t* tokens64, const int* tokens32, long tok_stride, int p, const float* emb,
                         const float* pos, float* x, long ldx, int R, int d, cudaStream_t st);
// embedding + LayerNorm (first decoder layer's norm_1) in one pass; d == 512
template <typename T>
cudaError_t launch_embed_ln(const int64_t* tokens64, const int* tokens32, long tok_stride, int p, const float* emb,
                            const float* pos, float* x, long ldx, const float* gamma, const float* beta, T* xn, long ldn, int R,
                            int d, cudaStream_t st);
// ln_out != nullptr (d == 512): also writes LayerNorm(x_out) in the operand type (the layer's norm_2)
template <typename T>
cudaError_t launch_dyn_exp_step(const DecState& s, int layer, int p, const float* qexp, const float* bexp,
                                int n_exp, const int* row_len, const float* x_in, long ldxi, float* x_out,
                                long ldxo, int d, const float* ln_g, const float* ln_b, T* ln_out, long ldn, cudaStream_t st);
// cross attention of one query position per row against per-image K/V (shared by the beams)
template <typename KvT, typename OutT>
cudaError_t launch_cross_attn_step(const float* q, long ldq, const KvT* kv, long ldkv, int k_off, int v_off,
                                   OutT* out, long ldo, int R, int rows_per_image, int n_keys, int heads, int dk,
                                   const int* n_valid, const int* row_len, int p, cudaStream_t st);
// beam-search state (beam.cu); declared here because the persistent whole-search kernel takes it too
struct BeamBufs {
  int* tokens[2];     // (B, beam, L) ping-pong
  float* lps[2];      // (B, beam, L)
  int* len[2];        // (B, beam)
  int* anc[2];        // (B*beam, L)   slot is local (0..beam-1) + b*beam
  float* cum[2];      // (B, beam)     running sum of the history log-probs
  int* eos[2];        // (B, beam)     history contains EOS
  int* all_done;      // [1]
  int* grew;          // [L]  grew[t] != 0: some beam was extended at time step t (else every beam had ended: the search is over)
  int* final_src;     // [1]  ping-pong index holding the state after the last EXECUTED step (steps may be skipped, see engine.cu)
};

// ---------------------------------------------------------------- persistent whole-position decoder kernel (decode_mega.cu)
// One launch = one decoder position for all rows (16-bit modes, d_model 512, head width 64): embedding, every decoder
// layer, reduce group, final norm + vocabulary projection; with topk > 0 also log-softmax + top-k (the logits are then
// never stored).  Phases are separated by a grid barrier; the grid is launched cooperatively.
constexpr int kMegaMaxLayers = 4;
struct MegaLayer {
  const float *n1g, *n1b, *n2g, *n2b, *n3g, *n3b, *qexp, *bexp;
  const void *w_dyn5, *w_wq, *w_wo, *w_ff1, *w_ff2;          // 16-bit, slab-packed (launch_mega_pack_weight)
  const float *b_dyn5, *b_wq, *b_wo, *b_ff1, *b_ff2;
};
struct MegaArgs {
  DecState s;
  int p, n_layers, d, ff, n_exp, heads, n_keys, vocab, R, rows_per_image;
  MegaLayer L[kMegaMaxLayers];
  const int64_t* tok64; const int* tok32; long tok_stride;
  const int* n_valid; const int* row_len;
  const float *emb, *pos;
  float *x0, *ycat, *q, *pre;                 // fp32: embedding, layer outputs side by side (ld d*n_layers), queries, reduce output
  void *xn, *att, *hid, *ycat16;              // 16-bit operands, slab-packed (mega_act_bytes): rows of 520, K slabs of 512
  const void* kv; long ldkv;                  // cross K/V of all layers per encoder token: [layer][K | V]
  const void* w_reduce; const float* b_reduce;
  const float *ng, *nb;
  const void* w_vocab; const float* b_vocab;
  float* logits; long ldl;                    // topk == 0: logits out
  int topk; void* parts; float* top_val; int* top_idx;      // topk > 0: log-prob / index of the k best words per row
  unsigned* bar;                              // kMegaBarBytes: grid-barrier words + split-K tile counters (zeroed once; self-restoring)
  int max_ctas_per_sm;                        // 0 / 2: two CTAs per SM; 1: one (two decode groups run two of these kernels side by side)
  int ksplit_ff2, ksplit_red; float* scratch; // split-K over CTAs of the two long-K projections (partial tiles in `scratch`)
  int dbg_mode;                               // timing experiments: bit 0 skip the MMAs, bit 1 skip the operand loads, bit 2 skip the LayerNorm fill
  unsigned long long* dbg;                    // optional: CTA 0 stores %globaltimer before / after every barrier (2 per phase)
};
bool mega_supported(const MegaArgs& a);
size_t mega_parts_bytes(int R, int vocab);
size_t mega_scratch_bytes(int R);
size_t mega_packed_bytes(int N, int K);          // slab-packed 16-bit copy of an (N x K) weight, K % 512 == 0
cudaError_t launch_mega_pack_weight(const float* w, void* out, int N, int K, int fp16, cudaStream_t st);
size_t mega_act_bytes(int R, int ff, int n_layers);
constexpr size_t kMegaBarBytes = 20 * 1024;
template <typename T>
cudaError_t launch_dec_step_mega(const MegaArgs& a, cudaStream_t st);
// the whole 'max' beam search over an encoder output in one launch (time steps loop inside the kernel); `a` as for one
// position with topk = beam, tokens / ancestry taken from bb (initialised by launch_beam_init)
struct MegaSearch {
  BeamBufs bb; int beam, L, eos, how_many, early_exit;
  int* r_tok; int* r_len; float* r_lp;         // results as launch_beam_finalize writes them
};
template <typename T>
cudaError_t launch_dec_search_mega(const MegaArgs& a, const MegaSearch& q, cudaStream_t st);
extern int g_mega_coop;                       // 1: cooperative launch (default)

// ensemble of up to kMaxEnsemble models: out = log(mean_m softmax(logits_m)) per row
constexpr int kMaxEnsemble = 8;
struct EnsembleLogits { const float* p[kMaxEnsemble]; int n; };
cudaError_t launch_ensemble_logprob(const EnsembleLogits& in, long ld, int rows, int V, float* out, long ldo, cudaStream_t st);
// write_mode bit 0: also store the log-probabilities; bit 1: the input already holds log-probabilities (top-k only)
cudaError_t launch_logsoftmax_topk(const float* logits, long ld, int rows, int V, int k, float* top_val,
                                   int* top_idx, float* logprob, long ldlp, int write_mode, cudaStream_t st);

// Sampling variant (reference captioning_model.py sample_or_max='sample' / mode='sampling'): k draws WITHOUT replacement
// from softmax(logits) per row, in draw order, by the Gumbel-top-k construction -- rank log p_i + G_i with i.i.d. standard
// Gumbel noise G_i from a counter-based generator keyed on (seed, row, step, i) -- which has exactly the distribution of
// torch.multinomial(p, k, replacement=False) (k = 1: a Categorical draw).  top_val receives the UNPERTURBED log p.
cudaError_t launch_gumbel_topk(const float* logits, long ld, int rows, int V, int k, uint64_t seed, int step, float* top_val,
                               int* top_idx, cudaStream_t st);

// ---------------------------------------------------------------- image preprocessing (Pillow-exact resize + normalise)
// Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter: bounds (out_size x {first, count}) and
// 22-bit fixed-point weights (out_size x ksize)
void resample_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk, int* ksize_out);
cudaError_t launch_preprocess_rgb8(const uint8_t* rgb_dev, int H, int W, int S, const int* bounds_x, const int* kk_x, int ksize_x,
                                   const int* bounds_y, const int* kk_y, int ksize_y, uint8_t* tmp_dev, float* out_dev,
                                   cudaStream_t st);

// batched form: one item per image, all pointers device memory
struct PreItem {
  const uint8_t* src; uint8_t* tmp; float* out;
  const int *bx, *kx, *by, *ky;
  int H, W, ksx, ksy;
};
cudaError_t launch_preprocess_rgb8_batch(const PreItem* items_dev, int n, int max_h, int S, cudaStream_t st);

// ---------------------------------------------------------------- beam search bookkeeping
// device-side early exit (CUDA-graph conditional nodes): sets the condition of the next step's IF node to grew[t] != 0
cudaError_t launch_beam_set_condition(unsigned long long cond_handle, const int* grew_t, cudaStream_t st);
cudaError_t launch_beam_init(const BeamBufs& bb, int B, int beam, int L, int sos, cudaStream_t st);
cudaError_t launch_beam_first(const BeamBufs& bb, const float* top_val, const int* top_idx, int B, int beam,
                              int L, int eos, cudaStream_t st);
cudaError_t launch_beam_step(const BeamBufs& bb, int src, const float* top_val, const int* top_idx, int B,
                             int beam, int L, int t, int eos, cudaStream_t st);
cudaError_t launch_beam_finalize(const BeamBufs& bb, int src, int B, int beam, int L, int t_final, int how_many,
                                 int* out_tokens, int* out_len, float* out_lp, cudaStream_t st);

// mode='sampling' bookkeeping (legacy_models/captioning_model.py:60-109): independent rows, one sampled word per step
cudaError_t launch_sample_append(const BeamBufs& bb, const float* top_val, const int* top_idx, int R, int L, int t, int eos,
                                 cudaStream_t st);
cudaError_t launch_sample_finalize(const BeamBufs& bb, int R, int L, int t_final, int* out_tokens, int* out_len, float* out_lp,
                                   cudaStream_t st);

}  // namespace xn
