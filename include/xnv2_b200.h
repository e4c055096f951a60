/*
 * xnv2_b200 -- C ABI of the B200-native ExpansionNet v2 captioning path.
 *
 * The reference (nighting0le01/On_Device_Image_Captioning) is pure Python/PyTorch and has
 * no FFI of its own; its "plugin boundary" for this path is the Python class surface
 *   End_ExpansionNet_v2.forward_enc / forward_dec      models/End_ExpansionNet_v2.py:121-209
 *   ExpansionNet_v2.forward_enc / forward_dec          models/ExpansionNet_v2.py:76-156
 *   CaptioningModel.forward(mode=...) / beam_search    legacy_models/captioning_model.py:24-57,111-241
 *   Captioner.__call__ / beam_search                   models/captioning_model.py:67-110,220-427
 * Every entry point below names the reference interface it replaces.  The Python shim in
 * on_device_image_captioning_b200/ binds these symbols with ctypes and re-exposes the
 * reference's class/method signatures.
 *
 * Beyond the hot path: xn_preprocess_rgb8 (utils/image_utils.py), xn_ensemble_beam_search
 * (models/ensemble_captioning_model.py).
 *
 * Conventions: every call returns 0 on success and a negative code on failure (never
 * throws); xn_last_error() gives the message.  All tensor arguments are caller-owned
 * DEVICE pointers unless the name says host; work is enqueued on the given CUDA stream
 * (passed as void* == cudaStream_t) and is stream-ordered: the caller synchronises the
 * stream before reading results.  A handle is bound to one device and is not
 * thread-safe.  Handles are independent of each other: calls on DIFFERENT handles may be
 * in flight at the same time on different streams (two handles with the same weights
 * taking alternate xn_caption_host_begin calls is the throughput serving pattern, see
 * INTEGRATION.md); xn_set_option names marked process-wide in the list below affect
 * every handle.  There is no CPU fallback: without a CUDA device xn_create fails.
 */
#ifndef XNV2_B200_H
#define XNV2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xn_handle xn_handle;

enum { XN_OK = 0, XN_ERR_ARG = -1, XN_ERR_CUDA = -2, XN_ERR_STATE = -3, XN_ERR_UNSUPPORTED = -4 };
enum { XN_PREC_FP32 = 0, XN_PREC_BF16 = 1, XN_PREC_FP16 = 2 };   /* 16-bit modes: tcgen05 operands, fp32 accumulate */
enum { XN_DTYPE_F32 = 0, XN_DTYPE_I64 = 1 };

/* Model geometry: the constructor arguments of End_ExpansionNet_v2 / ExpansionNet_v2
 * (models/End_ExpansionNet_v2.py:11-47, models/ExpansionNet_v2.py:10-25). */
typedef struct xn_config {
  int32_t has_swin;          /* 1: End_ExpansionNet_v2 (images in), 0: ExpansionNet_v2 (features in) */
  int32_t img_size, patch_size, in_chans, embed_dim;
  int32_t n_stages;
  int32_t depths[4];
  int32_t swin_heads[4];
  int32_t window_size;       /* 12 */
  float   mlp_ratio;         /* 4.0 */
  int32_t feat_dim;          /* final_swin_dim / img_feature_dim */
  int32_t d_model, n_enc, n_dec, ff, num_heads;
  int32_t n_exp_groups;
  int32_t exp_groups[8];     /* num_exp_enc_list */
  int32_t num_exp_dec;
  int32_t vocab;
  int32_t max_seq_len;       /* rows of pos_encoder */
  int32_t enc_len;           /* visual tokens entering the expansion encoder (144) */
} xn_config;

/* Replaces: the model constructor.  `device` is the CUDA ordinal ("rank"). */
int xn_create(const xn_config* cfg, int device, xn_handle** out);
int xn_destroy(xn_handle* h);
const char* xn_last_error(const xn_handle* h);   /* h may be NULL: last error of xn_create */

/* Replaces: model.load_state_dict(checkpoint["model_state_dict"]) (demo.py:100-104,
 * test.py:453-459).  One call per state_dict entry, key and shape exactly as in the
 * reference checkpoint.  `data` may be a host or a device pointer
 * (unified addressing).  Buffers (relative_position_index, attn_mask) are geometry-only:
 * they are accepted and ignored. */
int xn_load_tensor(xn_handle* h, const char* key, const void* data, int dtype,
                   const int64_t* shape, int ndim);
/* Packs the loaded tensors for the chosen precision; fails if a key is missing. */
int xn_finalize_weights(xn_handle* h, int precision);

/* Replaces: SwinTransformer.forward_features (models/swin_transformer_mod.py:801-813).
 * images (B,in_chans,S,S) f32 NCHW -> out (B, L_last, C_last) f32. */
int xn_forward_swin(xn_handle* h, const float* images, int B, float* out, void* stream);

/* Replaces: forward_enc (models/End_ExpansionNet_v2.py:121-153, models/ExpansionNet_v2.py:76-100).
 * input: images (has_swin) or features (B,enc_len,feat_dim).  enc_pads_host: B ints (host)
 * or NULL (== all 0; must be all 0 when has_swin).  out (B,enc_len,d_model) f32. */
int xn_forward_enc(xn_handle* h, const float* input, int B, const int32_t* enc_pads_host,
                   float* out, void* stream);

/* Replaces: forward_dec (models/End_ExpansionNet_v2.py:155-209, models/ExpansionNet_v2.py:102-156).
 * cross (R,enc_len,d_model) f32, tokens (R,t) i64, pads: R host ints or NULL.
 * out (R,t,vocab) f32: logits, or log-probabilities if apply_log_softmax != 0. */
int xn_forward_dec(xn_handle* h, const float* cross, int R, const int32_t* enc_pads_host,
                   const int64_t* tokens, int t, const int32_t* dec_pads_host,
                   int apply_log_softmax, float* out, void* stream);

/* Replaces: beam_search, 'max' branch (legacy_models/captioning_model.py:111-241 ==
 * models/captioning_model.py:220-427), including forward_enc.  Limits (XN_ERR_ARG beyond them): 1 <= how_many <= beam <= 8,
 * 2 <= max_len <= min(max_seq_len, 128).  `input` may be any device tensor: it is copied into a handle-owned staging
 * buffer first, so the call's cached CUDA graph does not depend on the caller's address.
 * out_tokens  (B,how_many,max_len) i32, -1 padded, SOS..EOS inclusive
 * out_len     (B,how_many) i32
 * out_logprob (B,how_many,max_len) f32, 0 padded                                        */
int xn_beam_search(xn_handle* h, const float* input, int B, const int32_t* enc_pads_host,
                   int beam, int max_len, int how_many, int sos_idx, int eos_idx,
                   int32_t* out_tokens, int32_t* out_len, float* out_logprob, void* stream);
/* Same, starting from an encoder output already on the device (B,enc_len,d_model). */
int xn_beam_search_from_enc(xn_handle* h, const float* enc_out, int B, const int32_t* enc_pads_host,
                            int beam, int max_len, int how_many, int sos_idx, int eos_idx,
                            int32_t* out_tokens, int32_t* out_len, float* out_logprob, void* stream);

/* End-to-end convenience for callers holding HOST buffers (the call bench.py's e2e leg
 * times): pinned/pageable host images -> host tokens; copies are issued on `stream` and the
 * call returns after the results have landed. */
int xn_caption_host(xn_handle* h, const float* input_host, int B, int beam, int max_len, int how_many,
                    int sos_idx, int eos_idx, int32_t* out_tokens_host, int32_t* out_len_host,
                    float* out_logprob_host, void* stream);

/* Pipelined form of xn_caption_host for a caller that streams batches (bench.py's e2e leg at N GPUs): _begin enqueues
 * the host-to-device copy of this batch on the handle's copy stream, the search on `stream` and the device-to-host copies
 * of the results, and returns a ticket (0 or 1) without waiting; _end(ticket) blocks until that call's results are in the
 * host buffers.  Two calls may be in flight: the copy of batch i+1 overlaps the compute of batch i (two staging slots).
 * The host buffers of a call must stay valid and untouched until its _end.  Pinned host memory is needed for overlap. */
int xn_caption_host_begin(xn_handle* h, const float* input_host, int B, int beam, int max_len, int how_many,
                          int sos_idx, int eos_idx, int32_t* out_tokens_host, int32_t* out_len_host,
                          float* out_logprob_host, void* stream);      /* >= 0: ticket; < 0: error */
int xn_caption_host_end(xn_handle* h, int ticket);

/* Replaces: beam_search with sample_or_max='sample' (legacy_models/captioning_model.py:131-133,168-170 ==
 * models/captioning_model.py:256-260,305-308): the per-step candidates of every beam are `beam` draws WITHOUT replacement
 * from its word distribution (torch.multinomial(exp(log_probs), beam, replacement=False)) instead of the top-k; the rest
 * of the search is xn_beam_search's.  Draws come from a counter-based generator keyed on (seed, row, step, word): the
 * same seed reproduces the same captions; the distribution, not torch's random stream, is what is matched. */
int xn_beam_search_sample(xn_handle* h, const float* input, int B, const int32_t* enc_pads_host, int beam, int max_len,
                          int how_many, int sos_idx, int eos_idx, uint64_t seed, int32_t* out_tokens, int32_t* out_len,
                          float* out_logprob, void* stream);
/* Replaces: mode='sampling' -> get_batch_multiple_sampled_prediction (legacy_models/captioning_model.py:60-109 ==
 * models/captioning_model.py:120-218; the SCST sampler, train.py:146-151): num_outputs (<= 8) independent ancestral
 * samples per image, one Categorical draw per step, up to max_len sampled words after SOS.
 * out_tokens  (B,num_outputs,max_len+1) i32, -1 padded: SOS .. first sampled EOS inclusive (all max_len+1 if none)
 * out_len     (B,num_outputs) i32
 * out_logprob (B,num_outputs,max_len+1) f32: log-probability of every sampled word, 0 at SOS and after the first EOS */
int xn_sample(xn_handle* h, const float* input, int B, const int32_t* enc_pads_host, int num_outputs, int max_len,
              int sos_idx, int eos_idx, uint64_t seed, int32_t* out_tokens, int32_t* out_len, float* out_logprob,
              void* stream);

/* JPEG files straight to the normalised tensor (SURVEY.md 8f N1 incl. the decode): nvJPEG decodes every stream on the GPU
 * into RGB8, then the batched resize / ToTensor / Normalize of xn_preprocess_rgb8_batch runs on it -- the pixels never
 * visit the host.  jpeg_ptrs_host: n host pointers to complete JPEG streams, jpeg_sizes: their byte counts.  heights_out /
 * widths_out (n ints, may be NULL) receive the decoded sizes.  A stream that is not 3-component (grayscale, CMYK) gives the
 * blank canvas the reference substitutes for non-RGB files (utils/image_utils.py:18-19).  nvJPEG's IDCT / chroma
 * upsampling is not bit-identical to libjpeg's (what PIL runs in the reference): pixels differ by a few grey levels, so
 * the bit-exact path remains "decode with PIL, xn_preprocess_rgb8_batch".  libnvjpeg is loaded with dlopen at first use;
 * without it this call returns XN_ERR_UNSUPPORTED and xn_jpeg_available() returns 0. */
int xn_jpeg_available(void);
int xn_preprocess_jpeg_batch(xn_handle* h, const uint8_t* const* jpeg_ptrs_host, const int64_t* jpeg_sizes, int n,
                             float* out, int out_size, int32_t* heights_out, int32_t* widths_out, void* stream);

/* 16-bit modes store QKV, attention outputs and MLP hidden activations as fp16 / bf16.  fp16 tops out at 65504: if a
 * checkpoint's activations exceed that, infinities reach the encoder output as NaN.  Every xn_forward_enc / xn_beam_search
 * / xn_caption_host call in a 16-bit mode checks its encoder output on the device and raises this flag when it holds a
 * non-finite value.  Reads (and optionally clears) the flag; synchronises the device.  Remedy: XN_PREC_BF16 or FP32. */
int xn_overflow_flag(xn_handle* h, int* flag_out, int clear);

/* Counters / introspection. */
/* Ensemble beam search (replaces EsembleCaptioningModel.forward(mode="beam_search"),
 * legacy_models/ensemble_captioning_model.py:19-241, built at test.py:334): up to 8 handles on one device with the same
 * vocabulary and input geometry decode in lock step; the step distribution is log(mean_m softmax(logits_m)), the
 * search itself (EOS rules, beam^2 merge, length bookkeeping) is xn_beam_search's.  'max' branch only. */
int xn_ensemble_beam_search(xn_handle* const* handles, int n_models, const float* input, int B, const int32_t* enc_pads_host,
                            int beam, int max_len, int how_many, int sos_idx, int eos_idx, int32_t* out_tokens,
                            int32_t* out_len, float* out_logprob, void* stream);

/* Image preprocessing (replaces utils/image_utils.py:5-23 preprocess_image after the decode):
 * RGB8 (H x W x 3, interleaved) -> float32 (3 x S x S) = Normalize(ToTensor(Resize((S,S))(image))), bit-identical to
 * Pillow's antialiased bilinear resize + torchvision's float32 tail.  `rgb` is a host pointer (rgb_on_device = 0: copied
 * inside the call, stream-ordered) or a device pointer; `out` is device memory.  Any H, W >= 1; S = out_size. */
int xn_preprocess_rgb8(xn_handle* h, const uint8_t* rgb, int rgb_on_device, int H, int W, float* out, int out_size, void* stream);
/* Batched form: n images of arbitrary sizes in ONE launch pair, no per-image synchronisation.  rgb_ptrs: n pointers (all
 * host or all device, per rgb_on_device), heights / widths: n host ints, out: (n,3,S,S) f32 device.  Host images are
 * copied inside the call, stream-ordered (pageable memory is staged by the driver; pinned memory must stay valid until
 * the stream has passed the call). */
int xn_preprocess_rgb8_batch(xn_handle* h, const uint8_t* const* rgb_ptrs, int rgb_on_device, const int32_t* heights,
                             const int32_t* widths, int n, float* out, int out_size, void* stream);

int64_t xn_kernel_launches(const xn_handle* h);      /* kernels of this library launched so far */
int64_t xn_workspace_bytes(const xn_handle* h);
/* Tuning / measurement switches (defaults are the measured best; see DESIGN.md and profiles/README.md):
 *   "swin_chunk", "enc_chunk"  images per Swin / encoder chunk (64)
 *   "use_graph"                1: calls are captured into a CUDA graph on their second occurrence and replayed
 *   "decode_groups"            image groups decoded on concurrent graph branches (0 = automatic: 2 from 32 images)
 *   "early_exit"               device-side early termination of the beam search (reference captioning_model.py:397): inside a
 *                              captured call the decode steps are grouped into CUDA-graph IF nodes of this many steps that are
 *                              skipped once every beam has ended (default 4; 0 = off: all steps always run, same results)
 *   "attn_tc"                  1: window attention on the tcgen05 kernel (default), 0: the mma.sync kernel (process-wide)
 *   "pdl"                      programmatic dependent launch (process-wide)
 *   "tc_pair"                  CTA-pair (cta_group::2) GEMM tiles for long-K shapes (process-wide)
 *   "use_skinny"               skinny mma.sync GEMM for decoder-step linears with <= 64 rows
 *   "ln_fuse"                  Swin norm1 / norm2 folded algebraically into the neighbouring tcgen05 GEMMs (16-bit modes): 1 = all but
 *                              each stage's first norm1, 2 (default) = those too (produced by the patch embedding / merge GEMM)
 *   "se_tc"                    static-expansion block of the encoder in the 16-bit modes (reference models/layers.py:20-102):
 *                              0 = per-image mma.sync contractions + (B,E,N)-layout normaliser kernels (round 1);
 *                              1 = scores as ONE transposed linear-layer launch on tcgen05 + 16-token slab kernels;
 *                              2 (default) = also class^T / out^T as batched tcgen05 GEMMs (3-D tensor maps, every operand K-major)
 *   "pe_tc"                    1 (default): patch embedding of the 16-bit modes on the tensor cores (TF32 mma.sync, patch width 4,
 *                              embed_dim 192, image side % 64 == 0); 0: the fp32 CUDA-core kernel (always used by the fp32 mode)
 *   "dec_splitk"               1: the long-K decoder-step linears (ff2, reduce group) run as K / 512 slices on the batched tcgen05
 *                              GEMM and are summed, in slice order, by the LayerNorm launch that follows; default 0 (measured: a tie)
 *   "ln_on_load"               decoder-step LayerNorm computed inside the consuming tcgen05 GEMM (off: measured slower)
 *   "use_mega"                 1: every decoder position of the 16-bit modes (d_model 512, head width 64, <= 20 positions,
 *                              16 expansion vectors) runs as ONE persistent cooperative kernel with grid barriers between its
 *                              phases (csrc/decode_mega.cu) instead of ~33 launches; default 0 (measured: a tie at <= 192 rows,
 *                              slower beyond).  "fuse_topk" (1): that kernel also does log-softmax + top-k of the 'max' search,
 *                              the R x V logits are never stored.  "mega_search" (0): all time steps of the search in one launch
 *                              (measured slower).  "mega_coop" (1): cooperative launch (process-wide).  "mega_dbg": phase timestamps for
 *                              xn_mega_timeline; "mega_dbg_mode": timing experiments (skip MMAs / loads / fills)
 *   "profile"                  1: event-time every tcgen05 GEMM; 2: event-time every kernel launch (both disable graphs)
 *   "tc_debug", "op_out16"     kernel timing experiments / test hooks */
int xn_set_option(xn_handle* h, const char* name, int64_t value);
/* With option "profile"=1 every tcgen05 GEMM launch is bracketed by CUDA events on its stream;
 * this returns the summed device time, the summed algorithmic FLOPs and the launch count since
 * the option was set (synchronises the device). */
int xn_profile_read(xn_handle* h, double* ms_total, double* flops_total, int64_t* count);
/* Same, restricted to launches of at least min_flops floating-point operations (separates the Swin / encoder GEMMs from
 * the latency-bound decoder-step GEMMs in the roofline report). */
int xn_profile_read_min(xn_handle* h, double min_flops, double* ms_total, double* flops_total, int64_t* count);
/* With option "profile"=2 EVERY kernel launch of the library is bracketed by CUDA events (graphs off) and attributed to
 * its launcher; this writes "launcher<TAB>launches<TAB>total_ms" lines into buf (bench.py's bandwidth-roofline leg). */
int xn_profile_kernels(xn_handle* h, char* buf, int cap);
/* With option "mega_dbg"=1 CTA 0 of the persistent decoder-position kernel (decode_mega.cu) stores the device's
 * %globaltimer (ns) at its start, before and after every grid barrier and at its end; this returns the number of
 * timestamps of the LAST launch (>= 0) and copies up to `cap` of them (synchronises the device). */
int xn_mega_timeline(xn_handle* h, uint64_t* out, int cap);

/* Single-operator entry points (kernel-level parity tests call these through the ABI).
 * All pointers device, row-major, f32 unless noted. */
int xn_op_layernorm(xn_handle* h, const float* x, const float* gamma, const float* beta, float* y,
                    int rows, int C, void* stream);
int xn_op_linear(xn_handle* h, const float* x, const float* w, const float* bias, const float* residual,
                 float* y, int M, int N, int K, int act /*0 none,1 gelu,2 relu*/, int precision, void* stream);
/* Decoder-step linear on the latency-oriented kernel (gemm_skinny.cu; 16-bit precisions, M <= 512):
 * y = act(LN(x; gamma, beta) . w^T + bias) + residual, LayerNorm optional (gamma == NULL: plain conversion of the fp32
 * rows; the fp32 path takes K slices <= 512).  x_is_16bit != 0 rounds x to the operand type first (the cp.async A path); 0 feeds the fp32 rows to the kernel.
 * Replaces nn.LayerNorm + nn.Linear pairs of reference models/layers.py:222-248 at one new position per row. */
int xn_op_linear_skinny(xn_handle* h, const float* x, const float* gamma, const float* beta, const float* w,
                        const float* bias, const float* residual, float* y, int M, int N, int K, int act,
                        int x_is_16bit, int precision, void* stream);
/* Measurement hook: one GEMM launch on operands that are already 16-bit on the device (no conversions), y fp32.
 * which = 0: tcgen05 kernel, 1: skinny kernel with a 16-bit A, 2: skinny kernel with fp32 A (+ LayerNorm if gamma),
 * 3: tcgen05 kernel with LayerNorm-on-load (a = fp32 rows, K = 512; gamma may be NULL for a plain conversion). */
int xn_op_gemm_raw(xn_handle* h, int which, const void* a, const float* gamma, const float* beta, const void* w16,
                   const float* bias, const float* residual, float* y, int M, int N, int K, int act, int precision, void* stream);
int xn_op_window_attention(xn_handle* h, const float* qkv, const float* bias_table, float* out,
                           int B, int H, int C, int heads, int shift, int precision, void* stream);
int xn_op_logsoftmax_topk(xn_handle* h, const float* logits, int rows, int V, int k,
                          float* top_val, int32_t* top_idx, float* logprob_or_null, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XNV2_B200_H */
