// xnv2_b200 engine: weight store, workspace, forward orchestration and the C ABI
// (include/xnv2_b200.h).  Host code only launches kernels of this library on the caller's
// stream; there is no CPU compute path.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <unordered_map>
#include <algorithm>
#include <type_traits>
#include <dlfcn.h>
#include <nvjpeg.h>

#include "../../include/xnv2_b200.h"
#include "kernels.h"

using namespace xn;

cudaError_t launch_widen_16(const void* x, float* y, long n, int fp16, cudaStream_t st);
cudaError_t launch_nonfinite_flag(const float* x, long n, int* flag, cudaStream_t st);

namespace {

std::string g_create_error;

struct DevTensor {
  float* p = nullptr;
  std::vector<int64_t> shape;
  size_t n = 0;
};

struct LinW {
  const float* w = nullptr;   // (N, K) fp32
  const void* wb = nullptr;   // (N, K) 16-bit copy (bf16 / fp16 modes)
  const void* wp = nullptr;   // slab-packed 16-bit copy for the persistent decoder kernel (decode_mega.cu)
  const float* b = nullptr;   // (N) or null
  int N = 0, K = 0;
};

struct SwinBlockW {
  const float *n1g, *n1b, *n2g, *n2b, *rpb, *rpb_t;
  LinW qkv, proj, fc1, fc2;
  LinW qkv_ln, fc1_ln;      // norm1 / norm2 folded into the weights (16-bit modes; kernels.h: TcGemmArgs::ln_stats)
};
// folded-LayerNorm hooks of one tcgen05 GEMM launch (producer and / or consumer side)
struct LnFuse { float* stats_out = nullptr; void* x16_out = nullptr; long ldx16 = 0; const float* ln_stats = nullptr; int ln_k = 0; float* stats_zero = nullptr; };
struct SwinStageW {
  int C, H, heads;
  std::vector<SwinBlockW> blocks;
  bool has_merge = false;
  const float *mg = nullptr, *mb = nullptr;
  LinW red;
};
struct EncLayerW { const float *n1g, *n1b, *n2g, *n2b, *qexp, *bexp; const void* qexp16; const float* bexpT; LinW kabs, ff1, ff2; };
struct DecLayerW { const float *n1g, *n1b, *n2g, *n2b, *n3g, *n3b, *qexp, *bexp; LinW dyn5, wq, wo, ff1, ff2; };

struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  bool overflow = false;
  void reset() { off = 0; overflow = false; }
  template <typename T> T* get(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    if (off > cap) overflow = true;
    return p;
  }
};

}  // namespace

// nvJPEG is loaded at first use with dlopen, so the library itself has no hard dependency on it: a machine without
// libnvjpeg keeps everything but xn_preprocess_jpeg_batch (which then reports XN_ERR_UNSUPPORTED).
struct NvJpegApi {
  void* lib = nullptr;
  bool tried = false;
  decltype(&nvjpegCreateSimple) create = nullptr;
  decltype(&nvjpegDestroy) destroy = nullptr;
  decltype(&nvjpegJpegStateCreate) state_create = nullptr;
  decltype(&nvjpegJpegStateDestroy) state_destroy = nullptr;
  decltype(&nvjpegGetImageInfo) image_info = nullptr;
  decltype(&nvjpegDecode) decode = nullptr;
  bool load() {
    if (tried) return lib != nullptr;
    tried = true;
    for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (lib) break;
    }
    if (!lib) return false;
    create = reinterpret_cast<decltype(create)>(dlsym(lib, "nvjpegCreateSimple"));
    destroy = reinterpret_cast<decltype(destroy)>(dlsym(lib, "nvjpegDestroy"));
    state_create = reinterpret_cast<decltype(state_create)>(dlsym(lib, "nvjpegJpegStateCreate"));
    state_destroy = reinterpret_cast<decltype(state_destroy)>(dlsym(lib, "nvjpegJpegStateDestroy"));
    image_info = reinterpret_cast<decltype(image_info)>(dlsym(lib, "nvjpegGetImageInfo"));
    decode = reinterpret_cast<decltype(decode)>(dlsym(lib, "nvjpegDecode"));
    if (!create || !destroy || !state_create || !state_destroy || !image_info || !decode) { dlclose(lib); lib = nullptr; }
    return lib != nullptr;
  }
};
static NvJpegApi g_nvjpeg;

struct xn_handle {
  xn_config cfg;
  int device = 0;
  int precision = -1;                 // -1: weights not finalised
  std::string err;
  std::unordered_map<std::string, DevTensor> raw;
  std::vector<void*> owned;           // packed buffers
  // packed views
  const float *pe_w, *pe_b, *pe_g, *pe_beta, *swin_ng, *swin_nb;
  const float* pe_wq = nullptr;       // patch-embed filter packed for the patch-width-4 kernel
  std::vector<SwinStageW> stages;
  std::vector<EncLayerW> enc;
  std::vector<DecLayerW> dec;
  LinW input_linear, vocab, enc_reduce, dec_reduce, kv_all;
  const float *enc_ng, *enc_nb, *dec_ng, *dec_nb, *emb, *pos;
  int* group_start_dev = nullptr;
  int n_exp_total = 0, exp_chunk = 8;
  bool se_t_ok = false;               // group layout admits the transposed-score static-expansion kernels (static_exp.cu)
  int64_t dec_splitk = 0;             // 16-bit modes: long-K decoder-step linears as K slices summed by the following LayerNorm (measured: a tie, off)
  int64_t pe_tc = 1;                  // 16-bit modes: patch embedding on the tensor cores (TF32 mma.sync)
  int64_t se_tc = 2;                  // 16-bit modes: 0 = (B,E,N) kernels + mma.sync contractions, 1 = scores on tcgen05 + slab kernels, 2 = + tcgen05 class / out contractions
  Arena ws;
  int64_t launches = 0;
  int64_t swin_chunk = 64, enc_chunk = 64;
  // CUDA-graph cache: the decode loop (kind 0, keyed by the encoder-output pointer) or the whole
  // images -> token ids forward (kind 1, keyed by the input pointer); one entry per distinct call shape / buffer set
  struct GraphKey {
    int kind; const void* in; int B, beam, L, how_many, sos, eos; const char* ws_base; size_t ws_cap;
    bool operator==(const GraphKey& o) const {
      return kind == o.kind && in == o.in && B == o.B && beam == o.beam && L == o.L && how_many == o.how_many && sos == o.sos &&
             eos == o.eos && ws_base == o.ws_base && ws_cap == o.ws_cap;
    }
  };
  // `h2d`: the host-to-device copy nodes of a captured xn_caption_host call (offset of their source inside the caller's
  // host buffer); on replay with another host buffer their source pointers are patched in place, so the graph does not
  // depend on the caller's buffer address.  The cudaGraph_t is kept alive because the node handles belong to it.
  struct H2DNode { cudaGraphNode_t node; void* dst; size_t src_off, bytes; };
  struct CachedGraph {
    GraphKey key; cudaGraphExec_t exec; int64_t launches;
    cudaGraph_t graph = nullptr; const char* host_base = nullptr; std::vector<H2DNode> h2d;
  };
  static constexpr size_t kMaxGraphs = 16;
  std::vector<CachedGraph> graphs;   // least recently used first
  int64_t use_graph = 1;
  int64_t op_out16 = 0;
  int64_t use_skinny = 1;
  // 16-bit decoder positions as ONE persistent cooperative kernel per position (decode_mega.cu) instead of ~33 dependent
  // launches; with fuse_topk the kernel also produces the log-softmax top-k of the 'max' beam search (logits never
  // stored).  OFF by default: measured on B200 it ties the per-operation path at 64 images x beam 3 (25.98 vs 26.14 ms per
  // call) and at batch 1 (5.89 vs 5.93 ms), and loses where the decoder has many rows (config 3, 1280 rows: 26.9 vs 19.3 ms;
  // config 4, 1536 rows: 199.6 vs 194.8 ms) -- its GEMM phases are mma.sync tiles, the per-operation path runs tcgen05
  // there.  DESIGN.md section 5 has the phase timeline and the ncu evidence.
  int64_t use_mega = 0, fuse_topk = 1;
  int64_t mega_search = 0;            // 1: the whole 'max' beam search in one launch (measured slower than one persistent launch per position: decode_mega.cu)
  unsigned* mega_bar = nullptr;       // grid-barrier counter of the persistent kernel
  int64_t mega_dbg_mode = 0;
  unsigned long long* mega_dbg = nullptr;   // option "mega_dbg": phase timestamps of the last persistent-kernel launch
  // Decoder-step LayerNorms folded into the consuming tcgen05 GEMM's A path.  Implemented and bit-identical to the
  // two-kernel path, but measured slower in the graph (14.4 us vs 6.5 us GEMM + 4.7 us LayerNorm kernel at M = 96: the
  // in-kernel normalisation sits on the critical path in front of the first MMA; 64-image call 29.5 vs 28.8 ms), so it
  // is off by default (profiles/r2_decoder_gemm_ln_on_load.txt).
  int64_t ln_on_load = 0;
  int64_t ln_fuse = 2;                // Swin norm1 / norm2 folded into the neighbouring tcgen05 GEMMs (16-bit modes); 2: also each stage's first norm1 (produced by the patch embedding / patch-merging reduction)
  // the decoder step is a chain of latency-bound kernels that fills a fraction of the machine: the batch is decoded as
  // up to kMaxDecodeGroups independent image groups on concurrent streams (parallel branches of the captured graph)
  static constexpr int kMaxDecodeGroups = 8;
  int64_t decode_groups = 0;          // 0 = automatic (by batch size)
  // device-side early exit: inside a captured call every decode step is the body of a CUDA-graph IF node whose condition
  // ("some beam was extended in the previous step", reference captioning_model.py:397) is set on the device
  int64_t early_exit = 4;                 // decode steps per conditional block (0 = off)
  cudaStream_t bstream[kMaxDecodeGroups + 1] = {};     // capture streams of the conditional bodies (one per decode group)
  cudaStream_t dstream[kMaxDecodeGroups] = {};
  cudaEvent_t d_fork = nullptr, d_join[kMaxDecodeGroups] = {};
  // xn_caption_host with pinned input: the images are copied chunk by chunk on a side stream inside the call's graph, so
  // the copy of Swin chunk c+1 overlaps the compute of chunk c
  static constexpr int kMaxCopyChunks = 64;
  cudaStream_t cstream = nullptr;
  cudaEvent_t c_fork = nullptr, c_ev[kMaxCopyChunks] = {};            // decoder-step linears on gemm_skinny.cu when rows <= 512 (16-bit modes)
  cudaStream_t gstream = nullptr;      // graphs are captured/replayed here (the caller's stream may be the legacy
  cudaEvent_t g_in = nullptr, g_out = nullptr;   // default stream, which cannot be captured); ordered with events
  void reset_side_streams() {
    for (int g = 0; g < kMaxDecodeGroups; ++g)
      if (dstream[g]) { cudaStreamDestroy(dstream[g]); cudaStreamCreateWithFlags(&dstream[g], cudaStreamNonBlocking); }
    if (cstream) { cudaStreamDestroy(cstream); cudaStreamCreateWithFlags(&cstream, cudaStreamNonBlocking); }
  }
  static void destroy_graph(CachedGraph& g) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
    g.exec = nullptr; g.graph = nullptr;
  }
  void drop_graphs() {
    for (auto& g : graphs) destroy_graph(g);
    graphs.clear();
  }
  // optional per-launch event timing of the tcgen05 GEMMs (bench.py roofline leg)
  int64_t profile = 0;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<double> prof_flops;
  size_t prof_used = 0;
  // "profile" = 2: every kernel launch of the library is bracketed by events and attributed to its launcher
  struct KernelSpan { const char* label; cudaEvent_t e0, e1; };
  std::vector<KernelSpan> spans;
  std::vector<cudaEvent_t> span_pool;
  cudaStream_t cur_st = nullptr;
  cudaEvent_t span_event() {
    if (span_pool.empty()) { cudaEvent_t e; cudaEventCreate(&e); return e; }
    cudaEvent_t e = span_pool.back(); span_pool.pop_back(); return e;
  }
  float* io_in = nullptr; size_t io_in_cap = 0;     // xn_caption_host staging
  // xn_caption_host_begin / _end: two slots (input staging, result staging, events) so that the host-to-device copy of
  // call i+1 runs on the copy stream while call i computes
  struct HostSlot { float* in = nullptr; size_t in_cap = 0; char* out = nullptr; size_t out_cap = 0;
                    cudaEvent_t ev_in = nullptr, ev_done = nullptr; bool pending = false; };
  HostSlot slots[2];
  int next_slot = 0;
  cudaStream_t hstream = nullptr;
  // set by a device-side check of the encoder output in the 16-bit modes: a non-finite value there means an fp16
  // intermediate overflowed somewhere in the backbone (xn_overflow_flag)
  int* flag_dev = nullptr;
  const float* staged_input = nullptr;   // set around a call whose device input already lives in a handle-owned buffer
  // xn_preprocess_rgb8: coefficient tables per (input size, output size), staging for the image and the first pass
  struct ResampleTable { int in_size, out_size, ksize; int* bounds; int* kk; };
  static constexpr size_t kMaxResampleTables = 64;
  std::vector<ResampleTable> rtables;
  uint8_t* pp_buf = nullptr; size_t pp_cap = 0;
  char* pp_host = nullptr; size_t pp_host_cap = 0;   // pinned staging of a preprocessing batch's item records + tables
  nvjpegHandle_t jpg = nullptr; nvjpegJpegState_t jpg_state = nullptr;     // xn_preprocess_jpeg_batch
  uint8_t* jpg_buf = nullptr; size_t jpg_cap = 0;                          // decoded RGB8 images of a batch
  cudaEvent_t pp_ev = nullptr;
  char* io_out = nullptr; size_t io_out_cap = 0;

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return h->fail(XN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define WS_CHECK()                                                                                 \
  do {                                                                                             \
    if (h->ws.overflow) return h->fail(XN_ERR_STATE, "workspace overflow: need %zu have %zu (%s:%d)", h->ws.off, h->ws.cap, __FILE__, __LINE__); \
  } while (0)
#define KL(n, expr)                                                                                \
  do {                                                                                             \
    cudaEvent_t pe0_ = nullptr, pe1_ = nullptr;                                                    \
    if (h->profile == 2) { pe0_ = h->span_event(); pe1_ = h->span_event(); cudaEventRecord(pe0_, h->cur_st); } \
    cudaError_t e_ = (expr);                                                                       \
    if (pe0_) { cudaEventRecord(pe1_, h->cur_st); h->spans.push_back({#expr, pe0_, pe1_}); }       \
    h->launches += (n);                                                                            \
    if (e_ != cudaSuccess) return h->fail(XN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

namespace {

// device staging of the call's input: captured graphs read it, so a reallocation invalidates them
int ensure_io_in(xn_handle* h, size_t bytes) {
  if (h->io_in_cap >= bytes) return 0;
  if (h->io_in) {
    CU(cudaDeviceSynchronize());
    h->drop_graphs();
    cudaFree(h->io_in);
    h->io_in = nullptr; h->io_in_cap = 0;
  }
  CU(cudaMalloc(&h->io_in, bytes));
  h->io_in_cap = bytes;
  return 0;
}

int ensure_ws(xn_handle* h, size_t bytes, cudaStream_t st) {
  bytes += 1 << 20;
  if (h->ws.cap < bytes) {
    if (h->ws.base) {
      CU(cudaStreamSynchronize(st));
      CU(cudaDeviceSynchronize());
      h->drop_graphs();
      CU(cudaFree(h->ws.base));
      h->ws.base = nullptr;
      h->ws.cap = 0;
    }
    CU(cudaMalloc(&h->ws.base, bytes));
    h->ws.cap = bytes;
  }
  h->ws.reset();
  return 0;
}

// ---- generic linear: y = act(x W^T / div + b) + res -------------------------------------------
int lin_f32(xn_handle* h, const float* x, long ldx, const LinW& w, const float* res, long ldr, float* y, long ldy,
            int M, int act, cudaStream_t st) {
  GemmArgs g{};
  g.A = x; g.lda = ldx; g.sA = 0;
  g.W = w.w; g.ldw = w.K; g.sW = 0;
  g.C = y; g.ldc = ldy; g.sC = 0;
  g.bias = w.b; g.res = res; g.ldr = ldr; g.sR = 0;
  g.M = M; g.N = w.N; g.K = w.K; g.batch = 1; g.div = 0.f; g.act = act; g.w_kn = 0;
  KL(1, launch_gemm_f32(g, st));
  return 0;
}
// every tcgen05 GEMM launch goes through here: with option "profile" = 1 it is bracketed by events (roofline accounting)
int tc_launch(xn_handle* h, const TcGemmArgs& g, cudaStream_t st) {
  if (h->profile == 1) {
    if (h->prof_used + 2 > h->prof_ev.size()) {
      for (int i = 0; i < 2; ++i) { cudaEvent_t e; CU(cudaEventCreate(&e)); h->prof_ev.push_back(e); }
    }
    CU(cudaEventRecord(h->prof_ev[h->prof_used], st));
    KL(1, launch_gemm_tc(g, st));
    CU(cudaEventRecord(h->prof_ev[h->prof_used + 1], st));
    h->prof_used += 2;
    h->prof_flops.push_back(2.0 * g.M * (double)g.N * g.K * (g.batch > 0 ? g.batch : 1));
    return 0;
  }
  KL(1, launch_gemm_tc(g, st));
  return 0;
}

int lin_tc(xn_handle* h, const void* x, long ldx, const LinW& w, const float* res, long ldr, float* yf, void* yb,
           long ldy, int M, int act, int fp16, cudaStream_t st, int w_static = 1, const float* a32 = nullptr, long lda32 = 0,
           const float* ln_g = nullptr, const float* ln_b = nullptr, const LnFuse* lf = nullptr, float div = 0.f) {
  TcGemmArgs g{};
  if (lf) { g.stats_out = lf->stats_out; g.x16_out = lf->x16_out; g.ldx16 = lf->ldx16; g.ln_stats = lf->ln_stats; g.ln_k = lf->ln_k; g.stats_zero = lf->stats_zero; }
  g.w_static = w_static;
  g.a32 = a32; g.lda32 = lda32; g.ln_g = ln_g; g.ln_b = ln_b;        // LayerNorm-on-load (x is then unused)
  g.A = x; g.lda = ldx; g.W = w.wb; g.ldw = w.K; g.Cf = yf; g.Cb = yb; g.ldc = ldy; g.fp16 = fp16;
  g.bias = w.b; g.res = res; g.ldr = ldr; g.M = M; g.N = w.N; g.K = w.K; g.div = div; g.act = act;
  return tc_launch(h, g, st);
}

// activation-type dispatch used by the templated Swin forward
template <typename T> struct ActOps;
template <> struct ActOps<float> {
  static int lin_act(xn_handle* h, const float* x, long ldx, const LinW& w, float* y, long ldy, int M, int act, cudaStream_t st, const LnFuse* = nullptr) {
    return lin_f32(h, x, ldx, w, nullptr, 0, y, ldy, M, act, st);
  }
  static int lin_res(xn_handle* h, const float* x, long ldx, const LinW& w, const float* res, long ldr, float* y, long ldy, int M, cudaStream_t st, const LnFuse* = nullptr) {
    return lin_f32(h, x, ldx, w, res, ldr, y, ldy, M, 0, st);
  }
  static cudaError_t attn(const float* qkv, const float* rpb, float* o, int B, int H, int C, int heads, int shift, cudaStream_t st) {
    return launch_window_attention<float>(qkv, rpb, o, B, H, C, heads, shift, st);
  }
};
template <> struct ActOps<bf16> {
  static int lin_act(xn_handle* h, const bf16* x, long ldx, const LinW& w, bf16* y, long ldy, int M, int act, cudaStream_t st, const LnFuse* lf = nullptr) {
    return lin_tc(h, x, ldx, w, nullptr, 0, nullptr, y, ldy, M, act, 0, st, 1, nullptr, 0, nullptr, nullptr, lf);
  }
  static int lin_res(xn_handle* h, const bf16* x, long ldx, const LinW& w, const float* res, long ldr, float* y, long ldy, int M, cudaStream_t st, const LnFuse* lf = nullptr) {
    return lin_tc(h, x, ldx, w, res, ldr, y, nullptr, ldy, M, 0, 0, st, 1, nullptr, 0, nullptr, nullptr, lf);
  }
  static cudaError_t attn(const bf16* qkv, const float* rpb, bf16* o, int B, int H, int C, int heads, int shift, cudaStream_t st) {
    if (g_attn_tc && window_attention_tc_supported(B, H, C, heads, shift)) return launch_window_attention_tc<bf16>(qkv, rpb, o, B, H, C, heads, shift, st);
    return launch_window_attention_mma<bf16>(qkv, rpb, o, B, H, C, heads, shift, st);
  }
};
template <> struct ActOps<f16> {
  static int lin_act(xn_handle* h, const f16* x, long ldx, const LinW& w, f16* y, long ldy, int M, int act, cudaStream_t st, const LnFuse* lf = nullptr) {
    return lin_tc(h, x, ldx, w, nullptr, 0, nullptr, y, ldy, M, act, 1, st, 1, nullptr, 0, nullptr, nullptr, lf);
  }
  static int lin_res(xn_handle* h, const f16* x, long ldx, const LinW& w, const float* res, long ldr, float* y, long ldy, int M, cudaStream_t st, const LnFuse* lf = nullptr) {
    return lin_tc(h, x, ldx, w, res, ldr, y, nullptr, ldy, M, 0, 1, st, 1, nullptr, 0, nullptr, nullptr, lf);
  }
  static cudaError_t attn(const f16* qkv, const float* rpb, f16* o, int B, int H, int C, int heads, int shift, cudaStream_t st) {
    if (g_attn_tc && window_attention_tc_supported(B, H, C, heads, shift)) return launch_window_attention_tc<f16>(qkv, rpb, o, B, H, C, heads, shift, st);
    return launch_window_attention_mma<f16>(qkv, rpb, o, B, H, C, heads, shift, st);
  }
};

// ---- Swin backbone --------------------------------------------------------------------------
size_t swin_ws_bytes(const xn_config& c, int Bc, int prec) {
  const size_t G = c.img_size / c.patch_size, L0 = G * G, C0 = c.embed_dim;
  const size_t act = prec == XN_PREC_FP32 ? 4 : 2;
  size_t tok = (size_t)Bc * L0 * C0;                 // elements of x at stage 0 (largest)
  size_t b = 0;
  b += 2 * tok * 4;                                   // x, x2 (fp32 residual stream, ping-pong over merges)
  b += tok * act;                                     // xn
  b += 3 * tok * act;                                 // qkv
  b += tok * act;                                     // attention out
  b += (size_t)(c.mlp_ratio * tok) * act + 4096;      // mlp hidden
  b += 2 * (size_t)Bc * L0 * 2 * 8;                     // per-row LayerNorm statistics (sum, sum of squares; 64-bit fixed point), norm1 and norm2
  return b + 17 * 256;
}

// the Swin LayerNorms under their own launcher name, so that the per-launcher profile (xn_profile_kernels) separates the
// bandwidth-bound big launches from the decoder's 512-wide row norms
template <typename T>
inline cudaError_t swin_layernorm(const float* x, long ldx, const float* g, const float* b, T* y, long ldy, long rows, int C, cudaStream_t st) {
  return launch_layernorm<T>(x, ldx, g, b, y, ldy, rows, C, st);
}

template <typename T>
int swin_forward_chunk(xn_handle* h, const float* img, int Bc, float* out, cudaStream_t st) {
  const xn_config& c = h->cfg;
  const int G = c.img_size / c.patch_size;
  const size_t tok0 = (size_t)Bc * G * G * c.embed_dim;
  float* x = h->ws.get<float>(tok0);
  float* x2 = h->ws.get<float>(tok0);
  T* xn = h->ws.get<T>(tok0);
  T* qkv = h->ws.get<T>(3 * tok0);
  T* ao = h->ws.get<T>(tok0);
  T* hid = h->ws.get<T>((size_t)(c.mlp_ratio * tok0) + 1024);
  float* stats = h->ws.get<float>((size_t)Bc * G * G * 4);      // (M x 2 64-bit fixed-point sums): norm1 statistics
  float* stats2 = h->ws.get<float>((size_t)Bc * G * G * 4);     // the same for norm2 (two buffers: see `alt` below)
  WS_CHECK();
  // 16-bit modes: norm1 / norm2 are folded into the GEMMs around them (option "ln_fuse").  The GEMM that produces the
  // residual stream (proj, fc2) also writes the raw rows in 16 bits into `xn` and their sum / sum of squares into `stats`;
  // the GEMM that consumes the normalised rows (fc1, the next block's qkv) reads those raw rows against weights that
  // carry gamma, are centred along K (the mean drops out) and scales its accumulator rows by 1/std.  Per block that
  // deletes two LayerNorm launches and the second read of the fp32 residual stream.
  const bool fuse = !std::is_same<T, float>::value && h->ln_fuse != 0;
  // The producers of a stage's FIRST norm1 are the patch embedding (stage 1) and the patch-merging reduction (later
  // stages): they emit the raw 16-bit rows and the row statistics too, so no LayerNorm launch is left in front of a qkv
  // GEMM (option "ln_fuse" >= 2, the default).  The raw rows of a stage's first block live in `ao` (the merge reads `xn`).
  const bool fuse_first = fuse && h->ln_fuse >= 2 && (c.embed_dim % 16 == 0);
  // `alt`: the statistics of norm1 (fc2 / patch embedding / merge -> qkv) and of norm2 (proj -> fc1) live in two buffers,
  // and each is cleared for its next producer by the GEMM that runs after its consumer (proj clears the norm1 buffer,
  // fc2 the norm2 buffer: TcGemmArgs::stats_zero) -- no memset node between the kernels of a block.
  const bool alt = fuse_first;
  float* st1 = stats;
  float* st2 = alt ? stats2 : stats;
  if (alt) CU(cudaMemsetAsync(st2, 0, (size_t)Bc * G * G * 2 * sizeof(long long), st));
  bool first_x16 = false;
  if (h->pe_wq) {
    first_x16 = fuse_first;
    if (!std::is_same<T, float>::value && h->pe_tc && patch_embed4_tc_supported(c.in_chans, c.img_size, c.embed_dim))
      KL(1, launch_patch_embed4_tc(img, h->pe_wq, h->pe_b, h->pe_g, h->pe_beta, x, Bc, c.in_chans, c.img_size, c.embed_dim, st,
                                   first_x16 ? (void*)ao : nullptr, std::is_same<T, f16>::value, first_x16 ? st1 : nullptr));
    else
    KL(1, launch_patch_embed4(img, h->pe_wq, h->pe_b, h->pe_g, h->pe_beta, x, Bc, c.in_chans, c.img_size, c.embed_dim, st,
                              first_x16 ? (void*)ao : nullptr, std::is_same<T, f16>::value, first_x16 ? st1 : nullptr));
  } else KL(1, launch_patch_embed(img, h->pe_w, h->pe_b, h->pe_g, h->pe_beta, x, Bc, c.in_chans, c.img_size, c.patch_size,
                                  c.embed_dim, st));
  for (size_t si = 0; si < h->stages.size(); ++si) {
    const SwinStageW& S = h->stages[si];
    const int C = S.C, H = S.H, M = Bc * H * H;
    bool have_x16 = false;                      // `xn` holds the raw 16-bit rows of x and `stats` their statistics
    for (size_t bi = 0; bi < S.blocks.size(); ++bi) {
      const SwinBlockW& W = S.blocks[bi];
      const int shift = (bi % 2 == 1 && H > c.window_size) ? c.window_size / 2 : 0;
      LnFuse consume1; consume1.ln_stats = st1; consume1.ln_k = C;
      LnFuse consume2; consume2.ln_stats = st2; consume2.ln_k = C;
      LnFuse produce2; produce2.stats_out = st2; produce2.x16_out = xn; produce2.ldx16 = C; produce2.stats_zero = alt ? st1 : nullptr;
      LnFuse produce1; produce1.stats_out = st1; produce1.x16_out = xn; produce1.ldx16 = C; produce1.stats_zero = alt ? st2 : nullptr;
      if (fuse && bi == 0 && first_x16) {
        if (int r = ActOps<T>::lin_act(h, ao, C, W.qkv_ln, qkv, 3 * C, M, 0, st, &consume1)) return r;
      } else if (fuse && have_x16) {
        if (int r = ActOps<T>::lin_act(h, xn, C, W.qkv_ln, qkv, 3 * C, M, 0, st, &consume1)) return r;
      } else {
        KL(1, swin_layernorm<T>(x, C, W.n1g, W.n1b, xn, C, M, C, st));
        if (int r = ActOps<T>::lin_act(h, xn, C, W.qkv, qkv, 3 * C, M, 0, st)) return r;
      }
      KL(1, ActOps<T>::attn(qkv, std::is_same<T, float>::value ? W.rpb : W.rpb_t, ao, Bc, H, C, S.heads, shift, st));
      if (fuse) {
        if (!alt) CU(cudaMemsetAsync(st2, 0, (size_t)M * 2 * sizeof(long long), st));
        if (int r = ActOps<T>::lin_res(h, ao, C, W.proj, x, C, x, C, M, st, &produce2)) return r;
        if (int r = ActOps<T>::lin_act(h, xn, C, W.fc1_ln, hid, W.fc1.N, M, 1, st, &consume2)) return r;
      } else {
        if (int r = ActOps<T>::lin_res(h, ao, C, W.proj, x, C, x, C, M, st)) return r;
        KL(1, swin_layernorm<T>(x, C, W.n2g, W.n2b, xn, C, M, C, st));
        if (int r = ActOps<T>::lin_act(h, xn, C, W.fc1, hid, W.fc1.N, M, 1, st)) return r;
      }
      have_x16 = fuse && bi + 1 < S.blocks.size();          // the last block of a stage feeds the merge / final norm (fp32)
      if (have_x16) {
        if (!alt) CU(cudaMemsetAsync(st1, 0, (size_t)M * 2 * sizeof(long long), st));
        if (int r = ActOps<T>::lin_res(h, hid, W.fc1.N, W.fc2, x, C, x, C, M, st, &produce1)) return r;
      } else if (alt) {
        LnFuse clear2; clear2.stats_zero = st2;              // not a producer, but the norm2 buffer still has to be cleared
        if (int r = ActOps<T>::lin_res(h, hid, W.fc1.N, W.fc2, x, C, x, C, M, st, &clear2)) return r;
      } else {
        if (int r = ActOps<T>::lin_res(h, hid, W.fc1.N, W.fc2, x, C, x, C, M, st)) return r;
      }
    }
    if (S.has_merge) {
      const int M4 = Bc * (H / 2) * (H / 2);
      KL(1, launch_merge_layernorm<T>(x, S.mg, S.mb, xn, Bc, H, C, st));     // xn viewed as (M4, 4C)
      first_x16 = fuse_first && ((2 * C) % 16 == 0);
      if (first_x16) {
        // the norm1 buffer is clear here: the stage's last proj cleared it and its last fc2 produced nothing
        LnFuse produce; produce.stats_out = st1; produce.x16_out = ao; produce.ldx16 = 2 * C;
        if (int r = ActOps<T>::lin_res(h, xn, 4 * C, S.red, nullptr, 0, x2, 2 * C, M4, st, &produce)) return r;
      } else if (int r = ActOps<T>::lin_res(h, xn, 4 * C, S.red, nullptr, 0, x2, 2 * C, M4, st)) return r;
      std::swap(x, x2);
    }
  }
  const SwinStageW& Sl = h->stages.back();
  KL(1, launch_layernorm<float>(x, Sl.C, h->swin_ng, h->swin_nb, out, Sl.C, (long)Bc * Sl.H * Sl.H, Sl.C, st));
  return 0;
}

// host_img != nullptr: `img` is device staging that has NOT been filled yet; every chunk's images are copied from the
// (pinned) host buffer on the handle's copy stream, all copies queued up front, and chunk c waits only for its own copy.
int swin_forward(xn_handle* h, const float* img, int B, float* out, cudaStream_t st, const float* host_img = nullptr, int host_chunk = 0) {
  const xn_config& c = h->cfg;
  const SwinStageW& Sl = h->stages.back();
  const size_t img_elems = (size_t)c.in_chans * c.img_size * c.img_size;
  const size_t out_elems = (size_t)Sl.H * Sl.H * Sl.C;
  int chunk = (int)std::max<int64_t>(1, h->swin_chunk);
  if (host_img && host_chunk > 0) chunk = std::min(chunk, host_chunk);
  const int n_chunks = (B + chunk - 1) / chunk;
  if (host_img) {
    if (n_chunks > xn_handle::kMaxCopyChunks) return h->fail(XN_ERR_ARG, "too many copy chunks (%d)", n_chunks);
    if (!h->cstream) {
      CU(cudaStreamCreateWithFlags(&h->cstream, cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&h->c_fork, cudaEventDisableTiming));
    }
    CU(cudaEventRecord(h->c_fork, st));
    CU(cudaStreamWaitEvent(h->cstream, h->c_fork, 0));
    for (int ci = 0; ci < n_chunks; ++ci) {
      const int b0 = ci * chunk, Bc = std::min(chunk, B - b0);
      if (!h->c_ev[ci]) CU(cudaEventCreateWithFlags(&h->c_ev[ci], cudaEventDisableTiming));
      CU(cudaMemcpyAsync(const_cast<float*>(img) + b0 * img_elems, host_img + b0 * img_elems, (size_t)Bc * img_elems * 4,
                         cudaMemcpyHostToDevice, h->cstream));
      CU(cudaEventRecord(h->c_ev[ci], h->cstream));
    }
  }
  for (int b0 = 0, ci = 0; b0 < B; b0 += chunk, ++ci) {
    const int Bc = std::min(chunk, B - b0);
    if (host_img) CU(cudaStreamWaitEvent(st, h->c_ev[ci], 0));
    h->ws.reset();
    int r = (h->precision == XN_PREC_BF16)   ? swin_forward_chunk<bf16>(h, img + b0 * img_elems, Bc, out + b0 * out_elems, st)
            : (h->precision == XN_PREC_FP16) ? swin_forward_chunk<f16>(h, img + b0 * img_elems, Bc, out + b0 * out_elems, st)
                                             : swin_forward_chunk<float>(h, img + b0 * img_elems, Bc, out + b0 * out_elems, st);
    if (r) return r;
  }
  return 0;
}

// ---- expansion encoder ----------------------------------------------------------------------
size_t enc_ws_bytes(const xn_config& c, int Bc) {
  const size_t N = c.enc_len, d = c.d_model, E = 0;
  (void)E;
  size_t e = 0;
  for (int i = 0; i < c.n_exp_groups; ++i) e += c.exp_groups[i];
  const size_t M = (size_t)Bc * N;
  size_t f = 0;
  f += M * d;                       // x0
  f += M * d * c.n_enc;             // xcat
  f += M * d;                       // xn
  f += M * 4 * d;                   // kabs
  f += 5 * (size_t)Bc * e * N;      // z, a_fw, b_fw, a_bw, b_bw
  f += 2 * (size_t)Bc * e * d;      // CA, CB
  f += 2 * M * d;                   // outA, outB
  f += M * c.ff;                    // ff hidden
  f += (size_t)Bc * c.n_exp_groups * 2 * N;   // group sums
  f += (size_t)Bc * ((N + 15) / 16) * 2 * e;  // per-slab column partials (transposed-score path)
  f += M * d;                                 // class_a / class_b projections transposed (16-bit, 2 d columns)
  f += M * d;                       // pre-norm output
  f += M * std::max<size_t>(d, c.feat_dim) + M * d * c.n_enc;   // 16-bit staging (counted at 4 B)
  return f * 4 + Bc * 4 + 40 * 256;
}

// T = float: everything fp32 (parity mode; CUDA-core GEMMs).  T = bf16/f16: the Linear layers run on tcgen05, the
// per-image expansion contractions (z, class, out) on the batched mma.sync kernel, all with 16-bit operands and fp32
// accumulation; normalisations, sums and the residual stream stay fp32.
template <typename T>
int enc_body_chunk(xn_handle* h, const float* feats, int Bc, const int* n_valid_dev, float* out, cudaStream_t st) {
  constexpr bool kF32 = std::is_same<T, float>::value;
  const xn_config& c = h->cfg;
  const int N = c.enc_len, d = c.d_model, E = h->n_exp_total, M = Bc * N, ne = c.n_enc;
  float* x0 = h->ws.get<float>((size_t)M * d);
  float* xcat = h->ws.get<float>((size_t)M * d * ne);
  T* xn = h->ws.get<T>((size_t)M * std::max(d, c.feat_dim));
  T* xcat16 = kF32 ? nullptr : h->ws.get<T>((size_t)M * d * ne);
  T* kabs = h->ws.get<T>((size_t)M * 4 * d);            // [key | class_a | class_b | selector] projections
  float* z = h->ws.get<float>((size_t)Bc * E * N);
  T* afw = h->ws.get<T>((size_t)Bc * E * N);
  T* bfw = h->ws.get<T>((size_t)Bc * E * N);
  T* abw = h->ws.get<T>((size_t)Bc * E * N);
  T* bbw = h->ws.get<T>((size_t)Bc * E * N);
  T* CA = h->ws.get<T>((size_t)Bc * E * d);
  T* CB = h->ws.get<T>((size_t)Bc * E * d);
  float* oA = h->ws.get<float>((size_t)M * d);
  float* oB = h->ws.get<float>((size_t)M * d);
  T* hid = h->ws.get<T>((size_t)M * c.ff);
  float* gs = h->ws.get<float>((size_t)Bc * c.n_exp_groups * 2 * N);
  float* colpart = h->ws.get<float>((size_t)Bc * ((N + 15) / 16) * 2 * E);
  T* abt = kF32 ? nullptr : h->ws.get<T>((size_t)M * 2 * d);       // [b][2 d][N]
  float* pre = h->ws.get<float>((size_t)M * d);
  WS_CHECK();
  const long ldc = (long)d * ne;
  const int fp16 = std::is_same<T, f16>::value;

  if (kF32) {
    if (int r = lin_f32(h, feats, c.feat_dim, h->input_linear, nullptr, 0, x0, d, M, 0, st)) return r;
  } else {
    KL(1, launch_cast<T>(feats, xn, (long)M * c.feat_dim, st));
    if (int r = lin_tc(h, xn, c.feat_dim, h->input_linear, nullptr, 0, x0, nullptr, d, M, 0, fp16, st)) return r;
  }
  for (int l = 0; l < ne; ++l) {
    const EncLayerW& W = h->enc[l];
    const float* xin = l == 0 ? x0 : xcat + (size_t)(l - 1) * d;
    const long ldi = l == 0 ? d : ldc;
    float* xout = xcat + (size_t)l * d;
    KL(1, launch_layernorm<T>(xin, ldi, W.n1g, W.n1b, xn, d, M, d, st));
    if (kF32) {
      const float* xn32 = reinterpret_cast<const float*>(xn);
      float* k32 = reinterpret_cast<float*>(kabs);
      float* afw32 = reinterpret_cast<float*>(afw); float* bfw32 = reinterpret_cast<float*>(bfw);
      float* abw32 = reinterpret_cast<float*>(abw); float* bbw32 = reinterpret_cast<float*>(bbw);
      float* CA32 = reinterpret_cast<float*>(CA); float* CB32 = reinterpret_cast<float*>(CB);
      if (int r = lin_f32(h, xn32, d, W.kabs, nullptr, 0, k32, 4 * d, M, 0, st)) return r;
      GemmArgs g{};
      // z[b] = Q (E x d) . key[b]^T / sqrt(d)            reference layers.py:52
      g.A = W.qexp; g.lda = d; g.sA = 0;
      g.W = k32; g.ldw = 4 * d; g.sW = (long)N * 4 * d; g.w_kn = 0;
      g.C = z; g.ldc = N; g.sC = (long)E * N;
      g.M = E; g.N = N; g.K = d; g.batch = Bc; g.div = sqrtf((float)d); g.act = 0;
      KL(1, launch_gemm_f32(g, st));
      KL(2, launch_static_exp_weights<float>(z, n_valid_dev, h->group_start_dev, c.n_exp_groups, afw32, bfw32, abw32, bbw32, gs,
                                             Bc, E, N, h->exp_chunk, st));
      // class_a = a_fw . A + bias_exp ; class_b = b_fw . B + bias_exp      layers.py:62-63
      for (int ab = 0; ab < 2; ++ab) {
        GemmArgs q{};
        q.A = ab ? bfw32 : afw32; q.lda = N; q.sA = (long)E * N;
        q.W = k32 + (size_t)(1 + ab) * d; q.ldw = 4 * d; q.sW = (long)N * 4 * d; q.w_kn = 1;
        q.C = ab ? CB32 : CA32; q.ldc = d; q.sC = (long)E * d;
        q.res = W.bexp; q.ldr = d; q.sR = 0;
        q.M = E; q.N = d; q.K = N; q.batch = Bc;
        KL(1, launch_gemm_f32(q, st));
      }
      // out = bw . class / n_groups                                          layers.py:82-83
      for (int ab = 0; ab < 2; ++ab) {
        GemmArgs q{};
        q.A = ab ? bbw32 : abw32; q.lda = E; q.sA = (long)N * E;
        q.W = ab ? CB32 : CA32; q.ldw = d; q.sW = (long)E * d; q.w_kn = 1;
        q.C = ab ? oB : oA; q.ldc = d; q.sC = (long)N * d;
        q.M = N; q.N = d; q.K = E; q.batch = Bc; q.div = (float)c.n_exp_groups;
        KL(1, launch_gemm_f32(q, st));
      }
      KL(1, launch_selector_mix<float>(xin, ldi, k32 + 3 * (size_t)d, 4 * d, oA, oB, d, xout, ldc, M, d, st));
    } else {
      if (int r = lin_tc(h, xn, d, W.kabs, nullptr, 0, nullptr, kabs, 4 * d, M, 0, fp16, st)) return r;
      if (!kF32 && h->se_tc >= 1 && h->se_t_ok) {
        // scores transposed, zT[b n][e] = key . q_e / sqrt(d): ONE linear-layer launch on tcgen05 over all images
        LinW qe{};
        qe.wb = W.qexp16; qe.N = E; qe.K = d;
        if (int r = lin_tc(h, kabs, 4 * d, qe, nullptr, 0, z, nullptr, E, M, 0, fp16, st, 1, nullptr, 0, nullptr, nullptr, nullptr, sqrtf((float)d))) return r;
        if constexpr (!kF32)
          KL(2, launch_static_exp_weights_t<T>(z, n_valid_dev, h->group_start_dev, c.n_exp_groups, afw, bfw, abw, bbw, gs, colpart, Bc, E, N, st));
      } else {
      Mma16Args g{};
      g.A = W.qexp16; g.lda = d; g.sA = 0;
      g.B = kabs; g.ldb = 4 * d; g.sB = (long)N * 4 * d; g.b_kn = 0;
      g.C = z; g.ldc = N; g.sC = (long)E * N;
      g.M = E; g.N = N; g.K = d; g.batch = Bc; g.scale = 1.0f / sqrtf((float)d);
      KL(1, (launch_gemm_mma16<T, float>(g, st)));
      KL(2, launch_static_exp_weights<T>(z, n_valid_dev, h->group_start_dev, c.n_exp_groups, afw, bfw, abw, bbw, gs, Bc, E, N,
                                         h->exp_chunk, st));
      }
      const bool se_tc2 = h->se_tc >= 2 && h->se_t_ok && (d % 64) == 0;
      if (se_tc2) {
        if constexpr (!kF32) {
          // class^T[b] (d x E) = A^T[b] (d x N) . fw[b]^T + bias_exp^T ;  out^T[b] (d x N) = class^T[b] . bw[b]^T / n_groups:
          // every operand K-major, batched over the images on tcgen05 (layers.py:62-63,82-83)
          KL(1, launch_transpose_ab<T>(kabs, 4 * d, d, abt, Bc, 2 * d, N, st));
          for (int ab = 0; ab < 2; ++ab) {
            TcGemmArgs q{};
            q.A = abt + (size_t)ab * d * N; q.lda = N; q.sA = (long)2 * d * N;
            q.W = ab ? bfw : afw; q.ldw = N; q.sW = (long)E * N;
            q.Cb = ab ? CB : CA; q.ldc = E; q.sC = (long)d * E; q.fp16 = fp16;
            q.res = W.bexpT; q.ldr = E; q.sR = 0;
            q.M = d; q.N = E; q.K = N; q.batch = Bc;
            if (int r = tc_launch(h, q, st)) return r;
          }
          for (int ab = 0; ab < 2; ++ab) {
            TcGemmArgs q{};
            q.A = ab ? CB : CA; q.lda = E; q.sA = (long)d * E;
            q.W = ab ? bbw : abw; q.ldw = E; q.sW = (long)N * E;
            q.Cf = ab ? oB : oA; q.ldc = N; q.sC = (long)d * N; q.fp16 = fp16;
            q.M = d; q.N = N; q.K = E; q.batch = Bc; q.div = (float)c.n_exp_groups;
            if (int r = tc_launch(h, q, st)) return r;
          }
          KL(1, launch_selector_mix_t<T>(xin, ldi, kabs + 3 * (size_t)d, 4 * d, oA, oB, xout, ldc, Bc, N, d, st));
        }
      } else {
      for (int ab = 0; ab < 2; ++ab) {
        Mma16Args q{};
        q.A = ab ? bfw : afw; q.lda = N; q.sA = (long)E * N;
        q.B = kabs + (size_t)(1 + ab) * d; q.ldb = 4 * d; q.sB = (long)N * 4 * d; q.b_kn = 1;
        q.C = ab ? CB : CA; q.ldc = d; q.sC = (long)E * d;
        q.res = W.bexp; q.ldr = d; q.sR = 0;
        q.M = E; q.N = d; q.K = N; q.batch = Bc; q.scale = 1.0f;
        KL(1, (launch_gemm_mma16<T, T>(q, st)));
      }
      for (int ab = 0; ab < 2; ++ab) {
        Mma16Args q{};
        q.A = ab ? bbw : abw; q.lda = E; q.sA = (long)N * E;
        q.B = ab ? CB : CA; q.ldb = d; q.sB = (long)E * d; q.b_kn = 1;
        q.C = ab ? oB : oA; q.ldc = d; q.sC = (long)N * d;
        q.M = N; q.N = d; q.K = E; q.batch = Bc; q.scale = 1.0f / (float)c.n_exp_groups;
        KL(1, (launch_gemm_mma16<T, float>(q, st)));
      }
      KL(1, launch_selector_mix<T>(xin, ldi, kabs + 3 * (size_t)d, 4 * d, oA, oB, d, xout, ldc, M, d, st));
      }
    }
    KL(1, launch_layernorm<T>(xout, ldc, W.n2g, W.n2b, xn, d, M, d, st));
    if (kF32) {
      if (int r = lin_f32(h, reinterpret_cast<const float*>(xn), d, W.ff1, nullptr, 0, reinterpret_cast<float*>(hid), c.ff, M, 2, st)) return r;
      if (int r = lin_f32(h, reinterpret_cast<const float*>(hid), c.ff, W.ff2, xout, ldc, xout, ldc, M, 0, st)) return r;
    } else {
      if (int r = lin_tc(h, xn, d, W.ff1, nullptr, 0, nullptr, hid, c.ff, M, 2, fp16, st)) return r;
      if (int r = lin_tc(h, hid, c.ff, W.ff2, xout, ldc, xout, nullptr, ldc, M, 0, fp16, st)) return r;
    }
  }
  if (kF32) {
    if (int r = lin_f32(h, xcat, ldc, h->enc_reduce, xcat + (size_t)(ne - 1) * d, ldc, pre, d, M, 0, st)) return r;
  } else {
    KL(1, launch_cast<T>(xcat, xcat16, (long)M * ldc, st));
    if (int r = lin_tc(h, xcat16, ldc, h->enc_reduce, xcat + (size_t)(ne - 1) * d, ldc, pre, nullptr, d, M, 0, fp16, st)) return r;
  }
  KL(1, launch_layernorm<float>(pre, d, h->enc_ng, h->enc_nb, out, d, M, d, st));
  return 0;
}

// encoder on features already on the device; pads handled via n_valid
int enc_body(xn_handle* h, const float* feats, int B, const int32_t* enc_pads_host, float* out, cudaStream_t st) {
  const xn_config& c = h->cfg;
  const int chunk = (int)std::max<int64_t>(1, h->enc_chunk);
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int Bc = std::min(chunk, B - b0);
    h->ws.reset();
    int* nv = nullptr;
    if (enc_pads_host) {
      bool any = false;
      std::vector<int> v(Bc);
      for (int i = 0; i < Bc; ++i) { v[i] = c.enc_len - enc_pads_host[b0 + i]; any |= enc_pads_host[b0 + i] != 0; }
      if (any) {
        nv = h->ws.get<int>(Bc);
        CU(cudaMemcpyAsync(nv, v.data(), Bc * sizeof(int), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));   // v is a stack-lifetime host buffer
      }
    }
    const float* fin = feats + (size_t)b0 * c.enc_len * c.feat_dim;
    float* fout = out + (size_t)b0 * c.enc_len * c.d_model;
    const int r = h->precision == XN_PREC_BF16   ? enc_body_chunk<bf16>(h, fin, Bc, nv, fout, st)
                  : h->precision == XN_PREC_FP16 ? enc_body_chunk<f16>(h, fin, Bc, nv, fout, st)
                                                 : enc_body_chunk<float>(h, fin, Bc, nv, fout, st);
    if (r) return r;
  }
  return 0;
}

// ---- decoder --------------------------------------------------------------------------------
struct DecBufs {
  DecState s;
  float *x0, *ycat, *q, *pre;
  void *xn, *att, *hid, *yn, *ycat16, *kv;      // fp32 in the parity mode, 16-bit operands otherwise
  void* e16;                                    // 16-bit copy of the encoder output (operand of the cross K/V projection)
  float* kpart;                                 // fp32 partial products of the split-K decoder-step linears: [K / 512][R][d]
  int R, P;
  // persistent-kernel path: when topk > 0 the step also leaves log-softmax top-k in topv / topi (topk_done is set by
  // dec_step when it did; otherwise the caller runs the separate log-softmax + top-k kernel on the logits)
  float* mega_scratch = nullptr; void* mega_act = nullptr;
  void* parts = nullptr; int topk = 0; float* topv = nullptr; int* topi = nullptr; bool topk_done = false;
};

size_t dec_ws_bytes(const xn_config& c, int R, int P, int n_images, bool own_logits) {
  const size_t d = c.d_model;
  size_t f = 0;
  f += (size_t)c.n_dec * P * R * 5 * d;
  f += (size_t)c.n_dec * P * R * 2 * c.num_exp_dec * P;
  f += (size_t)c.n_dec * P * R * c.num_exp_dec;
  f += (size_t)R * d * (6 + 2 * c.n_dec) + (size_t)R * c.ff;
  f += (size_t)R * d * (std::max<size_t>(c.ff, d * c.n_dec) / 512 + 1);          // split-K partial products
  f += (size_t)n_images * c.enc_len * c.n_dec * 2 * d;
  f += (size_t)n_images * c.enc_len * d;          // 16-bit copy of the encoder output
  if (own_logits) f += (size_t)R * c.vocab;
  return f * 4 + mega_parts_bytes(R, c.vocab) + mega_scratch_bytes(R) + mega_act_bytes(R, c.ff, c.n_dec) + 64 * 256;
}

int dec_alloc(xn_handle* h, DecBufs& D, int R, int P, int n_images) {
  const xn_config& c = h->cfg;
  const size_t d = c.d_model;
  D.R = R; D.P = P;
  D.s.cache = h->ws.get<float>((size_t)c.n_dec * P * R * 5 * d);
  D.s.cw = 5 * (int)d;
  D.s.fw = h->ws.get<float>((size_t)c.n_dec * P * R * 2 * c.num_exp_dec * P);
  D.s.qk = h->ws.get<float>((size_t)c.n_dec * P * R * c.num_exp_dec);
  D.s.anc = nullptr; D.s.P = P; D.s.R = R;
  D.x0 = h->ws.get<float>((size_t)R * d);
  D.ycat = h->ws.get<float>((size_t)R * d * c.n_dec);
  D.q = h->ws.get<float>((size_t)R * d);
  D.pre = h->ws.get<float>((size_t)R * d);
  D.xn = h->ws.get<float>((size_t)R * d);
  D.att = h->ws.get<float>((size_t)R * d);
  D.hid = h->ws.get<float>((size_t)R * c.ff);
  D.yn = h->ws.get<float>((size_t)R * d);
  D.ycat16 = h->ws.get<float>((size_t)R * d * c.n_dec);
  D.kpart = h->ws.get<float>((size_t)R * d * (std::max<size_t>(c.ff, d * c.n_dec) / 512 + 1));
  D.kv = h->ws.get<float>((size_t)n_images * c.enc_len * c.n_dec * 2 * d);
  // planned here, not taken from the arena at run time: the Swin / encoder chunks reset the arena offset in between
  D.e16 = h->ws.get<float>(((size_t)n_images * c.enc_len * d + 1) / 2);
  D.parts = h->ws.get<char>(mega_parts_bytes(R, c.vocab));
  D.mega_scratch = h->ws.get<float>(mega_scratch_bytes(R) / 4);
  D.mega_act = (c.ff % 512 == 0) ? h->ws.get<char>(mega_act_bytes(R, c.ff, c.n_dec)) : nullptr;
  D.topk = 0; D.topv = nullptr; D.topi = nullptr; D.topk_done = false;
  return 0;
}

// y = act(x W^T + b) + res with fp32 (parity) or tcgen05 16-bit operands; exactly one of yf / y16 is written
template <typename T>
int dec_lin(xn_handle* h, const T* x, long ldx, const LinW& w, const float* res, long ldr, float* yf, T* y16, long ldy,
            int M, int act, cudaStream_t st) {
  if (std::is_same<T, float>::value)
    return lin_f32(h, reinterpret_cast<const float*>(x), ldx, w, res, ldr, yf ? yf : reinterpret_cast<float*>(y16), ldy, M, act, st);
  return lin_tc(h, x, ldx, w, res, ldr, yf, yf ? nullptr : y16, ldy, M, act, std::is_same<T, f16>::value, st);
}

// decoder-step linear on the latency-oriented kernel (16-bit modes, M <= 512): optional LayerNorm / conversion of an
// fp32 A on load.  Returns 1 when the shape is not covered (the caller falls back to the tcgen05 path).
template <typename T>
int dec_lin_skinny(xn_handle* h, const T* a16, const float* a32, long lda, const float* ln_g, const float* ln_b, const LinW& w,
                   const float* res, long ldr, float* yf, T* y16, long ldy, int M, int act, cudaStream_t st) {
  SkinnyArgs g{};
  g.A16 = a16; g.A32 = a32; g.lda = lda; g.ln_g = ln_g; g.ln_b = ln_b;
  g.W = w.wb; g.ldw = w.K; g.bias = w.b; g.res = res; g.ldr = ldr; g.Cf = yf; g.Cb = yf ? nullptr : y16; g.ldc = ldy;
  g.M = M; g.N = w.N; g.K = w.K; g.act = act;
  // selected where it measured faster in-graph than the 128-row tcgen05 tiles: few rows, narrow outputs, no LayerNorm on
  // load.  (Extending it to one 128-row tile for the long-K projections -- 8.8 vs 12.8 us in isolation at M = 96 -- made
  // the two-group decode slower, 28.4 vs 27.9 ms per call: its 192 cluster CTAs crowd the other chain's kernels.)
  if (!h->use_skinny || M > 64 || w.N > 2048 || ln_g || !skinny_gemm_supported(g)) return 1;
  if (a32 && (w.K % 512)) return 1;
  if (h->profile == 1) {
    if (h->prof_used + 2 > h->prof_ev.size()) {
      for (int i = 0; i < 2; ++i) { cudaEvent_t e; CU(cudaEventCreate(&e)); h->prof_ev.push_back(e); }
    }
    CU(cudaEventRecord(h->prof_ev[h->prof_used], st));
    KL(1, launch_gemm_skinny<T>(g, st));
    CU(cudaEventRecord(h->prof_ev[h->prof_used + 1], st));
    h->prof_used += 2;
    h->prof_flops.push_back(-2.0 * M * (double)w.N * w.K);      // negative: counted separately from the tcgen05 launches
    return 0;
  }
  KL(1, launch_gemm_skinny<T>(g, st));
  return 0;
}

// cross K/V of all decoder layers, once per image (shared by the beams)
template <typename T>
int dec_project_kv(xn_handle* h, DecBufs& D, const float* enc_out, int n_images, cudaStream_t st) {
  const xn_config& c = h->cfg;
  const int d = c.d_model, M = n_images * c.enc_len;
  if (std::is_same<T, float>::value)
    return lin_f32(h, enc_out, d, h->kv_all, nullptr, 0, reinterpret_cast<float*>(D.kv), h->kv_all.N, M, 0, st);
  T* e16 = reinterpret_cast<T*>(D.e16);
  KL(1, launch_cast<T>(enc_out, e16, (long)M * d, st));
  return lin_tc(h, e16, d, h->kv_all, nullptr, 0, nullptr, D.kv, h->kv_all.N, M, 0, std::is_same<T, f16>::value, st);
}

// Arguments of the persistent decoder kernels (decode_mega.cu) for this handle / decode buffers.  Returns 1 when the
// geometry is not covered: the caller then runs the one-kernel-per-operation sequence.
int mega_build_args(xn_handle* h, DecBufs& D, int rows_per_image, const int* n_valid, const int* row_len, MegaArgs& a,
                    cudaStream_t st = nullptr) {
  const xn_config& c = h->cfg;
  if (!h->use_mega || (int)h->dec.size() > kMegaMaxLayers || c.ff % 512 || !D.mega_act) return 1;
  a = MegaArgs{};
  a.s = D.s; a.n_layers = c.n_dec; a.d = c.d_model; a.ff = c.ff; a.n_exp = c.num_exp_dec; a.heads = c.num_heads;
  a.n_keys = c.enc_len; a.vocab = c.vocab; a.R = D.R; a.rows_per_image = rows_per_image;
  for (int l = 0; l < c.n_dec; ++l) {
    const DecLayerW& W = h->dec[l];
    MegaLayer& m = a.L[l];
    m.n1g = W.n1g; m.n1b = W.n1b; m.n2g = W.n2g; m.n2b = W.n2b; m.n3g = W.n3g; m.n3b = W.n3b; m.qexp = W.qexp; m.bexp = W.bexp;
    m.w_dyn5 = W.dyn5.wp; m.w_wq = W.wq.wp; m.w_wo = W.wo.wp; m.w_ff1 = W.ff1.wp; m.w_ff2 = W.ff2.wp;
    if (!m.w_dyn5 || !m.w_wq || !m.w_wo || !m.w_ff1 || !m.w_ff2) return 1;
    m.b_dyn5 = W.dyn5.b; m.b_wq = W.wq.b; m.b_wo = W.wo.b; m.b_ff1 = W.ff1.b; m.b_ff2 = W.ff2.b;
    if (W.dyn5.N != 5 * c.d_model || W.dyn5.K != c.d_model || W.ff1.N != c.ff || W.ff2.K != c.ff) return 1;
  }
  a.n_valid = n_valid; a.row_len = row_len;
  a.emb = h->emb; a.pos = h->pos;
  a.x0 = D.x0; a.ycat = D.ycat; a.q = D.q; a.pre = D.pre;
  {
    const size_t slab = (size_t)((D.R + 31) / 32) * 32 * 520 * 2;          // one packed 16-bit slab (rows of 520 elements)
    char* base = reinterpret_cast<char*>(D.mega_act);
    a.xn = base; a.att = base + slab; a.hid = base + 2 * slab; a.ycat16 = base + (2 + (size_t)c.ff / 512) * slab;
  }
  a.kv = D.kv; a.ldkv = (long)c.n_dec * 2 * c.d_model;
  a.w_reduce = h->dec_reduce.wp; a.b_reduce = h->dec_reduce.b; a.ng = h->dec_ng; a.nb = h->dec_nb;
  a.w_vocab = h->vocab.wp; a.b_vocab = h->vocab.b;
  if (!a.w_reduce || !a.w_vocab) return 1;
  a.parts = D.parts;
  // fixed split factors (not chosen by row count): the summation order, hence every bit of the result, is independent of the batch
  a.ksplit_ff2 = c.ff / 512; a.ksplit_red = c.n_dec; a.scratch = D.mega_scratch;
  // decode groups run on their own streams: each gets its own barrier words, and one CTA per SM so that two groups' kernels
  // are resident side by side (the phases are latency-bound: two interleaved chains fill the gaps)
  int slot = 0;
  for (int g = 0; g < xn_handle::kMaxDecodeGroups; ++g) if (st && st == h->dstream[g]) slot = g + 1;
  for (int g = 0; g <= xn_handle::kMaxDecodeGroups; ++g) if (st && st == h->bstream[g]) slot = g;
  a.bar = h->mega_bar + (size_t)slot * (kMegaBarBytes / sizeof(unsigned));
  a.max_ctas_per_sm = slot > 0 ? 1 : 2;
  a.dbg = slot <= 1 ? h->mega_dbg : nullptr; a.dbg_mode = (int)h->mega_dbg_mode;
  if (h->dec_reduce.K != c.d_model * c.n_dec || h->dec_reduce.N != c.d_model || h->vocab.K != c.d_model) return 1;
  return 0;
}

// one decoder position for all rows as ONE persistent kernel (16-bit modes)
template <typename T>
int dec_step_mega(xn_handle* h, DecBufs& D, int p, const int64_t* tok64, const int* tok32, long tok_stride,
                  int rows_per_image, const int* n_valid, const int* row_len, float* logits, long ldl, cudaStream_t st) {
  MegaArgs a;
  if (mega_build_args(h, D, rows_per_image, n_valid, row_len, a, st)) return 1;
  a.p = p; a.tok64 = tok64; a.tok32 = tok32; a.tok_stride = tok_stride;
  a.logits = logits; a.ldl = ldl;
  const bool fuse = h->fuse_topk && D.topk > 0 && D.topv && D.topi && D.parts;
  a.topk = fuse ? D.topk : 0; a.top_val = D.topv; a.top_idx = D.topi;
  if (!mega_supported(a)) return 1;
  KL(1, launch_dec_step_mega<T>(a, st));
  D.topk_done = fuse;
  return 0;
}

// one decoder position for all rows -> logits (R, V) at `logits` with row stride ldl
template <typename T>
int dec_step_t(xn_handle* h, DecBufs& D, int p, const int64_t* tok64, const int* tok32, long tok_stride,
               int rows_per_image, const int* n_valid, const int* row_len, float* logits, long ldl, cudaStream_t st) {
  const xn_config& c = h->cfg;
  const int d = c.d_model, R = D.R, nd = c.n_dec;
  const long ldc = (long)d * nd, ldkv = (long)nd * 2 * d;
  T* xn = reinterpret_cast<T*>(D.xn);
  T* att = reinterpret_cast<T*>(D.att);
  T* hid = reinterpret_cast<T*>(D.hid);
  T* yn = reinterpret_cast<T*>(D.yn);
  constexpr bool k16 = !std::is_same<T, float>::value;
  D.topk_done = false;
  if constexpr (k16) {
    const int rm = dec_step_mega<T>(h, D, p, tok64, tok32, tok_stride, rows_per_image, n_valid, row_len, logits, ldl, st);
    if (rm <= 0) return rm;
  }
  // 16-bit modes with a few hundred rows: the latency-oriented kernel (LayerNorm / conversion of A fused into its load)
  auto skinny = [&](const T* a16, const float* a32, long lda, const float* g, const float* b, const LinW& w, const float* res,
                    long ldr, float* yf, T* y16, long ldy, int act) -> int {
    if constexpr (k16) return dec_lin_skinny<T>(h, a16, a32, lda, g, b, w, res, ldr, yf, y16, ldy, R, act, st);
    else return 1;
  };
  // Split-K (option "dec_splitk", 16-bit modes, tcgen05 path): a decoder-step linear with K >= 1024 onto d_model streams
  // its whole K through 8 CTAs (16 us at 96 rows for 0.2 GFLOP); as K / 512 slices (batched mode of the tcgen05 GEMM)
  // it runs on 4x the CTAs, and the partial products are summed -- in slice order -- by the LayerNorm launch that
  // follows.  Applied at every row count of the tcgen05 path, so a caption does not depend on its batch.
  // Measured in the graph (64 images, two decode groups): 5.20 ms of decode with, 5.00 - 5.36 ms without, 25.03 vs 25.08 ms
  // per call -- a tie (the 16 us are a cold-cache ncu figure; in the graph the weights are L2 hits), so it is off by default.
  int pend_parts = 0;
  const float* pend_bias = nullptr;
  auto splitk_ok = [&](const LinW& w) {
    return k16 && h->dec_splitk && !h->ln_on_load && (R > 64 || !h->use_skinny) && w.N == d && d % 128 == 0 && d <= 1024 && w.K >= 1024 && w.K % 512 == 0 &&
           w.K / 512 <= (int)(std::max<size_t>(c.ff, (size_t)d * c.n_dec) / 512 + 1);
  };
  auto lin_splitk = [&](const T* a, long lda, const LinW& w) -> int {
    TcGemmArgs q{};
    const int ns = w.K / 512;
    q.A = a; q.lda = lda; q.sA = 512; q.W = w.wb; q.ldw = w.K; q.sW = 512;
    q.Cf = D.kpart; q.ldc = d; q.sC = (long)R * d; q.fp16 = std::is_same<T, f16>::value;
    q.M = R; q.N = w.N; q.K = 512; q.batch = ns; q.w_static = 1;
    if (int r = tc_launch(h, q, st)) return r;
    pend_parts = ns; pend_bias = w.b;
    return 0;
  };
  const bool fuse_ln = d == 512;            // embedding + norm_1 of layer 0, dynamic expansion + norm_2: one kernel each
  if (fuse_ln) KL(1, launch_embed_ln<T>(tok64, tok32, tok_stride, p, h->emb, h->pos, D.x0, d, h->dec[0].n1g, h->dec[0].n1b, xn, d, R, d, st));
  else KL(1, launch_embed(tok64, tok32, tok_stride, p, h->emb, h->pos, D.x0, d, R, d, st));
  for (int l = 0; l < nd; ++l) {
    const DecLayerW& W = h->dec[l];
    const float* xin = l == 0 ? D.x0 : D.ycat + (size_t)(l - 1) * d;
    const long ldi = l == 0 ? d : ldc;
    float* xout = D.ycat + (size_t)l * d;
    float* crow = D.s.cache + (((size_t)l * D.P + p) * R) * D.s.cw;
    // 16-bit A: the skinny kernel where selected (few rows), else the tcgen05 kernel
    auto lin16 = [&](const T* a, long lda, const LinW& w, const float* res, long ldr, float* yf, T* y16, long ldy, int act) -> int {
      const int rs = skinny(a, nullptr, lda, nullptr, nullptr, w, res, ldr, yf, y16, ldy, act);
      if (rs <= 0) return rs;
      return dec_lin<T>(h, a, lda, w, res, ldr, yf, y16, ldy, R, act, st);
    };
    // LayerNorm + linear with the norm folded into the GEMM's A path (tcgen05, K = 512, one tile per CTA, more than 64 rows)
    auto ln_lin = [&](const float* x32, long ldx32, const float* g_, const float* b_, const LinW& w, float* yf, T* y16, long ldy, int act) -> int {
      if constexpr (k16) {
        if (pend_parts > 0) {
          // the previous layer's ff2 ran split along K: its partial products, bias and residual are summed here (into the
          // residual stream row x32 itself) by the LayerNorm launch that follows it anyway
          KL(1, launch_layernorm_sum<T>(D.kpart, pend_parts, (long)R * d, d, pend_bias, x32, ldx32, const_cast<float*>(x32), ldx32, g_, b_, xn, d, R, d, st));
          pend_parts = 0;
          return lin16(xn, d, w, nullptr, 0, yf, y16, ldy, act);
        }
        if (h->ln_on_load && R > 64 && tc_gemm_ln_supported(R, w.N, w.K))
          return lin_tc(h, nullptr, 0, w, nullptr, 0, yf, yf ? nullptr : y16, ldy, R, act, std::is_same<T, f16>::value, st, 1, x32, ldx32, g_, b_);
      }
      KL(1, launch_layernorm<T>(x32, ldx32, g_, b_, xn, d, R, d, st));
      return lin16(xn, d, w, nullptr, 0, yf, y16, ldy, act);
    };
    if (fuse_ln && l == 0) { if (int r = lin16(xn, d, W.dyn5, nullptr, 0, crow, nullptr, D.s.cw, 0)) return r; }
    else if (int r = ln_lin(xin, ldi, W.n1g, W.n1b, W.dyn5, crow, nullptr, D.s.cw, 0)) return r;
    if (fuse_ln) {
      KL(1, launch_dyn_exp_step<T>(D.s, l, p, W.qexp, W.bexp, c.num_exp_dec, row_len, xin, ldi, xout, ldc, d, W.n2g, W.n2b, xn, d, st));
    } else {
      KL(1, launch_dyn_exp_step<T>(D.s, l, p, W.qexp, W.bexp, c.num_exp_dec, row_len, xin, ldi, xout, ldc, d, nullptr, nullptr, (T*)nullptr, 0, st));
      KL(1, launch_layernorm<T>(xout, ldc, W.n2g, W.n2b, xn, d, R, d, st));
    }
    if (int r = lin16(xn, d, W.wq, nullptr, 0, D.q, nullptr, d, 0)) return r;
    KL(1, (launch_cross_attn_step<T, T>(D.q, d, reinterpret_cast<const T*>(D.kv), ldkv, l * 2 * d, l * 2 * d + d, att, d, R,
                                        rows_per_image, c.enc_len, c.num_heads, d / c.num_heads, n_valid, row_len, p, st)));
    if (int r = lin16(att, d, W.wo, xout, ldc, xout, nullptr, ldc, 0)) return r;
    if (int r = ln_lin(xout, ldc, W.n3g, W.n3b, W.ff1, nullptr, hid, c.ff, 2)) return r;
    if (l + 1 < nd && splitk_ok(W.ff2)) { if (int r = lin_splitk(hid, c.ff, W.ff2)) return r; }     // summed by layer l + 1's norm_1 launch
    else if (int r = lin16(hid, c.ff, W.ff2, xout, ldc, xout, nullptr, ldc, 0)) return r;
  }
  // reduce group: fp32 concatenation, converted on load by the skinny kernel where selected
  int rs = skinny(nullptr, D.ycat, ldc, nullptr, nullptr, h->dec_reduce, D.ycat + (size_t)(nd - 1) * d, ldc, D.pre, nullptr, d, 0);
  if (rs < 0) return rs;
  if (rs > 0) {
    const T* ycat_in = reinterpret_cast<const T*>(D.ycat);
    if (k16) {
      KL(1, launch_cast<T>(D.ycat, reinterpret_cast<T*>(D.ycat16), (long)R * ldc, st));
      ycat_in = reinterpret_cast<const T*>(D.ycat16);
    }
    if (splitk_ok(h->dec_reduce)) {
      if constexpr (k16) {
        if (int r = lin_splitk(ycat_in, ldc, h->dec_reduce)) return r;
        KL(1, launch_layernorm_sum<T>(D.kpart, pend_parts, (long)R * d, d, pend_bias, D.ycat + (size_t)(nd - 1) * d, ldc, D.pre, d, h->dec_ng, h->dec_nb,
                                      yn, d, R, d, st));
        pend_parts = 0;
        return dec_lin<T>(h, yn, d, h->vocab, nullptr, 0, logits, nullptr, ldl, R, 0, st);
      }
    }
    if (int r = dec_lin<T>(h, ycat_in, ldc, h->dec_reduce, D.ycat + (size_t)(nd - 1) * d, ldc, D.pre, nullptr, d, R, 0, st)) return r;
  }
  if constexpr (k16) {
    if (h->ln_on_load && R > 64 && d == 512 && tc_gemm_ln_supported(R, h->vocab.N, h->vocab.K))
      return lin_tc(h, nullptr, 0, h->vocab, nullptr, 0, logits, nullptr, ldl, R, 0, std::is_same<T, f16>::value, st, 1, D.pre, d,
                    h->dec_ng, h->dec_nb);
  }
  KL(1, launch_layernorm<T>(D.pre, d, h->dec_ng, h->dec_nb, yn, d, R, d, st));
  if (int r = dec_lin<T>(h, yn, d, h->vocab, nullptr, 0, logits, nullptr, ldl, R, 0, st)) return r;
  return 0;
}

int dec_step(xn_handle* h, DecBufs& D, int p, const int64_t* tok64, const int* tok32, long tok_stride,
             int rows_per_image, const int* n_valid, const int* row_len, float* logits, long ldl, cudaStream_t st) {
  if (h->precision == XN_PREC_BF16) return dec_step_t<bf16>(h, D, p, tok64, tok32, tok_stride, rows_per_image, n_valid, row_len, logits, ldl, st);
  if (h->precision == XN_PREC_FP16) return dec_step_t<f16>(h, D, p, tok64, tok32, tok_stride, rows_per_image, n_valid, row_len, logits, ldl, st);
  return dec_step_t<float>(h, D, p, tok64, tok32, tok_stride, rows_per_image, n_valid, row_len, logits, ldl, st);
}
int dec_project(xn_handle* h, DecBufs& D, const float* enc_out, int n_images, cudaStream_t st) {
  if (h->precision == XN_PREC_BF16) return dec_project_kv<bf16>(h, D, enc_out, n_images, st);
  if (h->precision == XN_PREC_FP16) return dec_project_kv<f16>(h, D, enc_out, n_images, st);
  return dec_project_kv<float>(h, D, enc_out, n_images, st);
}

int upload_ints(xn_handle* h, const std::vector<int>& v, int** dev, cudaStream_t st) {
  *dev = h->ws.get<int>(v.size());
  CU(cudaMemcpyAsync(*dev, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

// Runs `body(stream)` (kernel launches only: no allocation, no host sync) either directly on the caller's stream or,
// from the second identical call on, as a CUDA graph captured once and replayed on an internal stream that is ordered
// with the caller's stream by events.
template <typename F>
int run_graphed(xn_handle* h, const xn_handle::GraphKey& key, bool allow_graph, cudaStream_t user_st, F&& body,
                const void* host_src = nullptr, size_t host_bytes = 0) {
  if (!(h->use_graph && allow_graph && !h->profile)) return body(user_st);
  if (!h->gstream) {
    CU(cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->g_in, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->g_out, cudaEventDisableTiming));
  }
  xn_handle::CachedGraph* g = nullptr;
  for (size_t i = 0; i < h->graphs.size(); ++i)
    if (h->graphs[i].key == key) {                  // most recently used entry lives at the back
      if (i + 1 != h->graphs.size()) std::rotate(h->graphs.begin() + i, h->graphs.begin() + i + 1, h->graphs.end());
      g = &h->graphs.back();
      break;
    }
  if (!g) {                                         // first sight: run eagerly, remember the shape
    if (h->graphs.size() >= xn_handle::kMaxGraphs) {          // evict the least recently used entry only
      xn_handle::destroy_graph(h->graphs.front());
      h->graphs.erase(h->graphs.begin());
    }
    xn_handle::CachedGraph fresh{};
    fresh.key = key;
    h->graphs.push_back(fresh);
    return body(user_st);
  }
  cudaStream_t st = h->gstream;                     // hand over from the caller's stream to the graph stream
  CU(cudaEventRecord(h->g_in, user_st));
  CU(cudaStreamWaitEvent(st, h->g_in, 0));
  if (!g->exec) {                                   // second identical call: capture once, then replay from now on
    const int64_t l0 = h->launches;
    CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    const int rr = body(st);
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rr || ce != cudaSuccess) {
      // a failed capture may leave forked side streams in an invalidated capture: recreate them before the next call
      if (graph) cudaGraphDestroy(graph);
      (void)cudaGetLastError();
      h->reset_side_streams();
      h->launches = l0;
      if (rr) return rr;
      return h->fail(XN_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
    }
    g->launches = h->launches - l0;
    h->launches = l0;
    cudaError_t ie = cudaGraphInstantiate(&g->exec, graph, 0);
    if (ie != cudaSuccess) { cudaGraphDestroy(graph); g->exec = nullptr; return h->fail(XN_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ie)); }
    g->graph = graph;
    g->host_base = reinterpret_cast<const char*>(host_src);
    g->h2d.clear();
    if (host_src) {                                 // remember the copy nodes that read the caller's host buffer
      size_t n_nodes = 0;
      CU(cudaGraphGetNodes(graph, nullptr, &n_nodes));
      std::vector<cudaGraphNode_t> nodes(n_nodes);
      if (n_nodes) CU(cudaGraphGetNodes(graph, nodes.data(), &n_nodes));
      const char* hb = reinterpret_cast<const char*>(host_src);
      for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType ty;
        if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess) { (void)cudaGetLastError(); continue; }      // e.g. conditional nodes
        if (ty != cudaGraphNodeTypeMemcpy) continue;
        cudaMemcpy3DParms mp{};
        CU(cudaGraphMemcpyNodeGetParams(nd, &mp));
        const char* src = reinterpret_cast<const char*>(mp.srcPtr.ptr);
        if (!src || src < hb || src >= hb + host_bytes) continue;
        const size_t bytes = mp.extent.width * std::max<size_t>(1, mp.extent.height) * std::max<size_t>(1, mp.extent.depth);
        g->h2d.push_back({nd, mp.dstPtr.ptr, (size_t)(src - hb), bytes});
      }
    }
  } else if (host_src && g->host_base != reinterpret_cast<const char*>(host_src)) {
    const char* hb = reinterpret_cast<const char*>(host_src);
    for (auto& n : g->h2d)
      CU(cudaGraphExecMemcpyNodeSetParams1D(g->exec, n.node, n.dst, hb + n.src_off, n.bytes, cudaMemcpyHostToDevice));
    g->host_base = hb;
  }
  CU(cudaGraphLaunch(g->exec, st));
  h->launches += g->launches;
  CU(cudaEventRecord(h->g_out, st));                // and back
  CU(cudaStreamWaitEvent(user_st, h->g_out, 0));
  return 0;
}

// beam search over an encoder output: arena plan (host arithmetic only), kernel sequence, result copies
struct BeamPlan {
  DecBufs D;
  float *logits, *topv; int* topi;
  BeamBufs bb;
  int* nv = nullptr;
  int32_t *r_tok, *r_len; float* r_lp;
};

int beam_plan(xn_handle* h, BeamPlan& P, int B, const int32_t* enc_pads_host, int beam, int L, int how_many, cudaStream_t st,
              bool reset_ws) {
  const xn_config& c = h->cfg;
  if (beam < 1 || beam > 8 || how_many > beam || how_many < 1) return h->fail(XN_ERR_ARG, "requested output per sequence must be lower than beam width (beam<=8)");
  if (L < 2 || L > c.max_seq_len || L > 128) return h->fail(XN_ERR_ARG, "max_seq_len %d outside [2, %d]", L, std::min(c.max_seq_len, 128));
  if (beam > c.vocab) return h->fail(XN_ERR_ARG, "beam > vocab");
  const int R = B * beam;
  if (reset_ws) h->ws.reset();
  dec_alloc(h, P.D, R, L, B);
  P.logits = h->ws.get<float>((size_t)R * c.vocab);
  P.topv = h->ws.get<float>((size_t)R * beam);
  P.topi = h->ws.get<int>((size_t)R * beam);
  for (int s = 0; s < 2; ++s) {
    P.bb.tokens[s] = h->ws.get<int>((size_t)R * L);
    P.bb.lps[s] = h->ws.get<float>((size_t)R * L);
    P.bb.len[s] = h->ws.get<int>(R);
    P.bb.anc[s] = h->ws.get<int>((size_t)R * L);
    P.bb.cum[s] = h->ws.get<float>(R);
    P.bb.eos[s] = h->ws.get<int>(R);
  }
  P.bb.all_done = h->ws.get<int>(1);
  P.bb.grew = h->ws.get<int>(L);
  P.bb.final_src = h->ws.get<int>(1);
  // results land in arena buffers (stable addresses -> graph-capturable), then are copied to the caller
  P.r_tok = h->ws.get<int32_t>((size_t)B * how_many * L);
  P.r_len = h->ws.get<int32_t>((size_t)B * how_many);
  P.r_lp = h->ws.get<float>((size_t)B * how_many * L);
  P.nv = nullptr;
  if (enc_pads_host) {
    std::vector<int> v(B);
    bool any = false;
    for (int i = 0; i < B; ++i) { v[i] = c.enc_len - enc_pads_host[i]; any |= enc_pads_host[i] != 0; }
    if (any) if (int r = upload_ints(h, v, &P.nv, st)) return r;
  }
  WS_CHECK();
  return 0;
}

// candidate selection of one step: top-k of the log-probabilities ('max'), or k draws without replacement ('sample',
// reference captioning_model.py:131-133,168-170)
struct SampleOpt { bool on = false; uint64_t seed = 0; };

int beam_run(xn_handle* h, BeamPlan& P, const float* enc_out, int B, int beam, int L, int how_many, int sos, int eos,
             cudaStream_t st, SampleOpt smp = SampleOpt()) {
  const xn_config& c = h->cfg;
  const int R = B * beam;
  DecBufs& D = P.D;
  BeamBufs& bb = P.bb;
  // cross K/V of all decoder layers, once per image (shared by the beams)
  if (int r = dec_project(h, D, enc_out, B, st)) return r;
  KL(1, launch_beam_init(bb, B, beam, L, sos, st));
  // 16-bit 'max' search on the covered geometry: ALL time steps in one persistent launch (decode_mega.cu); the early
  // `break` of the reference is a branch inside the kernel
  if (!smp.on && h->precision != XN_PREC_FP32 && h->mega_search && h->fuse_topk) {
    MegaArgs a;
    D.s.anc = bb.anc[0];
    if (mega_build_args(h, D, beam, P.nv, nullptr, a, st) == 0) {
      a.p = 0; a.tok32 = bb.tokens[0]; a.tok_stride = L;
      a.topk = beam; a.top_val = P.topv; a.top_idx = P.topi;
      MegaSearch q{};
      q.bb = bb; q.beam = beam; q.L = L; q.eos = eos; q.how_many = how_many; q.early_exit = h->early_exit != 0;
      q.r_tok = P.r_tok; q.r_len = P.r_len; q.r_lp = P.r_lp;
      if (mega_supported(a)) {
        if (h->precision == XN_PREC_FP16) KL(1, launch_dec_search_mega<f16>(a, q, st));
        else KL(1, launch_dec_search_mega<bf16>(a, q, st));
        return 0;
      }
    }
  }
  int src = 0;
  // step 0: every beam row decodes [SOS]
  D.s.anc = bb.anc[0];
  D.topk = smp.on ? 0 : beam; D.topv = P.topv; D.topi = P.topi;      // 'max': the persistent kernel fuses log-softmax + top-k
  if (int r = dec_step(h, D, 0, nullptr, bb.tokens[0], L, beam, P.nv, nullptr, P.logits, c.vocab, st)) return r;
  if (smp.on) KL(1, launch_gumbel_topk(P.logits, c.vocab, R, c.vocab, beam, smp.seed, 1, P.topv, P.topi, st));
  else if (!D.topk_done) KL(1, launch_logsoftmax_topk(P.logits, c.vocab, R, c.vocab, beam, P.topv, P.topi, nullptr, 0, 0, st));
  KL(1, launch_beam_first(bb, P.topv, P.topi, B, beam, L, eos, st));
  // one loop iteration of the reference (captioning_model.py:295-397) on stream `s2`
  auto one_step = [&](int t, cudaStream_t s2) -> int {
    D.s.anc = bb.anc[src];
    if (int r = dec_step(h, D, t - 1, nullptr, bb.tokens[src], L, beam, P.nv, nullptr, P.logits, c.vocab, s2)) return r;
    if (smp.on) KL(1, launch_gumbel_topk(P.logits, c.vocab, R, c.vocab, beam, smp.seed, t, P.topv, P.topi, s2));
    else if (!D.topk_done) KL(1, launch_logsoftmax_topk(P.logits, c.vocab, R, c.vocab, beam, P.topv, P.topi, nullptr, 0, 0, s2));
    KL(1, launch_beam_step(bb, src, P.topv, P.topi, B, beam, L, t, eos, s2));
    src ^= 1;
    return 0;
  };
  // Early termination (captioning_model.py:397: the reference breaks when no beam was extended).  Without a host in the
  // loop the break is a device-side condition: while the call is being CAPTURED, every step becomes the body graph of an
  // IF node; the step's last kernel sets the next node's condition to "some beam grew", a skipped step leaves the rest at
  // their default 0.  The finaliser reads the ping-pong index the last executed step left behind.  Eager (uncaptured)
  // calls run all steps: the result is the same, finished beams only re-append their 0.0 candidate.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  CU(cudaStreamIsCapturing(st, &cap));
  // An IF node costs ~18 us of graph scheduling (measured: 36 nodes added 0.7 ms to a 64-image call), so the steps are
  // grouped: `early_exit` = steps per conditional block (default 4; 0 = off); the first block runs unconditionally.
  const int blk = (int)std::max<int64_t>(0, h->early_exit);
  const bool conditional = blk > 0 && cap == cudaStreamCaptureStatusActive && L > 2 + blk;
  if (!conditional) {
    for (int t = 2; t < L; ++t)
      if (int r = one_step(t, st)) return r;
  } else {
    int slot = 0;
    for (int g = 0; g < xn_handle::kMaxDecodeGroups; ++g) if (st == h->dstream[g]) slot = g + 1;
    if (!h->bstream[slot]) CU(cudaStreamCreateWithFlags(&h->bstream[slot], cudaStreamNonBlocking));
    cudaStream_t sb = h->bstream[slot];
    cudaStreamCaptureStatus cs;
    unsigned long long cid = 0;
    cudaGraph_t graph = nullptr;
    const cudaGraphNode_t* deps = nullptr;
    size_t ndeps = 0;
    int t = 2;
    for (; t < 2 + blk; ++t)                               // first block: always runs
      if (int r = one_step(t, st)) return r;
    cudaGraphConditionalHandle hnd = 0;
    CU(cudaStreamGetCaptureInfo_v2(st, &cs, &cid, &graph, &deps, &ndeps));
    CU(cudaGraphConditionalHandleCreate(&hnd, graph, 0u, cudaGraphCondAssignDefault));
    KL(1, launch_beam_set_condition((unsigned long long)hnd, bb.grew + (t - 1), st));
    while (t < L) {
      const int t_end = std::min(L, t + blk);
      cudaGraphConditionalHandle next = 0;
      CU(cudaStreamGetCaptureInfo_v2(st, &cs, &cid, &graph, &deps, &ndeps));
      if (t_end < L) CU(cudaGraphConditionalHandleCreate(&next, graph, 0u, cudaGraphCondAssignDefault));
      cudaGraphNodeParams np = {};
      np.type = cudaGraphNodeTypeConditional;
      np.conditional.handle = hnd;
      np.conditional.type = cudaGraphCondTypeIf;
      np.conditional.size = 1;
      cudaGraphNode_t node = nullptr;
      CU(cudaGraphAddNode(&node, graph, deps, ndeps, &np));
      CU(cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies));
      CU(cudaStreamBeginCaptureToGraph(sb, np.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
      int rr = 0;
      for (; t < t_end && !rr; ++t) rr = one_step(t, sb);
      if (!rr && t_end < L) {
        cudaError_t e_ = launch_beam_set_condition((unsigned long long)next, bb.grew + (t_end - 1), sb);
        h->launches += 1;
        if (e_ != cudaSuccess) rr = h->fail(XN_ERR_CUDA, "launch_beam_set_condition failed: %s", cudaGetErrorString(e_));
      }
      cudaError_t ee = cudaStreamEndCapture(sb, nullptr);
      if (rr) return rr;
      if (ee != cudaSuccess) return h->fail(XN_ERR_CUDA, "capture of a decode-step body failed: %s", cudaGetErrorString(ee));
      t = t_end;
      hnd = next;
    }
  }
  KL(1, launch_beam_finalize(bb, src, B, beam, L, L, how_many, P.r_tok, P.r_len, P.r_lp, st));
  return 0;
}

int beam_copy_out(xn_handle* h, const BeamPlan& P, int B, int L, int how_many, int32_t* out_tokens, int32_t* out_len,
                  float* out_lp, cudaStream_t st) {
  CU(cudaMemcpyAsync(out_tokens, P.r_tok, (size_t)B * how_many * L * 4, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(out_len, P.r_len, (size_t)B * how_many * 4, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(out_lp, P.r_lp, (size_t)B * how_many * L * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int beam_from_enc(xn_handle* h, const float* enc_out, int B, const int32_t* enc_pads_host, int beam, int L, int how_many,
                  int sos, int eos, int32_t* out_tokens, int32_t* out_len, float* out_lp, cudaStream_t st, bool reset_ws) {
  BeamPlan P;
  if (int r = beam_plan(h, P, B, enc_pads_host, beam, L, how_many, st, reset_ws)) return r;
  const xn_handle::GraphKey key{0, enc_out, B, beam, L, how_many, sos, eos, h->ws.base, h->ws.cap};
  if (int r = run_graphed(h, key, P.nv == nullptr, st, [&](cudaStream_t s2) { return beam_run(h, P, enc_out, B, beam, L, how_many, sos, eos, s2); }))
    return r;
  return beam_copy_out(h, P, B, L, how_many, out_tokens, out_len, out_lp, st);
}

size_t beam_ws_bytes(const xn_config& c, int B, int beam, int L) {
  const int R = B * beam;
  return dec_ws_bytes(c, R, L, B, true) + (size_t)R * beam * 8 + (size_t)R * L * 24 + R * 8 + B * 4 + (size_t)R * L * 8 + R * 4 + R * 16 + 80 * 256;
}

const float* rawp(xn_handle* h, const std::string& k, std::vector<int64_t> shape, int* rc) {
  auto it = h->raw.find(k);
  if (it == h->raw.end()) { *rc = h->fail(XN_ERR_STATE, "missing state_dict entry '%s'", k.c_str()); return nullptr; }
  if (it->second.shape != shape) {
    std::string got, want;
    for (auto s : it->second.shape) got += std::to_string(s) + ",";
    for (auto s : shape) want += std::to_string(s) + ",";
    *rc = h->fail(XN_ERR_STATE, "shape mismatch for '%s': got (%s) expected (%s)", k.c_str(), got.c_str(), want.c_str());
    return nullptr;
  }
  return it->second.p;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int xn_create(const xn_config* cfg, int device, xn_handle** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return XN_ERR_ARG; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev) {
    g_create_error = std::string("no usable CUDA device ") + std::to_string(device) + ": " +
                     (e != cudaSuccess ? cudaGetErrorString(e) : "ordinal out of range") +
                     " (xnv2_b200 has no CPU fallback)";
    return XN_ERR_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_error = "xnv2_b200 is built for sm_100a (B200) only; device reports sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
    return XN_ERR_UNSUPPORTED;
  }
  if (cfg->window_size != 12) { g_create_error = "only swin_window_size == 12 is supported"; return XN_ERR_UNSUPPORTED; }
  if (cfg->d_model % 128 || cfg->d_model % cfg->num_heads) { g_create_error = "d_model must be a multiple of 128 and of num_heads"; return XN_ERR_UNSUPPORTED; }
  if (cfg->has_swin) {
    if (cfg->n_stages < 1 || cfg->n_stages > 4) { g_create_error = "1..4 swin stages"; return XN_ERR_ARG; }
    for (int s = 0; s < cfg->n_stages; ++s) {
      const int C = cfg->embed_dim << s, H = (cfg->img_size / cfg->patch_size) >> s;
      if (C != cfg->swin_heads[s] * 32) { g_create_error = "swin head_dim must be 32"; return XN_ERR_UNSUPPORTED; }
      if (H % 12) { g_create_error = "every swin stage resolution must be a multiple of the window (12)"; return XN_ERR_UNSUPPORTED; }
    }
    const int Hl = (cfg->img_size / cfg->patch_size) >> (cfg->n_stages - 1);
    if (Hl * Hl != cfg->enc_len || (cfg->embed_dim << (cfg->n_stages - 1)) != cfg->feat_dim) {
      g_create_error = "swin output (L,C) must equal (enc_len, feat_dim)";
      return XN_ERR_ARG;
    }
  }
  for (int i = 0; i < cfg->n_exp_groups; ++i)
    if (cfg->exp_groups[i] % 8) { g_create_error = "num_exp_enc_list entries must be multiples of 8"; return XN_ERR_UNSUPPORTED; }
  xn_handle* h = new xn_handle();
  h->cfg = *cfg;
  h->device = device;
  cudaSetDevice(device);
  if (cudaMalloc(&h->flag_dev, sizeof(int)) != cudaSuccess || cudaMemset(h->flag_dev, 0, sizeof(int)) != cudaSuccess ||
      cudaMalloc(&h->mega_bar, kMegaBarBytes * (xn_handle::kMaxDecodeGroups + 1)) != cudaSuccess ||
      cudaMemset(h->mega_bar, 0, kMegaBarBytes * (xn_handle::kMaxDecodeGroups + 1)) != cudaSuccess) {
    g_create_error = "cudaMalloc failed";
    delete h;
    return XN_ERR_CUDA;
  }
  *out = h;
  return XN_OK;
}

int xn_destroy(xn_handle* h) {
  if (!h) return XN_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& kv : h->raw) cudaFree(kv.second.p);
  for (void* p : h->owned) cudaFree(p);
  if (h->group_start_dev) cudaFree(h->group_start_dev);
  if (h->ws.base) cudaFree(h->ws.base);
  if (h->io_in) cudaFree(h->io_in);
  for (auto& sl : h->slots) {
    if (sl.in) cudaFree(sl.in);
    if (sl.out) cudaFree(sl.out);
    if (sl.ev_in) cudaEventDestroy(sl.ev_in);
    if (sl.ev_done) cudaEventDestroy(sl.ev_done);
  }
  if (h->hstream) cudaStreamDestroy(h->hstream);
  if (h->flag_dev) cudaFree(h->flag_dev);
  if (h->mega_bar) cudaFree(h->mega_bar);
  if (h->mega_dbg) cudaFree(h->mega_dbg);
  if (h->pp_buf) cudaFree(h->pp_buf);
  if (h->pp_host) cudaFreeHost(h->pp_host);
  if (h->jpg_state) g_nvjpeg.state_destroy(h->jpg_state);
  if (h->jpg) g_nvjpeg.destroy(h->jpg);
  if (h->jpg_buf) cudaFree(h->jpg_buf);
  if (h->pp_ev) cudaEventDestroy(h->pp_ev);
  for (auto& t : h->rtables) { cudaFree(t.bounds); cudaFree(t.kk); }
  if (h->io_out) cudaFree(h->io_out);
  h->drop_graphs();
  if (h->gstream) { cudaStreamDestroy(h->gstream); cudaEventDestroy(h->g_in); cudaEventDestroy(h->g_out); }
  for (int g = 0; g < xn_handle::kMaxDecodeGroups; ++g) {
    if (h->dstream[g]) cudaStreamDestroy(h->dstream[g]);
    if (h->d_join[g]) cudaEventDestroy(h->d_join[g]);
  }
  if (h->d_fork) cudaEventDestroy(h->d_fork);
  for (auto& bs : h->bstream) if (bs) cudaStreamDestroy(bs);
  if (h->cstream) cudaStreamDestroy(h->cstream);
  if (h->c_fork) cudaEventDestroy(h->c_fork);
  for (int i = 0; i < xn_handle::kMaxCopyChunks; ++i) if (h->c_ev[i]) cudaEventDestroy(h->c_ev[i]);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  for (auto& sp : h->spans) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
  for (cudaEvent_t e : h->span_pool) cudaEventDestroy(e);
  delete h;
  return XN_OK;
}

const char* xn_last_error(const xn_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int xn_load_tensor(xn_handle* h, const char* key, const void* data, int dtype, const int64_t* shape, int ndim) {
  if (!h || !key || !data) return XN_ERR_ARG;
  cudaSetDevice(h->device);
  std::string k(key);
  auto ends = [&](const char* s) { size_t n = strlen(s); return k.size() >= n && k.compare(k.size() - n, n, s) == 0; };
  if (ends("relative_position_index") || ends("attn_mask")) return XN_OK;   // geometry-only buffers
  if (dtype != XN_DTYPE_F32) return h->fail(XN_ERR_ARG, "tensor '%s': only f32 parameters are accepted", key);
  DevTensor t;
  t.n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); t.n *= (size_t)shape[i]; }
  auto it = h->raw.find(k);
  if (it != h->raw.end()) { cudaFree(it->second.p); h->raw.erase(it); }
  CU(cudaMalloc(&t.p, std::max<size_t>(t.n, 1) * sizeof(float)));
  CU(cudaMemcpy(t.p, data, t.n * sizeof(float), cudaMemcpyDefault));
  h->raw[k] = t;
  h->precision = -1;
  return XN_OK;
}

int xn_finalize_weights(xn_handle* h, int precision) {
  if (!h) return XN_ERR_ARG;
  if (precision != XN_PREC_FP32 && precision != XN_PREC_BF16 && precision != XN_PREC_FP16) return h->fail(XN_ERR_ARG, "bad precision");
  cudaSetDevice(h->device);
  const xn_config& c = h->cfg;
  h->drop_graphs();
  for (void* p : h->owned) cudaFree(p);
  h->owned.clear();
  h->stages.clear(); h->enc.clear(); h->dec.clear();
  int rc = 0;
  const int64_t d = c.d_model, ff = c.ff, V = c.vocab;
  auto P = [&](const std::string& k, std::vector<int64_t> s) { return rc ? nullptr : rawp(h, k, s, &rc); };
  auto lin = [&](const std::string& name, int64_t N, int64_t K, bool bias = true) {
    LinW l;
    l.w = P(name + ".weight", {N, K});
    l.b = bias ? P(name + ".bias", {N}) : nullptr;
    l.N = (int)N; l.K = (int)K;
    return l;
  };
  auto to_bf16 = [&](LinW& l) -> int {      // 16-bit operand copy in the mode's format
    if (rc || !l.w) return rc;
    void* p = nullptr;
    CU(cudaMalloc(&p, (size_t)l.N * l.K * 2));
    h->owned.push_back(p);
    if (precision == XN_PREC_FP16) KL(1, launch_cast<f16>(l.w, reinterpret_cast<f16*>(p), (long)l.N * l.K, 0));
    else KL(1, launch_cast<bf16>(l.w, reinterpret_cast<bf16*>(p), (long)l.N * l.K, 0));
    l.wb = p;
    return 0;
  };
  auto to_packed = [&](LinW& l) -> int {    // decoder linears: also the persistent kernel's slab-packed layout
    if (rc || !l.w || (l.K % 512)) return rc;
    void* p = nullptr;
    CU(cudaMalloc(&p, mega_packed_bytes(l.N, l.K)));
    h->owned.push_back(p);
    KL(1, launch_mega_pack_weight(l.w, p, l.N, l.K, precision == XN_PREC_FP16, 0));
    l.wp = p;
    return 0;
  };
  // concatenate row blocks of several (N_i x K) weights (+ biases) into one (sum N_i x K) weight
  auto concat = [&](const std::vector<LinW>& parts, LinW& outl) -> int {
    if (rc) return rc;
    int N = 0, K = parts[0].K;
    for (auto& p : parts) N += p.N;
    float *w = nullptr, *b = nullptr;
    CU(cudaMalloc(&w, (size_t)N * K * sizeof(float)));
    h->owned.push_back(w);
    CU(cudaMalloc(&b, (size_t)N * sizeof(float)));
    h->owned.push_back(b);
    int n0 = 0;
    for (auto& p : parts) {
      CU(cudaMemcpy(w + (size_t)n0 * K, p.w, (size_t)p.N * K * sizeof(float), cudaMemcpyDeviceToDevice));
      CU(cudaMemcpy(b + n0, p.b, (size_t)p.N * sizeof(float), cudaMemcpyDeviceToDevice));
      n0 += p.N;
    }
    outl.w = w; outl.b = b; outl.N = N; outl.K = K;
    return 0;
  };

  if (c.has_swin) {
    const std::string p = "swin_transf.";
    h->pe_w = P(p + "patch_embed.proj.weight", {c.embed_dim, c.in_chans, c.patch_size, c.patch_size});
    h->pe_b = P(p + "patch_embed.proj.bias", {c.embed_dim});
    h->pe_wq = nullptr;
    const size_t pe4_smem = ((size_t)c.in_chans * 4 * (c.img_size / 4) + (size_t)c.in_chans * 4 * c.embed_dim) * 16;
    if (!rc && c.patch_size == 4 && c.img_size % 16 == 0 && c.embed_dim % 32 == 0 && c.embed_dim <= 256 && pe4_smem <= 200 * 1024) {
      float* wq = nullptr;
      CU(cudaMalloc(&wq, (size_t)c.in_chans * 16 * c.embed_dim * sizeof(float)));
      h->owned.push_back(wq);
      KL(1, launch_patch_filter_pack4(h->pe_w, wq, c.in_chans, c.embed_dim, 0));
      h->pe_wq = wq;
    }
    h->pe_g = P(p + "patch_embed.norm.weight", {c.embed_dim});
    h->pe_beta = P(p + "patch_embed.norm.bias", {c.embed_dim});
    for (int s = 0; s < c.n_stages; ++s) {
      SwinStageW S;
      S.C = c.embed_dim << s; S.H = (c.img_size / c.patch_size) >> s; S.heads = c.swin_heads[s];
      const int64_t C = S.C, hid = (int64_t)(C * c.mlp_ratio);
      for (int b = 0; b < c.depths[s]; ++b) {
        const std::string q = p + "layers." + std::to_string(s) + ".blocks." + std::to_string(b) + ".";
        SwinBlockW W;
        W.n1g = P(q + "norm1.weight", {C}); W.n1b = P(q + "norm1.bias", {C});
        W.n2g = P(q + "norm2.weight", {C}); W.n2b = P(q + "norm2.bias", {C});
        W.rpb = P(q + "attn.relative_position_bias_table", {23 * 23, S.heads});
        W.qkv = lin(q + "attn.qkv", 3 * C, C); W.proj = lin(q + "attn.proj", C, C);
        W.fc1 = lin(q + "mlp.fc1", hid, C); W.fc2 = lin(q + "mlp.fc2", C, hid);
        W.rpb_t = nullptr;
        if (precision != XN_PREC_FP32) {
          if (to_bf16(W.qkv) || to_bf16(W.proj) || to_bf16(W.fc1) || to_bf16(W.fc2)) return XN_ERR_CUDA;
          if (!rc) {
            auto fold = [&](const LinW& src, const float* g_, const float* b_, LinW& dst) -> int {
              void* w16 = nullptr; float* bo = nullptr;
              CU(cudaMalloc(&w16, (size_t)src.N * src.K * 2)); h->owned.push_back(w16);
              CU(cudaMalloc(&bo, (size_t)src.N * sizeof(float))); h->owned.push_back(bo);
              KL(1, launch_fold_ln_weight(src.w, g_, b_, src.b, w16, bo, src.N, src.K, precision == XN_PREC_FP16, 0));
              dst = src; dst.w = nullptr; dst.wb = w16; dst.b = bo;
              return 0;
            };
            if (fold(W.qkv, W.n1g, W.n1b, W.qkv_ln) || fold(W.fc1, W.n2g, W.n2b, W.fc1_ln)) return XN_ERR_CUDA;
          }
          if (!rc) {
            float* bt = nullptr;
            CU(cudaMalloc(&bt, bias_derived_floats(S.heads) * sizeof(float)));
            h->owned.push_back(bt);
            KL(1, launch_transpose_bias(W.rpb, bt, S.heads, 0));
            W.rpb_t = bt;
          }
        }
        S.blocks.push_back(W);
      }
      if (s < c.n_stages - 1) {
        const std::string q = p + "layers." + std::to_string(s) + ".downsample.";
        S.has_merge = true;
        S.mg = P(q + "norm.weight", {4 * C}); S.mb = P(q + "norm.bias", {4 * C});
        S.red = lin(q + "reduction", 2 * C, 4 * C, false);
        if (precision != XN_PREC_FP32 && to_bf16(S.red)) return XN_ERR_CUDA;
      }
      h->stages.push_back(S);
    }
    h->swin_ng = P(p + "norm.weight", {c.feat_dim});
    h->swin_nb = P(p + "norm.bias", {c.feat_dim});
  }
  int64_t E = 0;
  std::vector<int> gstart(1, 0);
  int gcd = 32;
  for (int i = 0; i < c.n_exp_groups; ++i) { E += c.exp_groups[i]; gstart.push_back((int)E); }
  auto gcdf = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
  for (int v : gstart) if (v) gcd = gcdf(gcd, v);
  h->n_exp_total = (int)E; h->exp_chunk = gcd;
  h->se_t_ok = static_exp_t_supported(gstart.data(), c.n_exp_groups, (int)E, c.enc_len);
  if (h->group_start_dev) cudaFree(h->group_start_dev);
  CU(cudaMalloc(&h->group_start_dev, gstart.size() * sizeof(int)));
  CU(cudaMemcpy(h->group_start_dev, gstart.data(), gstart.size() * sizeof(int), cudaMemcpyHostToDevice));

  for (int i = 0; i < c.n_enc; ++i) {
    const std::string q = "encoders." + std::to_string(i) + ".";
    EncLayerW W;
    W.n1g = P(q + "norm_1.weight", {d}); W.n1b = P(q + "norm_1.bias", {d});
    W.n2g = P(q + "norm_2.weight", {d}); W.n2b = P(q + "norm_2.bias", {d});
    W.qexp = P(q + "stc_exp.query_exp_vectors.weight", {E, d});
    W.bexp = P(q + "stc_exp.bias_exp_vectors.weight", {E, d});
    std::vector<LinW> parts = {lin(q + "stc_exp.key_embed", d, d), lin(q + "stc_exp.class_a_embed", d, d),
                               lin(q + "stc_exp.class_b_embed", d, d), lin(q + "stc_exp.selector_embed", d, d)};
    W.ff1 = lin(q + "ff.linear_1", ff, d); W.ff2 = lin(q + "ff.linear_2", d, ff);
    if (rc) return rc;
    if (concat(parts, W.kabs)) return XN_ERR_CUDA;
    W.qexp16 = nullptr; W.bexpT = nullptr;
    if (precision != XN_PREC_FP32) {
      if (to_bf16(W.kabs) || to_bf16(W.ff1) || to_bf16(W.ff2)) return XN_ERR_CUDA;
      LinW qe; qe.w = W.qexp; qe.N = (int)E; qe.K = (int)d;       // 16-bit copy of the expansion queries (left operand of z)
      if (to_bf16(qe)) return XN_ERR_CUDA;
      W.qexp16 = qe.wb;
      float* bt = nullptr;                                          // bias_exp transposed [d][E]: residual of class^T
      CU(cudaMalloc(&bt, (size_t)E * d * sizeof(float))); h->owned.push_back(bt);
      CU(launch_transpose_f32(W.bexp, bt, (int)E, (int)d, nullptr));
      W.bexpT = bt;
    }
    h->enc.push_back(W);
  }
  std::vector<LinW> kvparts;
  for (int i = 0; i < c.n_dec; ++i) {
    const std::string q = "decoders." + std::to_string(i) + ".";
    DecLayerW W;
    W.n1g = P(q + "norm_1.weight", {d}); W.n1b = P(q + "norm_1.bias", {d});
    W.n2g = P(q + "norm_2.weight", {d}); W.n2b = P(q + "norm_2.bias", {d});
    W.n3g = P(q + "norm_3.weight", {d}); W.n3b = P(q + "norm_3.bias", {d});
    W.qexp = P(q + "dyn_exp.query_exp_vectors.weight", {c.num_exp_dec, d});
    W.bexp = P(q + "dyn_exp.bias_exp_vectors.weight", {c.num_exp_dec, d});
    std::vector<LinW> parts = {lin(q + "dyn_exp.cond_embed", d, d), lin(q + "dyn_exp.key_linear", d, d),
                               lin(q + "dyn_exp.class_a_embed", d, d), lin(q + "dyn_exp.class_b_embed", d, d),
                               lin(q + "dyn_exp.selector_embed", d, d)};
    W.wq = lin(q + "mha.Wq", d, d); W.wo = lin(q + "mha.out_linear", d, d);
    kvparts.push_back(lin(q + "mha.Wk", d, d));
    kvparts.push_back(lin(q + "mha.Wv", d, d));
    W.ff1 = lin(q + "ff.linear_1", ff, d); W.ff2 = lin(q + "ff.linear_2", d, ff);
    if (rc) return rc;
    if (concat(parts, W.dyn5)) return XN_ERR_CUDA;
    if (precision != XN_PREC_FP32 && (to_bf16(W.dyn5) || to_bf16(W.wq) || to_bf16(W.wo) || to_bf16(W.ff1) || to_bf16(W.ff2))) return XN_ERR_CUDA;
    if (precision != XN_PREC_FP32 && (to_packed(W.dyn5) || to_packed(W.wq) || to_packed(W.wo) || to_packed(W.ff1) || to_packed(W.ff2))) return XN_ERR_CUDA;
    h->dec.push_back(W);
  }
  h->input_linear = lin("input_linear", d, c.feat_dim);
  h->vocab = lin("vocab_linear", V, d);
  h->enc_reduce = lin("enc_reduce_group", d, d * c.n_enc);
  if (precision != XN_PREC_FP32 && (to_bf16(h->input_linear) || to_bf16(h->enc_reduce))) return XN_ERR_CUDA;
  h->dec_reduce = lin("dec_reduce_group", d, d * c.n_dec);
  h->enc_ng = P("enc_reduce_norm.weight", {d}); h->enc_nb = P("enc_reduce_norm.bias", {d});
  h->dec_ng = P("dec_reduce_norm.weight", {d}); h->dec_nb = P("dec_reduce_norm.bias", {d});
  h->emb = P("out_embedder.embed.weight", {V, d});
  h->pos = P("pos_encoder.weight", {c.max_seq_len, d});
  if (rc) return rc;
  if (concat(kvparts, h->kv_all)) return XN_ERR_CUDA;
  if (precision != XN_PREC_FP32 && (to_bf16(h->kv_all) || to_bf16(h->vocab) || to_bf16(h->dec_reduce))) return XN_ERR_CUDA;
  if (precision != XN_PREC_FP32 && (to_packed(h->vocab) || to_packed(h->dec_reduce))) return XN_ERR_CUDA;
  CU(cudaDeviceSynchronize());
  h->precision = precision;
  return XN_OK;
}

#define NEED_READY()                                                                   \
  if (!h) return XN_ERR_ARG;                                                           \
  if (h->precision < 0) return h->fail(XN_ERR_STATE, "weights not finalised");         \
  cudaSetDevice(h->device);                                                            \
  cudaStream_t st = (cudaStream_t)stream;                                              \
  h->cur_st = st;

int xn_forward_swin(xn_handle* h, const float* images, int B, float* out, void* stream) {
  NEED_READY();
  if (!h->cfg.has_swin) return h->fail(XN_ERR_STATE, "model has no Swin backbone");
  const int Bc = (int)std::min<int64_t>(B, h->swin_chunk);
  if (int r = ensure_ws(h, swin_ws_bytes(h->cfg, Bc, h->precision), st)) return r;
  return swin_forward(h, images, B, out, st);
}

int xn_forward_enc(xn_handle* h, const float* input, int B, const int32_t* enc_pads_host, float* out, void* stream) {
  NEED_READY();
  const xn_config& c = h->cfg;
  const int Bs = (int)std::min<int64_t>(B, h->swin_chunk), Be = (int)std::min<int64_t>(B, h->enc_chunk);
  const size_t feat_bytes = c.has_swin ? (size_t)B * c.enc_len * c.feat_dim * 4 + 4096 : 0;
  size_t need = std::max(c.has_swin ? swin_ws_bytes(c, Bs, h->precision) : 0, enc_ws_bytes(c, Be));
  if (int r = ensure_ws(h, need + feat_bytes, st)) return r;
  const float* feats = input;
  if (c.has_swin) {
    if (enc_pads_host)
      for (int i = 0; i < B; ++i)
        if (enc_pads_host[i] != 0) return h->fail(XN_ERR_ARG, "End to End case have no padding");
    // feature buffer lives at the top of the arena, below it the per-chunk scratch
    float* fb = reinterpret_cast<float*>(h->ws.base + ((h->ws.cap - feat_bytes) & ~size_t(255)));
    const size_t keep = h->ws.cap;
    h->ws.cap = (h->ws.cap - feat_bytes) & ~size_t(255);
    int r = swin_forward(h, input, B, fb, st);
    if (!r) r = enc_body(h, fb, B, nullptr, out, st);
    h->ws.cap = keep;
    if (!r && h->precision != XN_PREC_FP32) KL(1, launch_nonfinite_flag(out, (long)B * c.enc_len * c.d_model, h->flag_dev, st));
    return r;
  }
  if (int r = enc_body(h, feats, B, enc_pads_host, out, st)) return r;
  if (h->precision != XN_PREC_FP32) KL(1, launch_nonfinite_flag(out, (long)B * c.enc_len * c.d_model, h->flag_dev, st));
  return XN_OK;
}

int xn_forward_dec(xn_handle* h, const float* cross, int R, const int32_t* enc_pads_host, const int64_t* tokens, int t,
                   const int32_t* dec_pads_host, int apply_log_softmax, float* out, void* stream) {
  NEED_READY();
  const xn_config& c = h->cfg;
  if (t < 1 || t > c.max_seq_len || t > 128) return h->fail(XN_ERR_ARG, "sequence length %d outside [1, %d]", t, std::min(c.max_seq_len, 128));
  if (int r = ensure_ws(h, dec_ws_bytes(c, R, t, R, false) + (size_t)R * 8, st)) return r;
  DecBufs D;
  dec_alloc(h, D, R, t, R);
  WS_CHECK();
  int *nv = nullptr, *rl = nullptr;
  if (enc_pads_host && !c.has_swin) {
    std::vector<int> v(R);
    bool any = false;
    for (int i = 0; i < R; ++i) { v[i] = c.enc_len - enc_pads_host[i]; any |= enc_pads_host[i] != 0; }
    if (any) if (int r = upload_ints(h, v, &nv, st)) return r;
  }
  if (dec_pads_host) {
    std::vector<int> v(R);
    bool any = false;
    for (int i = 0; i < R; ++i) { v[i] = t - dec_pads_host[i]; any |= dec_pads_host[i] != 0; }
    if (any) if (int r = upload_ints(h, v, &rl, st)) return r;
  }
  if (int r = dec_project(h, D, cross, R, st)) return r;
  const long ldl = (long)t * c.vocab;
  for (int p = 0; p < t; ++p) {
    float* lg = out + (size_t)p * c.vocab;
    if (int r = dec_step(h, D, p, tokens, nullptr, t, 1, nv, rl, lg, ldl, st)) return r;
    if (apply_log_softmax) KL(1, launch_logsoftmax_topk(lg, ldl, R, c.vocab, 0, nullptr, nullptr, lg, ldl, 1, st));
  }
  return XN_OK;
}

int xn_beam_search_from_enc(xn_handle* h, const float* enc_out, int B, const int32_t* enc_pads_host, int beam, int max_len,
                            int how_many, int sos_idx, int eos_idx, int32_t* out_tokens, int32_t* out_len,
                            float* out_logprob, void* stream) {
  NEED_READY();
  if (int r = ensure_ws(h, beam_ws_bytes(h->cfg, B, beam, max_len), st)) return r;
  return beam_from_enc(h, enc_out, B, h->cfg.has_swin ? nullptr : enc_pads_host, beam, max_len, how_many, sos_idx, eos_idx,
                       out_tokens, out_len, out_logprob, st, true);
}

}  // extern "C"

// restores the arena capacity on every exit path (the feature / encoder-output buffers are carved off its top per call)
struct ArenaCapGuard {
  Arena& a; size_t keep;
  explicit ArenaCapGuard(Arena& ar) : a(ar), keep(ar.cap) {}
  ~ArenaCapGuard() { a.cap = keep; }
};

// images -> captions.  host_src != nullptr: the images come from that pinned host buffer, copied chunk-wise inside the
// call (xn_caption_host).  Otherwise `input` is the caller's device tensor: it is first copied into the handle's own
// staging buffer (one device-to-device copy, ~35 us per 113 MB) so that everything after it -- and therefore the cached
// CUDA graph of the call shape -- reads a stable address: any caller tensor replays the same graph.
static int beam_search_impl(xn_handle* h, const float* input, const float* host_src, int B, const int32_t* enc_pads_host, int beam,
                            int max_len, int how_many, int sos_idx, int eos_idx, int32_t* out_tokens, int32_t* out_len,
                            float* out_logprob, void* stream) {
  NEED_READY();
  const xn_config& c = h->cfg;
  if (B < 1) return h->fail(XN_ERR_ARG, "empty batch");
  if (how_many > beam) return h->fail(XN_ERR_ARG, "requested output per sequence must be lower than beam width");
  const int Bs = (int)std::min<int64_t>(B, h->swin_chunk), Be = (int)std::min<int64_t>(B, h->enc_chunk);
  const size_t enc_bytes = ((size_t)B * c.enc_len * c.d_model * 4 + 4095) & ~size_t(255);
  const size_t feat_bytes = c.has_swin ? (((size_t)B * c.enc_len * c.feat_dim * 4 + 4095) & ~size_t(255)) : 0;
  size_t scratch = std::max(std::max(c.has_swin ? swin_ws_bytes(c, Bs, h->precision) : 0, enc_ws_bytes(c, Be)),
                            beam_ws_bytes(c, B, beam, max_len) + xn_handle::kMaxDecodeGroups * (size_t)(96 * 256));   // + per-group rounding
  if (int r = ensure_ws(h, scratch + enc_bytes + feat_bytes, st)) return r;
  bool pads = false;
  if (enc_pads_host)
    for (int i = 0; i < B; ++i) pads |= enc_pads_host[i] != 0;
  if (c.has_swin && pads) return h->fail(XN_ERR_ARG, "End to End case have no padding");
  const size_t in_elems = c.has_swin ? (size_t)B * c.in_chans * c.img_size * c.img_size : (size_t)B * c.enc_len * c.feat_dim;
  const bool stage_input = !host_src && !pads && h->use_graph && !h->profile && input != h->staged_input;
  if (host_src || stage_input)
    if (int r = ensure_io_in(h, in_elems * 4)) return r;
  ArenaCapGuard cap_guard(h->ws);
  const size_t keep = h->ws.cap;
  char* top = h->ws.base + (keep & ~size_t(255));
  float* enc_out = reinterpret_cast<float*>(top - enc_bytes);
  float* fb = reinterpret_cast<float*>(top - enc_bytes - feat_bytes);
  h->ws.cap = (size_t)((top - enc_bytes - feat_bytes) - h->ws.base);
  if (pads) {
    // encoder padding (features-in model): per-image valid lengths are uploaded with a host sync -> plain stream order
    if (int r = enc_body(h, input, B, enc_pads_host, enc_out, st)) return r;
    return beam_from_enc(h, enc_out, B, enc_pads_host, beam, max_len, how_many, sos_idx, eos_idx, out_tokens, out_len, out_logprob, st, true);
  }
  const float* in_dev = input;
  if (host_src) in_dev = h->io_in;
  else if (stage_input) {
    if (input != h->io_in) CU(cudaMemcpyAsync(h->io_in, input, in_elems * 4, cudaMemcpyDeviceToDevice, st));
    in_dev = h->io_in;
  }
  // The beam buffers are carved from the bottom of the arena, where the Swin / encoder chunks also put their scratch:
  // both run strictly before the decoder, in stream order.  The images are decoded in G independent groups (beam search
  // never mixes images), each with its own buffers, on G streams forked from the encoder's stream.
  int G = (int)h->decode_groups;
  // measured at batch 64 (profiles/r2_quick_time_decode_groups.txt): 2 groups -0.6 ms, 4 groups +0.9 ms, 8 groups +3.8 ms --
  // beyond two chains the graph's kernel-node dispatch rate, not kernel latency, is the limit
  if (G <= 0) G = B >= 32 ? 2 : 1;
  if (h->profile) G = 1;                                    // per-launch event timing wants one stream
  G = std::max(1, std::min(G, std::min(B, (int)xn_handle::kMaxDecodeGroups)));
  std::vector<BeamPlan> P(G);
  std::vector<int> g0(G + 1, 0);
  for (int g = 0; g < G; ++g) g0[g + 1] = g0[g] + B / G + (g < B % G ? 1 : 0);
  for (int g = 0; g < G; ++g)
    if (int r = beam_plan(h, P[g], g0[g + 1] - g0[g], nullptr, beam, max_len, how_many, st, g == 0)) return r;
  if (G > 1 && !h->d_fork) {
    CU(cudaEventCreateWithFlags(&h->d_fork, cudaEventDisableTiming));
    for (int g = 0; g < xn_handle::kMaxDecodeGroups; ++g) {
      CU(cudaStreamCreateWithFlags(&h->dstream[g], cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&h->d_join[g], cudaEventDisableTiming));
    }
  }
  const int host_chunk = 32;                                // copy granularity of the host path (chunk c+1 lands during chunk c)
  const xn_handle::GraphKey key{1 + 16 * G + (host_src ? 4096 : 0), in_dev, B, beam, max_len, how_many,
                                sos_idx, eos_idx, h->ws.base, keep};
  int r = run_graphed(h, key, true, st, [&](cudaStream_t s2) -> int {
    if (c.has_swin) {
      if (int rr = swin_forward(h, in_dev, B, fb, s2, host_src, host_chunk)) return rr;
      if (int rr = enc_body(h, fb, B, nullptr, enc_out, s2)) return rr;
    } else {
      if (int rr = enc_body(h, in_dev, B, nullptr, enc_out, s2)) return rr;
    }
    if (h->precision != XN_PREC_FP32) KL(1, launch_nonfinite_flag(enc_out, (long)B * c.enc_len * c.d_model, h->flag_dev, s2));
    if (G == 1) return beam_run(h, P[0], enc_out, B, beam, max_len, how_many, sos_idx, eos_idx, s2);
    // fork / join: every branch is joined back into s2 whatever happens in between (a capture must not end forked);
    // the first error is remembered and reported after the join
    int rr = 0;
    auto note = [&](cudaError_t e, const char* what) {
      if (e != cudaSuccess && !rr) rr = h->fail(XN_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
    };
    note(cudaEventRecord(h->d_fork, s2), "cudaEventRecord(fork)");
    for (int g = 0; g < G; ++g) {                           // issue order interleaves nothing: each group is one chain
      cudaStream_t sg = h->dstream[g];
      note(cudaStreamWaitEvent(sg, h->d_fork, 0), "cudaStreamWaitEvent(fork)");
      if (!rr) rr = beam_run(h, P[g], enc_out + (size_t)g0[g] * c.enc_len * c.d_model, g0[g + 1] - g0[g], beam, max_len, how_many,
                             sos_idx, eos_idx, sg);
      note(cudaEventRecord(h->d_join[g], sg), "cudaEventRecord(join)");
      note(cudaStreamWaitEvent(s2, h->d_join[g], 0), "cudaStreamWaitEvent(join)");
    }
    return rr;
  }, host_src, in_elems * 4);
  for (int g = 0; g < G && !r; ++g) {
    const size_t o = (size_t)g0[g] * how_many;
    r = beam_copy_out(h, P[g], g0[g + 1] - g0[g], max_len, how_many, out_tokens + o * max_len, out_len + o, out_logprob + o * max_len, st);
  }
  return r;
}

extern "C" {

int xn_beam_search(xn_handle* h, const float* input, int B, const int32_t* enc_pads_host, int beam, int max_len, int how_many,
                   int sos_idx, int eos_idx, int32_t* out_tokens, int32_t* out_len, float* out_logprob, void* stream) {
  return beam_search_impl(h, input, nullptr, B, enc_pads_host, beam, max_len, how_many, sos_idx, eos_idx, out_tokens, out_len,
                          out_logprob, stream);
}

int xn_ensemble_beam_search(xn_handle* const* hs, int n_models, const float* input, int B, const int32_t* enc_pads_host, int beam,
                            int max_len, int how_many, int sos_idx, int eos_idx, int32_t* out_tokens, int32_t* out_len,
                            float* out_logprob, void* stream) {
  if (!hs || n_models < 1 || n_models > kMaxEnsemble || !hs[0]) return XN_ERR_ARG;
  xn_handle* h = hs[0];                       // owner of the shared beam state; errors are reported on it
  cudaStream_t st = (cudaStream_t)stream;
  const xn_config& c = h->cfg;
  for (int m = 0; m < n_models; ++m) {
    if (!hs[m] || hs[m]->precision < 0) return h->fail(XN_ERR_STATE, "ensemble model %d: weights not finalised", m);
    const xn_config& cm = hs[m]->cfg;
    if (hs[m]->device != h->device || cm.vocab != c.vocab || cm.has_swin != c.has_swin || cm.enc_len != c.enc_len ||
        cm.max_seq_len < max_len || (c.has_swin ? (cm.img_size != c.img_size || cm.in_chans != c.in_chans) : cm.feat_dim != c.feat_dim))
      return h->fail(XN_ERR_ARG, "ensemble model %d is not compatible with model 0 (device, vocabulary or input geometry)", m);
  }
  cudaSetDevice(h->device);
  if (how_many > beam) return h->fail(XN_ERR_ARG, "requested output per sequence must be lower than beam width");
  bool pads = false;
  if (enc_pads_host)
    for (int i = 0; i < B; ++i) pads |= enc_pads_host[i] != 0;
  if (c.has_swin && pads) return h->fail(XN_ERR_ARG, "End to End case have no padding");
  const int R = B * beam;
  std::vector<size_t> keep(n_models);
  std::vector<float*> enc_out(n_models);
  std::vector<DecBufs> D(n_models);
  std::vector<float*> logits(n_models);
  std::vector<int*> nv(n_models, nullptr);
  BeamPlan P;
  int rc = 0;
  // ---- per model: arena, encoder (image / feature input is shared), decoder buffers
  for (int m = 0; m < n_models && !rc; ++m) {
    xn_handle* hm = hs[m];
    hm->cur_st = st;
    const xn_config& cm = hm->cfg;
    const int Bs = (int)std::min<int64_t>(B, hm->swin_chunk), Be = (int)std::min<int64_t>(B, hm->enc_chunk);
    const size_t enc_bytes = ((size_t)B * cm.enc_len * cm.d_model * 4 + 4095) & ~size_t(255);
    const size_t feat_bytes = cm.has_swin ? (((size_t)B * cm.enc_len * cm.feat_dim * 4 + 4095) & ~size_t(255)) : 0;
    const size_t scratch = std::max(std::max(cm.has_swin ? swin_ws_bytes(cm, Bs, hm->precision) : 0, enc_ws_bytes(cm, Be)),
                                    beam_ws_bytes(cm, B, beam, max_len) + (size_t)B * beam * cm.vocab * 4 + 4096);   // + combined log-probs
    if ((rc = ensure_ws(hm, scratch + enc_bytes + feat_bytes, st))) { if (hm != h) h->err = hm->err; break; }
    keep[m] = hm->ws.cap;
    char* top = hm->ws.base + (keep[m] & ~size_t(255));
    enc_out[m] = reinterpret_cast<float*>(top - enc_bytes);
    float* fb = reinterpret_cast<float*>(top - enc_bytes - feat_bytes);
    hm->ws.cap = (size_t)((top - enc_bytes - feat_bytes) - hm->ws.base);
    if (cm.has_swin) {
      rc = swin_forward(hm, input, B, fb, st);
      if (!rc) rc = enc_body(hm, fb, B, nullptr, enc_out[m], st);
    } else {
      rc = enc_body(hm, input, B, enc_pads_host, enc_out[m], st);
    }
    if (rc) { if (hm != h) h->err = hm->err; break; }
    if (m == 0) {
      rc = beam_plan(hm, P, B, cm.has_swin ? nullptr : enc_pads_host, beam, max_len, how_many, st, true);
      D[0] = P.D; logits[0] = P.logits; nv[0] = P.nv;
    } else {
      hm->ws.reset();
      dec_alloc(hm, D[m], R, max_len, B);
      logits[m] = hm->ws.get<float>((size_t)R * cm.vocab);
      if (!cm.has_swin && pads) {
        std::vector<int> v(B);
        for (int i = 0; i < B; ++i) v[i] = cm.enc_len - enc_pads_host[i];
        rc = upload_ints(hm, v, &nv[m], st);
      }
      if (!rc && hm->ws.overflow) rc = hm->fail(XN_ERR_STATE, "workspace overflow in ensemble model %d", m);
    }
    if (rc && hm != h) h->err = hm->err;
  }
  // ---- the search of beam_run with the step distribution log(mean_m softmax(logits_m))  (ensemble_captioning_model.py:55-84)
  float* lp_comb = nullptr;
  if (!rc) {
    lp_comb = h->ws.get<float>((size_t)R * c.vocab);
    if (h->ws.overflow) rc = h->fail(XN_ERR_STATE, "workspace overflow (ensemble log-probabilities)");
  }
  auto step_all = [&](int p, const int* tok32, int src_anc) -> int {
    EnsembleLogits el{};
    el.n = n_models;
    for (int m = 0; m < n_models; ++m) {
      xn_handle* hm = hs[m];
      D[m].s.anc = P.bb.anc[src_anc];
      if (int r = dec_step(hm, D[m], p, nullptr, tok32, max_len, beam, nv[m], nullptr, logits[m], hm->cfg.vocab, st)) { if (hm != h) h->err = hm->err; return r; }
      el.p[m] = logits[m];
    }
    KL(1, launch_ensemble_logprob(el, c.vocab, R, c.vocab, lp_comb, c.vocab, st));
    KL(1, launch_logsoftmax_topk(lp_comb, c.vocab, R, c.vocab, beam, P.topv, P.topi, nullptr, 0, 2, st));
    return 0;
  };
  if (!rc) {
    for (int m = 0; m < n_models && !rc; ++m) {
      rc = dec_project(hs[m], D[m], enc_out[m], B, st);
      if (rc && hs[m] != h) h->err = hs[m]->err;
    }
  }
  if (!rc) {
    auto run = [&]() -> int {
      KL(1, launch_beam_init(P.bb, B, beam, max_len, sos_idx, st));
      int src = 0;
      if (int r = step_all(0, P.bb.tokens[0], 0)) return r;
      KL(1, launch_beam_first(P.bb, P.topv, P.topi, B, beam, max_len, eos_idx, st));
      int t_final = 2;
      for (int t = 2; t < max_len; ++t) {
        if (int r = step_all(t - 1, P.bb.tokens[src], src)) return r;
        KL(1, launch_beam_step(P.bb, src, P.topv, P.topi, B, beam, max_len, t, eos_idx, st));
        src ^= 1;
        t_final = t + 1;
      }
      KL(1, launch_beam_finalize(P.bb, src, B, beam, max_len, t_final, how_many, P.r_tok, P.r_len, P.r_lp, st));
      return 0;
    };
    rc = run();
  }
  if (!rc) rc = beam_copy_out(h, P, B, max_len, how_many, out_tokens, out_len, out_logprob, st);
  for (int m = 0; m < n_models; ++m)
    if (keep[m]) hs[m]->ws.cap = keep[m];
  return rc;
}

int xn_caption_host(xn_handle* h, const float* input_host, int B, int beam, int max_len, int how_many, int sos_idx,
                    int eos_idx, int32_t* out_tokens_host, int32_t* out_len_host, float* out_logprob_host, void* stream) {
  NEED_READY();
  const xn_config& c = h->cfg;
  const size_t in_elems = c.has_swin ? (size_t)B * c.in_chans * c.img_size * c.img_size : (size_t)B * c.enc_len * c.feat_dim;
  const size_t n_out = (size_t)B * how_many * max_len;
  if (int r = ensure_io_in(h, in_elems * 4)) return r;
  const size_t out_bytes = n_out * 8 + (size_t)B * how_many * 4 + 1024;
  if (h->io_out_cap < out_bytes) {
    if (h->io_out) { CU(cudaDeviceSynchronize()); cudaFree(h->io_out); h->io_out = nullptr; h->io_out_cap = 0; }
    CU(cudaMalloc(&h->io_out, out_bytes));
    h->io_out_cap = out_bytes;
  }
  float* din = h->io_in;
  char* dout = h->io_out;
  int32_t* d_tok = reinterpret_cast<int32_t*>(dout);
  float* d_lp = reinterpret_cast<float*>(dout + n_out * 4);
  int32_t* d_len = reinterpret_cast<int32_t*>(dout + n_out * 8);
  // pinned host images of an end-to-end model: copied chunk by chunk inside the call's graph, overlapping the Swin chunks;
  // anything else (pageable memory, features-in model) is copied up front
  cudaPointerAttributes pa{};
  const bool pinned = cudaPointerGetAttributes(&pa, input_host) == cudaSuccess && pa.type == cudaMemoryTypeHost;
  (void)cudaGetLastError();
  if (pinned && c.has_swin && !h->profile) {
    if (int r = beam_search_impl(h, nullptr, input_host, B, nullptr, beam, max_len, how_many, sos_idx, eos_idx, d_tok, d_len, d_lp, stream)) return r;
  } else {
    CU(cudaMemcpyAsync(din, input_host, in_elems * 4, cudaMemcpyHostToDevice, st));
    if (int r = xn_beam_search(h, din, B, nullptr, beam, max_len, how_many, sos_idx, eos_idx, d_tok, d_len, d_lp, stream)) return r;
  }
  CU(cudaMemcpyAsync(out_tokens_host, d_tok, n_out * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out_len_host, d_len, (size_t)B * how_many * 4, cudaMemcpyDeviceToHost, st));
  if (out_logprob_host) CU(cudaMemcpyAsync(out_logprob_host, d_lp, n_out * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return XN_OK;
}

// Coefficient tables are cached per (input size, output size), least-recently-used first.  The entry is returned BY VALUE
// and `pinned` (the other table of the same call) is never evicted, so a later insertion cannot free what a pending
// launch of this call reads; an evicted table may still be read by launches already enqueued, hence the device sync.
static int resample_table(xn_handle* h, int in_size, int out_size, xn_handle::ResampleTable* out,
                          const xn_handle::ResampleTable* pinned = nullptr) {
  for (size_t i = 0; i < h->rtables.size(); ++i)
    if (h->rtables[i].in_size == in_size && h->rtables[i].out_size == out_size) {
      const xn_handle::ResampleTable t = h->rtables[i];
      h->rtables.erase(h->rtables.begin() + i);            // move to the most-recently-used end
      h->rtables.push_back(t);
      *out = t;
      return 0;
    }
  std::vector<int> bounds, kk;
  xn_handle::ResampleTable t{in_size, out_size, 0, nullptr, nullptr};
  resample_coeffs(in_size, out_size, bounds, kk, &t.ksize);
  CU(cudaMalloc(&t.bounds, bounds.size() * sizeof(int)));
  CU(cudaMalloc(&t.kk, kk.size() * sizeof(int)));
  CU(cudaMemcpy(t.bounds, bounds.data(), bounds.size() * sizeof(int), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(t.kk, kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (h->rtables.size() >= xn_handle::kMaxResampleTables) {
    size_t victim = 0;                                      // least recently used entry that this call does not hold
    if (pinned && h->rtables[victim].bounds == pinned->bounds) victim = 1;
    CU(cudaDeviceSynchronize());
    cudaFree(h->rtables[victim].bounds); cudaFree(h->rtables[victim].kk);
    h->rtables.erase(h->rtables.begin() + victim);
  }
  h->rtables.push_back(t);
  *out = t;
  return 0;
}

int xn_preprocess_rgb8(xn_handle* h, const uint8_t* rgb, int rgb_on_device, int H, int W, float* out, int out_size, void* stream) {
  if (!h || !rgb || !out) return XN_ERR_ARG;
  cudaSetDevice(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  h->cur_st = st;
  if (H < 1 || W < 1 || out_size < 1 || (long)H * W > (1L << 28)) return h->fail(XN_ERR_ARG, "bad image size %d x %d -> %d", H, W, out_size);
  xn_handle::ResampleTable tx{}, ty{};
  if (int r = resample_table(h, W, out_size, &tx)) return r;
  if (int r = resample_table(h, H, out_size, &ty, &tx)) return r;
  const size_t in_bytes = (size_t)H * W * 3, tmp_bytes = (size_t)H * out_size * 3;
  const size_t need = ((in_bytes + 255) & ~size_t(255)) + tmp_bytes;
  if (h->pp_cap < need) {
    if (h->pp_buf) { CU(cudaDeviceSynchronize()); cudaFree(h->pp_buf); h->pp_buf = nullptr; h->pp_cap = 0; }
    CU(cudaMalloc(&h->pp_buf, need));
    h->pp_cap = need;
  }
  const uint8_t* src = rgb;
  uint8_t* tmp = h->pp_buf + ((in_bytes + 255) & ~size_t(255));
  if (!rgb_on_device) {
    CU(cudaMemcpyAsync(h->pp_buf, rgb, in_bytes, cudaMemcpyHostToDevice, st));
    src = h->pp_buf;
  }
  KL(2, launch_preprocess_rgb8(src, H, W, out_size, tx.bounds, tx.kk, tx.ksize, ty.bounds, ty.kk, ty.ksize, tmp, out, st));
  return XN_OK;
}

// encoder output of a batch at the top of the arena (images or features in; pads for the features-in model)
static int encode_for_search(xn_handle* h, const float* input, int B, const int32_t* enc_pads_host, size_t scratch_bytes,
                             float** enc_out_p, bool* pads_p, cudaStream_t st) {
  const xn_config& c = h->cfg;
  const int Bs = (int)std::min<int64_t>(B, h->swin_chunk), Be = (int)std::min<int64_t>(B, h->enc_chunk);
  const size_t enc_bytes = ((size_t)B * c.enc_len * c.d_model * 4 + 4095) & ~size_t(255);
  const size_t feat_bytes = c.has_swin ? (((size_t)B * c.enc_len * c.feat_dim * 4 + 4095) & ~size_t(255)) : 0;
  const size_t scratch = std::max(std::max(c.has_swin ? swin_ws_bytes(c, Bs, h->precision) : 0, enc_ws_bytes(c, Be)), scratch_bytes);
  if (int r = ensure_ws(h, scratch + enc_bytes + feat_bytes, st)) return r;
  bool pads = false;
  if (enc_pads_host)
    for (int i = 0; i < B; ++i) pads |= enc_pads_host[i] != 0;
  if (c.has_swin && pads) return h->fail(XN_ERR_ARG, "End to End case have no padding");
  char* top = h->ws.base + (h->ws.cap & ~size_t(255));
  float* enc_out = reinterpret_cast<float*>(top - enc_bytes);
  float* fb = reinterpret_cast<float*>(top - enc_bytes - feat_bytes);
  h->ws.cap = (size_t)((top - enc_bytes - feat_bytes) - h->ws.base);        // the caller holds an ArenaCapGuard
  if (c.has_swin) {
    if (int r = swin_forward(h, input, B, fb, st)) return r;
    if (int r = enc_body(h, fb, B, nullptr, enc_out, st)) return r;
  } else {
    if (int r = enc_body(h, input, B, pads ? enc_pads_host : nullptr, enc_out, st)) return r;
  }
  *enc_out_p = enc_out;
  *pads_p = pads;
  return 0;
}

int xn_beam_search_sample(xn_handle* h, const float* input, int B, const int32_t* enc_pads_host, int beam, int max_len, int how_many,
                          int sos_idx, int eos_idx, uint64_t seed, int32_t* out_tokens, int32_t* out_len, float* out_logprob,
                          void* stream) {
  NEED_READY();
  if (B < 1) return h->fail(XN_ERR_ARG, "empty batch");
  if (how_many > beam) return h->fail(XN_ERR_ARG, "requested output per sequence must be lower than beam width");
  ArenaCapGuard guard(h->ws);
  float* enc_out = nullptr;
  bool pads = false;
  if (int r = encode_for_search(h, input, B, enc_pads_host, beam_ws_bytes(h->cfg, B, beam, max_len), &enc_out, &pads, st)) return r;
  BeamPlan P;
  if (int r = beam_plan(h, P, B, pads ? enc_pads_host : nullptr, beam, max_len, how_many, st, true)) return r;
  SampleOpt so; so.on = true; so.seed = seed;
  if (int r = beam_run(h, P, enc_out, B, beam, max_len, how_many, sos_idx, eos_idx, st, so)) return r;
  return beam_copy_out(h, P, B, max_len, how_many, out_tokens, out_len, out_logprob, st);
}

int xn_sample(xn_handle* h, const float* input, int B, const int32_t* enc_pads_host, int num_outputs, int max_len, int sos_idx,
              int eos_idx, uint64_t seed, int32_t* out_tokens, int32_t* out_len, float* out_logprob, void* stream) {
  NEED_READY();
  const xn_config& c = h->cfg;
  if (B < 1) return h->fail(XN_ERR_ARG, "empty batch");
  if (num_outputs < 1 || num_outputs > 8) return h->fail(XN_ERR_ARG, "num_outputs must be in [1, 8]");
  const int L = max_len + 1;                         // SOS + max_len sampled words
  if (max_len < 1 || max_len > c.max_seq_len || L > 128) return h->fail(XN_ERR_ARG, "sample_max_seq_len %d outside [1, %d]", max_len, std::min(c.max_seq_len, 127));
  ArenaCapGuard guard(h->ws);
  float* enc_out = nullptr;
  bool pads = false;
  if (int r = encode_for_search(h, input, B, enc_pads_host, beam_ws_bytes(c, B, num_outputs, L), &enc_out, &pads, st)) return r;
  BeamPlan P;
  // the plan of a beam search with beam = how_many = num_outputs: R = B * num_outputs independent rows
  {
    const int keep_max = h->cfg.max_seq_len;
    h->cfg.max_seq_len = std::max(keep_max, L);      // the history holds L tokens; only positions < max_len are decoded
    const int r = beam_plan(h, P, B, pads ? enc_pads_host : nullptr, num_outputs, L, num_outputs, st, true);
    h->cfg.max_seq_len = keep_max;
    if (r) return r;
  }
  const int R = B * num_outputs;
  if (int r = dec_project(h, P.D, enc_out, B, st)) return r;
  KL(1, launch_beam_init(P.bb, B, num_outputs, L, sos_idx, st));
  P.D.s.anc = P.bb.anc[0];
  for (int t = 1; t <= max_len; ++t) {
    if (int r = dec_step(h, P.D, t - 1, nullptr, P.bb.tokens[0], L, num_outputs, P.nv, nullptr, P.logits, c.vocab, st)) return r;
    KL(1, launch_gumbel_topk(P.logits, c.vocab, R, c.vocab, 1, seed, t, P.topv, P.topi, st));
    KL(1, launch_sample_append(P.bb, P.topv, P.topi, R, L, t, eos_idx, st));
  }
  KL(1, launch_sample_finalize(P.bb, R, L, L, P.r_tok, P.r_len, P.r_lp, st));
  return beam_copy_out(h, P, B, L, num_outputs, out_tokens, out_len, out_logprob, st);
}

int xn_preprocess_rgb8_batch(xn_handle* h, const uint8_t* const* rgb_ptrs, int rgb_on_device, const int32_t* heights,
                             const int32_t* widths, int n, float* out, int out_size, void* stream) {
  if (!h || !rgb_ptrs || !heights || !widths || !out) return XN_ERR_ARG;
  cudaSetDevice(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  h->cur_st = st;
  const int S = out_size;
  if (n < 1 || n > 65535 || S < 1) return h->fail(XN_ERR_ARG, "bad batch (%d images -> %d)", n, S);
  // ---- plan: image / intermediate staging, item records, one coefficient table per distinct input size
  size_t img_bytes = 0, tmp_bytes = 0;
  int max_h = 0;
  std::vector<std::pair<int, size_t>> sizes;           // (input size, int offset of its table) in the packed table block
  std::vector<int> tab;                                // [bounds (2S) | kk (S * ksize)] per distinct size
  std::vector<int> ksz;
  auto table_of = [&](int in_size) -> int {
    for (size_t i = 0; i < sizes.size(); ++i) if (sizes[i].first == in_size) return (int)i;
    std::vector<int> b, k;
    int ks = 0;
    resample_coeffs(in_size, S, b, k, &ks);
    sizes.push_back({in_size, tab.size()});
    ksz.push_back(ks);
    tab.insert(tab.end(), b.begin(), b.end());
    tab.insert(tab.end(), k.begin(), k.end());
    return (int)sizes.size() - 1;
  };
  std::vector<int> tx(n), ty(n);
  std::vector<size_t> src_off(n), tmp_off(n);
  for (int i = 0; i < n; ++i) {
    const int H = heights[i], W = widths[i];
    if (H < 1 || W < 1 || H > 65535 || (long)H * W > (1L << 28) || !rgb_ptrs[i]) return h->fail(XN_ERR_ARG, "bad image %d: %d x %d", i, H, W);
    tx[i] = table_of(W); ty[i] = table_of(H);
    max_h = std::max(max_h, H);
    src_off[i] = img_bytes;
    if (!rgb_on_device) img_bytes += ((size_t)H * W * 3 + 255) & ~size_t(255);
    tmp_off[i] = tmp_bytes;
    tmp_bytes += ((size_t)H * S * 3 + 255) & ~size_t(255);
  }
  const size_t items_bytes = ((size_t)n * sizeof(PreItem) + 255) & ~size_t(255);
  const size_t meta_bytes = items_bytes + tab.size() * sizeof(int);
  const size_t need = img_bytes + tmp_bytes + ((meta_bytes + 255) & ~size_t(255));
  if (h->pp_cap < need) {
    if (h->pp_buf) { CU(cudaDeviceSynchronize()); cudaFree(h->pp_buf); h->pp_buf = nullptr; h->pp_cap = 0; }
    CU(cudaMalloc(&h->pp_buf, need));
    h->pp_cap = need;
  }
  if (h->pp_host_cap < meta_bytes) {
    if (h->pp_host) { CU(cudaDeviceSynchronize()); cudaFreeHost(h->pp_host); h->pp_host = nullptr; h->pp_host_cap = 0; }
    CU(cudaMallocHost(&h->pp_host, meta_bytes));
    h->pp_host_cap = meta_bytes;
  }
  if (!h->pp_ev) CU(cudaEventCreateWithFlags(&h->pp_ev, cudaEventDisableTiming));
  else CU(cudaEventSynchronize(h->pp_ev));             // the previous batch's metadata upload has left the pinned staging
  uint8_t* d_img = h->pp_buf;
  uint8_t* d_tmp = h->pp_buf + img_bytes;
  char* d_meta = reinterpret_cast<char*>(h->pp_buf + img_bytes + tmp_bytes);
  const int* d_tab = reinterpret_cast<const int*>(d_meta + items_bytes);
  PreItem* items = reinterpret_cast<PreItem*>(h->pp_host);
  for (int i = 0; i < n; ++i) {
    PreItem& it = items[i];
    it.src = rgb_on_device ? rgb_ptrs[i] : d_img + src_off[i];
    it.tmp = d_tmp + tmp_off[i];
    it.out = out + (size_t)i * 3 * S * S;
    it.bx = d_tab + sizes[tx[i]].second; it.kx = it.bx + 2 * S; it.ksx = ksz[tx[i]];
    it.by = d_tab + sizes[ty[i]].second; it.ky = it.by + 2 * S; it.ksy = ksz[ty[i]];
    it.H = heights[i]; it.W = widths[i];
  }
  memcpy(h->pp_host + items_bytes, tab.data(), tab.size() * sizeof(int));
  CU(cudaMemcpyAsync(d_meta, h->pp_host, meta_bytes, cudaMemcpyHostToDevice, st));
  CU(cudaEventRecord(h->pp_ev, st));
  if (!rgb_on_device)
    for (int i = 0; i < n; ++i)
      CU(cudaMemcpyAsync(d_img + src_off[i], rgb_ptrs[i], (size_t)heights[i] * widths[i] * 3, cudaMemcpyHostToDevice, st));
  KL(2, launch_preprocess_rgb8_batch(reinterpret_cast<const PreItem*>(d_meta), n, max_h, S, st));
  return XN_OK;
}

int xn_jpeg_available(void) { return g_nvjpeg.load() ? 1 : 0; }

int xn_preprocess_jpeg_batch(xn_handle* h, const uint8_t* const* jpeg_ptrs_host, const int64_t* jpeg_sizes, int n, float* out,
                             int out_size, int32_t* heights_out, int32_t* widths_out, void* stream) {
  if (!h || !jpeg_ptrs_host || !jpeg_sizes || !out) return XN_ERR_ARG;
  cudaSetDevice(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  h->cur_st = st;
  if (n < 1 || n > 65535) return h->fail(XN_ERR_ARG, "bad batch (%d images)", n);
  if (!g_nvjpeg.load()) return h->fail(XN_ERR_UNSUPPORTED, "libnvjpeg could not be loaded: decode on the host (PIL) and call xn_preprocess_rgb8_batch");
  if (!h->jpg) {
    if (g_nvjpeg.create(&h->jpg) != NVJPEG_STATUS_SUCCESS) { h->jpg = nullptr; return h->fail(XN_ERR_CUDA, "nvjpegCreateSimple failed"); }
    if (g_nvjpeg.state_create(h->jpg, &h->jpg_state) != NVJPEG_STATUS_SUCCESS) return h->fail(XN_ERR_CUDA, "nvjpegJpegStateCreate failed");
  }
  std::vector<int> H(n), W(n), comps(n);
  std::vector<size_t> off(n);
  size_t total = 0;
  for (int i = 0; i < n; ++i) {
    int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t sub;
    if (!jpeg_ptrs_host[i] || jpeg_sizes[i] <= 0 ||
        g_nvjpeg.image_info(h->jpg, jpeg_ptrs_host[i], (size_t)jpeg_sizes[i], &nc, &sub, ws, hs) != NVJPEG_STATUS_SUCCESS)
      return h->fail(XN_ERR_ARG, "image %d is not a JPEG stream nvJPEG can parse", i);
    H[i] = hs[0]; W[i] = ws[0]; comps[i] = nc;
    if (H[i] < 1 || W[i] < 1 || H[i] > 65535 || (long)H[i] * W[i] > (1L << 28)) return h->fail(XN_ERR_ARG, "bad image %d: %d x %d", i, H[i], W[i]);
    off[i] = total;
    total += ((size_t)H[i] * W[i] * 3 + 255) & ~size_t(255);
    if (heights_out) heights_out[i] = H[i];
    if (widths_out) widths_out[i] = W[i];
  }
  if (h->jpg_cap < total) {
    if (h->jpg_buf) { CU(cudaDeviceSynchronize()); cudaFree(h->jpg_buf); h->jpg_buf = nullptr; h->jpg_cap = 0; }
    CU(cudaMalloc(&h->jpg_buf, total));
    h->jpg_cap = total;
  }
  std::vector<const uint8_t*> ptrs(n);
  for (int i = 0; i < n; ++i) {
    uint8_t* dst = h->jpg_buf + off[i];
    ptrs[i] = dst;
    if (comps[i] == 3) {
      nvjpegImage_t img{};
      img.channel[0] = dst;
      img.pitch[0] = (size_t)W[i] * 3;
      const nvjpegStatus_t rc = g_nvjpeg.decode(h->jpg, h->jpg_state, jpeg_ptrs_host[i], (size_t)jpeg_sizes[i], NVJPEG_OUTPUT_RGBI, &img, st);
      if (rc != NVJPEG_STATUS_SUCCESS) return h->fail(XN_ERR_CUDA, "nvjpegDecode failed on image %d (status %d)", i, (int)rc);
    } else {
      // reference utils/image_utils.py:18-19: a file whose PIL mode is not RGB (grayscale = 1 component, CMYK = 4) is
      // replaced by a blank RGB canvas of the same size
      CU(cudaMemsetAsync(dst, 0, (size_t)H[i] * W[i] * 3, st));
    }
  }
  return xn_preprocess_rgb8_batch(h, ptrs.data(), 1, H.data(), W.data(), n, out, out_size, stream);
}

int xn_overflow_flag(xn_handle* h, int* flag_out, int clear) {
  if (!h || !flag_out) return XN_ERR_ARG;
  cudaSetDevice(h->device);
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(flag_out, h->flag_dev, sizeof(int), cudaMemcpyDeviceToHost));
  if (clear) CU(cudaMemset(h->flag_dev, 0, sizeof(int)));
  return XN_OK;
}

int xn_caption_host_begin(xn_handle* h, const float* input_host, int B, int beam, int max_len, int how_many, int sos_idx,
                          int eos_idx, int32_t* out_tokens_host, int32_t* out_len_host, float* out_logprob_host, void* stream) {
  NEED_READY();
  const xn_config& c = h->cfg;
  if (!input_host || !out_tokens_host || !out_len_host || B < 1) return h->fail(XN_ERR_ARG, "null host buffer / empty batch");
  const size_t in_bytes = (c.has_swin ? (size_t)B * c.in_chans * c.img_size * c.img_size : (size_t)B * c.enc_len * c.feat_dim) * 4;
  const size_t n_out = (size_t)B * how_many * max_len;
  const size_t out_bytes = n_out * 8 + (size_t)B * how_many * 4 + 1024;
  if (!h->hstream) CU(cudaStreamCreateWithFlags(&h->hstream, cudaStreamNonBlocking));
  const int k = h->next_slot;
  xn_handle::HostSlot& sl = h->slots[k];
  if (!sl.ev_in) {
    CU(cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
  }
  if (sl.pending) return h->fail(XN_ERR_STATE, "caption_host slot %d still has an un-ended call (at most two calls in flight)", k);
  if (sl.in_cap < in_bytes || sl.out_cap < out_bytes) {
    if (sl.in || sl.out) {                            // replacing a buffer that cached graphs may read: drop them
      CU(cudaDeviceSynchronize());
      h->drop_graphs();
    }
    if (sl.in_cap < in_bytes) { if (sl.in) cudaFree(sl.in); sl.in = nullptr; sl.in_cap = 0; CU(cudaMalloc(&sl.in, in_bytes)); sl.in_cap = in_bytes; }
    if (sl.out_cap < out_bytes) { if (sl.out) cudaFree(sl.out); sl.out = nullptr; sl.out_cap = 0; CU(cudaMalloc(&sl.out, out_bytes)); sl.out_cap = out_bytes; }
  }
  // copy stream: after the slot's previous user has finished computing (device-side order, no host wait)
  CU(cudaStreamWaitEvent(h->hstream, sl.ev_done, 0));
  CU(cudaMemcpyAsync(sl.in, input_host, in_bytes, cudaMemcpyHostToDevice, h->hstream));
  CU(cudaEventRecord(sl.ev_in, h->hstream));
  CU(cudaStreamWaitEvent(st, sl.ev_in, 0));
  int32_t* d_tok = reinterpret_cast<int32_t*>(sl.out);
  float* d_lp = reinterpret_cast<float*>(sl.out + n_out * 4);
  int32_t* d_len = reinterpret_cast<int32_t*>(sl.out + n_out * 8);
  h->staged_input = sl.in;                          // beam_search_impl reads the slot's staging directly (no second copy)
  const int r = beam_search_impl(h, sl.in, nullptr, B, nullptr, beam, max_len, how_many, sos_idx, eos_idx, d_tok, d_len, d_lp, stream);
  h->staged_input = nullptr;
  if (r) return r;
  CU(cudaMemcpyAsync(out_tokens_host, d_tok, n_out * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out_len_host, d_len, (size_t)B * how_many * 4, cudaMemcpyDeviceToHost, st));
  if (out_logprob_host) CU(cudaMemcpyAsync(out_logprob_host, d_lp, n_out * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaEventRecord(sl.ev_done, st));
  sl.pending = true;
  h->next_slot = k ^ 1;
  return k;                                         // ticket
}

int xn_caption_host_end(xn_handle* h, int ticket) {
  if (!h || ticket < 0 || ticket > 1) return XN_ERR_ARG;
  cudaSetDevice(h->device);
  xn_handle::HostSlot& sl = h->slots[ticket];
  if (!sl.pending) return h->fail(XN_ERR_STATE, "caption_host ticket %d is not in flight", ticket);
  CU(cudaEventSynchronize(sl.ev_done));
  sl.pending = false;
  return XN_OK;
}

int64_t xn_kernel_launches(const xn_handle* h) { return h ? h->launches : 0; }
int64_t xn_workspace_bytes(const xn_handle* h) { return h ? (int64_t)h->ws.cap : 0; }

int xn_set_option(xn_handle* h, const char* name, int64_t value) {
  if (!h || !name) return XN_ERR_ARG;
  std::string n(name);
  if (n == "use_graph") { h->use_graph = value; h->drop_graphs(); return XN_OK; }
  if (n == "op_out16") { h->op_out16 = value; return XN_OK; }
  if (n == "use_skinny") { h->use_skinny = value; h->drop_graphs(); return XN_OK; }
  if (n == "use_mega") { h->use_mega = value; h->drop_graphs(); return XN_OK; }
  if (n == "mega_search") { h->mega_search = value; h->drop_graphs(); return XN_OK; }
  if (n == "fuse_topk") { h->fuse_topk = value; h->drop_graphs(); return XN_OK; }
  if (n == "mega_dbg") {
    if (value && !h->mega_dbg) { if (cudaMalloc(&h->mega_dbg, 128 * 8) != cudaSuccess) return h->fail(XN_ERR_CUDA, "cudaMalloc failed"); cudaMemset(h->mega_dbg, 0, 128 * 8); }
    if (!value && h->mega_dbg) { cudaDeviceSynchronize(); cudaFree(h->mega_dbg); h->mega_dbg = nullptr; }
    h->drop_graphs();
    return XN_OK;
  }
  if (n == "mega_dbg_mode") { h->mega_dbg_mode = value; h->drop_graphs(); return XN_OK; }
  if (n == "mega_coop") { g_mega_coop = value != 0; h->drop_graphs(); return XN_OK; }
  if (n == "ln_on_load") { h->ln_on_load = value; h->drop_graphs(); return XN_OK; }
  if (n == "dec_splitk") { h->dec_splitk = value; h->drop_graphs(); return XN_OK; }
  if (n == "pe_tc") { h->pe_tc = value; h->drop_graphs(); return XN_OK; }
  if (n == "se_tc") { h->se_tc = value; h->drop_graphs(); return XN_OK; }
  if (n == "ln_fuse") { h->ln_fuse = value; h->drop_graphs(); return XN_OK; }
  if (n == "decode_groups") { h->decode_groups = std::max<int64_t>(0, std::min<int64_t>(value, xn_handle::kMaxDecodeGroups)); h->drop_graphs(); return XN_OK; }
  if (n == "pdl") { g_pdl_enabled = value != 0; h->drop_graphs(); return XN_OK; }
  if (n == "tc_debug") { set_tc_debug((int)value); return XN_OK; }
  if (n == "tc_pair") { set_tc_pair((int)value); h->drop_graphs(); return XN_OK; }
  if (n == "attn_tc") { g_attn_tc = value != 0; h->drop_graphs(); return XN_OK; }
  if (n == "early_exit") { h->early_exit = value; h->drop_graphs(); return XN_OK; }
  if (n == "attn_tc_dbg") { g_attn_tc_dbg = (int)value; h->drop_graphs(); return XN_OK; }
  if (n == "profile") {
    h->profile = value; h->prof_used = 0; h->prof_flops.clear();
    for (auto& sp : h->spans) { h->span_pool.push_back(sp.e0); h->span_pool.push_back(sp.e1); }
    h->spans.clear();
  }
  else if (n == "swin_chunk") { h->swin_chunk = std::max<int64_t>(1, value); h->drop_graphs(); }
  else if (n == "enc_chunk") { h->enc_chunk = std::max<int64_t>(1, value); h->drop_graphs(); }
  else return h->fail(XN_ERR_ARG, "unknown option '%s'", name);
  return XN_OK;
}

int xn_mega_timeline(xn_handle* h, uint64_t* out, int cap) {
  if (!h || !out || cap < 1) return XN_ERR_ARG;
  if (!h->mega_dbg) return h->fail(XN_ERR_STATE, "option mega_dbg is off");
  CU(cudaDeviceSynchronize());
  unsigned long long buf[128];
  CU(cudaMemcpy(buf, h->mega_dbg, sizeof buf, cudaMemcpyDeviceToHost));
  const int n = (int)std::min<unsigned long long>(buf[127], 64);
  for (int i = 0; i < cap; ++i) out[i] = i < n ? buf[i] : (i >= 64 && i < (int)std::min<unsigned long long>(buf[126], 124) ? buf[i] : (i == 124 || i == 125 ? buf[i] : 0));
  return n;
}

int xn_profile_read(xn_handle* h, double* ms_total, double* flops_total, int64_t* count) {
  return xn_profile_read_min(h, 0.0, ms_total, flops_total, count);
}

int xn_profile_read_min(xn_handle* h, double min_flops, double* ms_total, double* flops_total, int64_t* count) {
  if (!h) return XN_ERR_ARG;
  cudaSetDevice(h->device);
  CU(cudaDeviceSynchronize());
  double ms = 0, fl = 0;
  int64_t n_tc = 0;
  for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
    if (h->prof_flops[i / 2] < 0) continue;          // skinny mma.sync launches are not part of the tcgen05 roofline
    if (h->prof_flops[i / 2] < min_flops) continue;
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, h->prof_ev[i], h->prof_ev[i + 1]));
    ms += t;
    fl += h->prof_flops[i / 2];
    ++n_tc;
  }
  if (ms_total) *ms_total = ms;
  if (flops_total) *flops_total = fl;
  if (count) *count = n_tc;
  return XN_OK;
}

int xn_profile_kernels(xn_handle* h, char* buf, int cap) {
  if (!h || !buf || cap <= 0) return XN_ERR_ARG;
  cudaSetDevice(h->device);
  CU(cudaDeviceSynchronize());
  // aggregate by launcher name: the text of the launch expression up to its argument list
  std::vector<std::string> names;
  std::vector<double> ms;
  std::vector<int64_t> cnt;
  for (auto& sp : h->spans) {
    std::string n(sp.label);
    size_t b = n.find_first_not_of("( ");
    n = n.substr(b == std::string::npos ? 0 : b);
    n = n.substr(0, n.find('('));
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, sp.e0, sp.e1));
    size_t i = 0;
    for (; i < names.size(); ++i) if (names[i] == n) break;
    if (i == names.size()) { names.push_back(n); ms.push_back(0); cnt.push_back(0); }
    ms[i] += t; cnt[i] += 1;
  }
  std::string out;
  for (size_t i = 0; i < names.size(); ++i) {
    char line[256];
    snprintf(line, sizeof line, "%s\t%lld\t%.6f\n", names[i].c_str(), (long long)cnt[i], ms[i]);
    out += line;
  }
  if ((int)out.size() + 1 > cap) return h->fail(XN_ERR_ARG, "profile buffer too small: need %zu", out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return XN_OK;
}

// ---- single-operator entry points ------------------------------------------------------------
#define OP_READY()                                   \
  if (!h) return XN_ERR_ARG;                         \
  cudaSetDevice(h->device);                          \
  cudaStream_t st = (cudaStream_t)stream;            \
  h->cur_st = st;

int xn_op_layernorm(xn_handle* h, const float* x, const float* gamma, const float* beta, float* y, int rows, int C, void* stream) {
  OP_READY();
  KL(1, launch_layernorm<float>(x, C, gamma, beta, y, C, rows, C, st));
  return XN_OK;
}

int xn_op_linear(xn_handle* h, const float* x, const float* w, const float* bias, const float* residual, float* y, int M, int N,
                 int K, int act, int precision, void* stream) {
  OP_READY();
  if (precision == XN_PREC_FP32) {
    LinW l; l.w = w; l.b = bias; l.N = N; l.K = K;
    return lin_f32(h, x, K, l, residual, N, y, N, M, act, st);
  }
  if (!tc_gemm_supported(M, N, K)) return h->fail(XN_ERR_UNSUPPORTED, "tcgen05 GEMM needs K %% 64 == 0 and N %% 8 == 0 (M=%d N=%d K=%d)", M, N, K);
  if (int r = ensure_ws(h, ((size_t)M * K + (size_t)N * K + (size_t)M * N) * 2 + 8192, st)) return r;
  bf16* xb = h->ws.get<bf16>((size_t)M * K);
  bf16* wb = h->ws.get<bf16>((size_t)N * K);
  const int fp16 = precision == XN_PREC_FP16;
  if (fp16) {
    KL(1, launch_cast<f16>(x, reinterpret_cast<f16*>(xb), (long)M * K, st));
    KL(1, launch_cast<f16>(w, reinterpret_cast<f16*>(wb), (long)N * K, st));
  } else {
    KL(1, launch_cast<bf16>(x, xb, (long)M * K, st));
    KL(1, launch_cast<bf16>(w, wb, (long)N * K, st));
  }
  LinW l; l.wb = wb; l.b = bias; l.N = N; l.K = K;
  if (h->op_out16) {            // exercise the 16-bit-output epilogue, then widen for the caller
    if (h->ws.off + (size_t)M * N * 2 + 512 > h->ws.cap) return h->fail(XN_ERR_STATE, "op_out16: workspace too small");
    bf16* y16 = h->ws.get<bf16>((size_t)M * N);
    if (int r = lin_tc(h, xb, K, l, residual, N, nullptr, y16, N, M, act, fp16, st, 0)) return r;
    KL(1, launch_widen_16(y16, y, (long)M * N, fp16, st));
    return XN_OK;
  }
  return lin_tc(h, xb, K, l, residual, N, y, nullptr, N, M, act, fp16, st, 0);
}

int xn_op_linear_skinny(xn_handle* h, const float* x, const float* gamma, const float* beta, const float* w, const float* bias,
                        const float* residual, float* y, int M, int N, int K, int act, int x_is_16bit, int precision, void* stream) {
  OP_READY();
  if (precision != XN_PREC_FP16 && precision != XN_PREC_BF16) return h->fail(XN_ERR_ARG, "skinny GEMM is a 16-bit-operand kernel");
  if (x_is_16bit && gamma) return h->fail(XN_ERR_ARG, "LayerNorm fusion takes the fp32 rows");
  if (int r = ensure_ws(h, ((size_t)M * K + (size_t)N * K) * 2 + 8192, st)) return r;
  bf16* xb = h->ws.get<bf16>((size_t)M * K);
  bf16* wb = h->ws.get<bf16>((size_t)N * K);
  const bool fp16 = precision == XN_PREC_FP16;
  if (fp16) KL(1, launch_cast<f16>(w, reinterpret_cast<f16*>(wb), (long)N * K, st));
  else KL(1, launch_cast<bf16>(w, wb, (long)N * K, st));
  if (x_is_16bit) {
    if (fp16) KL(1, launch_cast<f16>(x, reinterpret_cast<f16*>(xb), (long)M * K, st));
    else KL(1, launch_cast<bf16>(x, xb, (long)M * K, st));
  }
  SkinnyArgs g{};
  g.A16 = x_is_16bit ? xb : nullptr; g.A32 = x_is_16bit ? nullptr : x; g.lda = K; g.ln_g = gamma; g.ln_b = beta;
  g.W = wb; g.ldw = K; g.bias = bias; g.res = residual; g.ldr = N; g.Cf = y; g.Cb = nullptr; g.ldc = N;
  g.M = M; g.N = N; g.K = K; g.act = act;
  if (!skinny_gemm_supported(g)) return h->fail(XN_ERR_UNSUPPORTED, "skinny GEMM does not cover M=%d N=%d K=%d", M, N, K);
  if (fp16) KL(1, launch_gemm_skinny<f16>(g, st));
  else KL(1, launch_gemm_skinny<bf16>(g, st));
  return XN_OK;
}

int xn_op_gemm_raw(xn_handle* h, int which, const void* a, const float* gamma, const float* beta, const void* w16, const float* bias,
                   const float* residual, float* y, int M, int N, int K, int act, int precision, void* stream) {
  OP_READY();
  const bool fp16 = precision == XN_PREC_FP16;
  if (which == 0 || which == 3) {
    LinW l; l.wb = w16; l.b = bias; l.N = N; l.K = K;
    if (which == 3) {            // LayerNorm-on-load: a = fp32 rows
      if (!tc_gemm_ln_supported(M, N, K)) return h->fail(XN_ERR_UNSUPPORTED, "LayerNorm-on-load GEMM does not cover M=%d N=%d K=%d", M, N, K);
      return lin_tc(h, nullptr, 0, l, residual, N, y, nullptr, N, M, act, fp16, st, 1, reinterpret_cast<const float*>(a), K, gamma, beta);
    }
    return lin_tc(h, a, K, l, residual, N, y, nullptr, N, M, act, fp16, st);
  }
  SkinnyArgs g{};
  g.A16 = which == 1 ? a : nullptr; g.A32 = which == 2 ? reinterpret_cast<const float*>(a) : nullptr; g.lda = K;
  g.ln_g = which == 2 ? gamma : nullptr; g.ln_b = which == 2 ? beta : nullptr;
  g.W = w16; g.ldw = K; g.bias = bias; g.res = residual; g.ldr = N; g.Cf = y; g.Cb = nullptr; g.ldc = N;
  g.M = M; g.N = N; g.K = K; g.act = act;
  if (!skinny_gemm_supported(g)) return h->fail(XN_ERR_UNSUPPORTED, "skinny GEMM does not cover M=%d N=%d K=%d", M, N, K);
  if (fp16) KL(1, launch_gemm_skinny<f16>(g, st));
  else KL(1, launch_gemm_skinny<bf16>(g, st));
  return XN_OK;
}

int xn_op_window_attention(xn_handle* h, const float* qkv, const float* bias_table, float* out, int B, int H, int C, int heads,
                           int shift, int precision, void* stream) {
  OP_READY();
  if (precision == XN_PREC_FP32) {
    KL(1, launch_window_attention<float>(qkv, bias_table, out, B, H, C, heads, shift, st));
    return XN_OK;
  }
  const size_t n = (size_t)B * H * H * C;
  if (int r = ensure_ws(h, n * 4 * 2 + 8192 + bias_derived_floats(heads) * 4, st)) return r;
  bf16* qb = h->ws.get<bf16>(3 * n);
  bf16* ob = h->ws.get<bf16>(n);
  float* bias_t = h->ws.get<float>(bias_derived_floats(heads));
  KL(1, launch_transpose_bias(bias_table, bias_t, heads, st));
  bias_table = bias_t;
  if (precision == XN_PREC_FP16) {
    KL(1, launch_cast<f16>(qkv, reinterpret_cast<f16*>(qb), (long)(3 * n), st));
    KL(1, ActOps<f16>::attn(reinterpret_cast<f16*>(qb), bias_table, reinterpret_cast<f16*>(ob), B, H, C, heads, shift, st));
  } else {
    KL(1, launch_cast<bf16>(qkv, qb, (long)(3 * n), st));
    KL(1, ActOps<bf16>::attn(qb, bias_table, ob, B, H, C, heads, shift, st));
  }
  KL(1, launch_widen_16(ob, out, (long)n, precision == XN_PREC_FP16, st));
  return XN_OK;
}

int xn_op_logsoftmax_topk(xn_handle* h, const float* logits, int rows, int V, int k, float* top_val, int32_t* top_idx,
                          float* logprob_or_null, void* stream) {
  OP_READY();
  KL(1, launch_logsoftmax_topk(logits, V, rows, V, k, top_val, top_idx, logprob_or_null, V, logprob_or_null ? 1 : 0, st));
  return XN_OK;
}

}  // extern "C"

__global__ void nonfinite_flag_kernel(const float* __restrict__ x, long n, int* __restrict__ flag) {
  bool bad = false;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long)gridDim.x * blockDim.x * 4) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
    } else {
      for (long j = i; j < n; ++j) bad |= !isfinite(x[j]);
    }
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
cudaError_t launch_nonfinite_flag(const float* x, long n, int* flag, cudaStream_t st) {
  const long blocks = std::min<long>(148 * 8, (n / 4 + 255) / 256 + 1);
  nonfinite_flag_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, flag);
  return cudaGetLastError();
}

__global__ void widen_16_kernel(const void* __restrict__ x, float* __restrict__ y, long n, int fp16) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fp16 ? __half2float(reinterpret_cast<const f16*>(x)[i]) : __bfloat162float(reinterpret_cast<const bf16*>(x)[i]);
}
cudaError_t launch_widen_16(const void* x, float* y, long n, int fp16, cudaStream_t st) {
  widen_16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, n, fp16);
  return cudaGetLastError();
}
