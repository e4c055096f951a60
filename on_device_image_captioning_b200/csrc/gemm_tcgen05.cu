// 16-bit (fp16 by default, or bf16) GEMM on the 5th-generation tensor cores: tcgen05.mma with the accumulator in TMEM,
// operands staged in shared memory by TMA (128B swizzle), warp-specialised and persistent.
//
//   C (M x N) = act( A (M x K) . W^T (N x K) / div + bias ) + res
//
// A and W are fp16 or bf16 (ep.fp16) with K contiguous (nn.Linear layout on both sides), accumulation is fp32.
// This is the contraction behind every nn.Linear of the path in the 16-bit modes, and (batched: TcGemmArgs::batch) the
// per-image contractions of the static-expansion block
// (reference models/swin_transformer_mod.py:214-216,229-233,270 qkv/proj; :109-119 fc1/fc2;
// :479,499 patch-merging reduction).
//
// CTA = 18 warps: warp 0  TMA producer (one elected lane)
//                 warp 1  TMEM allocator + MMA issuer (one elected lane)
//                 warps 2-17 epilogue (four per TMEM lane quarter, a quarter of the columns each; the epilogue is
//                            latency-bound per warp, so it wants many warps in flight):
//                            tcgen05.ld (16 columns, double-buffered) -> smem transpose -> fused
//                            bias/GELU/residual (residual prefetched one step ahead) -> coalesced stores
// Pipelines: a kStages-deep smem ring (full/empty mbarriers, TMA <-> MMA) and a 2-deep TMEM
// accumulator ring (tmem_full/tmem_empty, MMA <-> epilogue) so the epilogue of tile i overlaps
// the MMAs of tile i+1.  Tiles are 128 x BN (BN in {128,192,256}), K step 64 (one 128-byte
// swizzle atom), UMMA shape 128 x BN x 16, cta_group::1.
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>
#include "kernels.h"
#include "common.cuh"

namespace xn {

constexpr int kBM = 128;
constexpr int kBK = 64;                       // bf16 elements = 128 bytes = swizzle span
constexpr int kUmmaK = 16;
constexpr int kEpiWarps = 16;                 // 4 per TMEM lane quarter, each draining a quarter of the tile's columns
constexpr int kTcThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kABytes = kBM * kBK * 2;   // 16 KB

// CTAS = 2: a CTA pair (cluster of two, the two SMs of a TPC) shares one 256 x BN tile through tcgen05.mma.cta_group::2.
// Each CTA stages its own 128 rows of A and only HALF of the W tile, which cuts the shared-memory traffic per MMA by a
// third (the single-CTA kernel measured 66-71 % tensor-pipe utilisation with TMA writes + MMA reads at the smem limit).
template <int BN, int CTAS> struct TcCfg {
  static constexpr uint32_t kBBytes = (BN / CTAS) * kBK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kStageBytes <= 32768) ? 6 : 4;
  static constexpr uint32_t kAccCols = 256;                 // column stride between the two accumulators
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kStagingBytes = kEpiWarps * 32 * 16 * 4;   // per epilogue warp: 32 rows x 16 floats, XOR-swizzled 16-B slots
  static constexpr uint32_t kBarBytes = 384;   // pipeline + accumulator barriers, TMEM pointer, 16 residual-slab barriers
  static constexpr uint32_t kBiasBytes = 2 * BN * 4;          // the tile's bias row, double-buffered by accumulator parity
  // no alignment slack: the kernel has no static smem, so the dynamic window starts 1024-aligned (checked on device)
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarBytes + kBiasBytes;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");
};

struct TcEpilogue {
  float* Cf; void* Cb; long ldc;
  int fp16;
  const float* bias;
  const float* res; long ldr;
  float scale;  // accumulator multiplier (1 / div); a multiply, never a speculated division
  int w_static;        // weight tiles may be loaded before the dependency wait
  const float* a32; long lda32; const float* ln_g; const float* ln_b;   // LayerNorm-on-load source of A (or nullptr)
  float* stats_out; void* x16_out; long ldx16;                          // producer side of the folded LayerNorm (fp32 output path)
  float* stats_zero;                                                    // the OTHER statistics buffer: its rows are zeroed here (fp32 output path)
  const float* ln_stats; float ln_inv_k;                                // consumer side (16-bit output path)
  int use_tma_store;   // 16-bit output without residual: write through TMA (needs ldc % 8 == 0)
  int dbg;     // timing experiments only: 1 = skip the epilogue's global traffic, 2 = skip MMA issue, 4 = skip TMA loads
  int batch; long sC, sR;   // batched mode (3-D operand maps, generic epilogue): C / res element strides between problems
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// CTA-pair variants: the transaction bytes of both CTAs' loads land on the LEADER's barrier (shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {     // shared::cluster address of a peer's smem
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  return ra;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {     // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// GELU(erf) for outputs that are rounded to 16 bits anyway:  gelu(x) = x * Phi(x) = x * sigmoid(u) = 0.5 x (1 + tanh(u / 2))
// with u = logit(Phi(x)), an odd function fitted by x * (a0 + a1 x^2 + a2 x^4) on |x| <= 5 (x^2 clamped beyond, where the
// sigmoid is saturated): max |error| of the fit 3.0e-5 absolute (tools/fit_gelu.py), an order of magnitude below the
// output rounding; tanh.approx.f32 adds 2^-11 relative.  7 FP32 instructions + 1 MUFU op per element, against 16 + 2
// for the Abramowitz-Stegun erfc it replaced (and 7 + 2 for the ex2/rcp form of the same sigmoid): the epilogue of the
// GELU GEMMs is issue/MUFU/power-bound, so the instruction count is what matters.  Measured: fc1 launches -8..-27 %,
// fp16 Swin features 8.2e-4 rel-fro before and after (profiles/).  The fp32 parity mode keeps erff().
__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = fminf(x * x, 25.0f);
  float q = fmaf(-0.000717442621f * 0.5f, x2, 0.0741005620f * 0.5f);
  q = fmaf(q, x2, 1.59491707f * 0.5f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(q * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024)   [46,48) version = 1
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1,
// A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
// A/B format field: 0 = F16, 1 = BF16.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int ab_fmt) {
  return (1u << 4) | ((uint32_t)ab_fmt << 7) | ((uint32_t)ab_fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// OUT: 0 = fp32 output, 1 = 16-bit output (bf16, or fp16 when ep.fp16).  ACT: 0 none, 1 GELU(erf), 2 ReLU.
template <int BN, int OUT, int ACT, int CTAS>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_r, TcEpilogue ep,
               int M, int N, int K) {
  using Cfg = TcCfg<BN, CTAS>;
  constexpr int S = Cfg::kStages;
  constexpr int kTileM = kBM * CTAS;               // rows of one (pair) tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);                             // SWIZZLE_128B tiles need 1024-B alignment
  if (base & 1023u) __trap();
  uint8_t* gen_base = smem_raw;
  const uint32_t staging = base + S * Cfg::kStageBytes;
  float* staging_gen = reinterpret_cast<float*>(gen_base + S * Cfg::kStageBytes);
  const uint32_t bars = staging + Cfg::kStagingBytes;
  // barrier layout: full[S] | empty[S] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto tfull_bar = [&](int b) { return bars + 8u * (2 * S + b); };
  auto tempty_bar = [&](int b) { return bars + 8u * (2 * S + 2 + b); };
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(gen_base + S * Cfg::kStageBytes + Cfg::kStagingBytes + 8 * (2 * S + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + kTileM - 1) / kTileM, n_tiles = (N + BN - 1) / BN;
  const int tiles_pb = m_tiles * n_tiles;                              // tiles per problem
  const int total_tiles = tiles_pb * (ep.batch > 0 ? ep.batch : 1);
  const int nkb = (K + kBK - 1) / kBK;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;           // 0 = leader (issues the MMAs)
  const int tile0 = blockIdx.x / CTAS, tile_step = gridDim.x / CTAS;   // persistent loop over (pair) tiles
  const int row_off = (int)rank * kBM;                                 // this CTA's rows inside the pair tile
  // LayerNorm-on-load mode (single CTA, one tile per CTA, K = 8 k-blocks): A lives in the first 8 x 16 KB of the ring
  // area for the whole kernel, the remaining ring bytes form the B pipeline
  constexpr int kAResBlocks = 8;
  const bool a_res = CTAS == 1 && ep.a32 != nullptr;
  constexpr uint32_t kBRingOff = kAResBlocks * kABytes;
  constexpr int kSB = (int)((S * Cfg::kStageBytes - kBRingOff) / Cfg::kBBytes) > 0 ? (int)((S * Cfg::kStageBytes - kBRingOff) / Cfg::kBBytes) : 1;
  const uint32_t afull_bar = bars + 8u * (2 * S + 21);

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
    if (ep.use_tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_c)) : "memory");
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), kEpiWarps * CTAS); }
    for (int b = 0; b < 16; ++b) mbar_init(bars + 8u * (2 * S + 5 + b), 1);     // residual-slab barriers (fp32 TMA epilogue)
    mbar_init(bars + 8u * (2 * S + 21), kEpiWarps);                              // A tile produced in-kernel (LayerNorm-on-load)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_gen)),
                   "r"(Cfg::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_gen)),
                   "r"(Cfg::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();      // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  // Barrier init and the TMEM allocation above overlap the previous kernel's tail (programmatic dependent launch).
  // The weight operand does not depend on the previous kernel either: the single-CTA producer streams the W tiles of
  // its first pipeline stages BEFORE the dependency wait and adds the A tiles after it, so a latency-bound decoder-step
  // GEMM starts its MMAs one L2 round trip after the previous kernel has drained.
  int pre_stages = 0;
  if (CTAS == 1 && ep.w_static && warp == 0 && lane == 0 && tile0 < total_tiles && !(ep.dbg & 4)) {
    const int ring = a_res ? kSB : S;
    pre_stages = nkb < ring ? nkb : ring;
    const int bz0 = ep.batch > 0 ? tile0 / tiles_pb : 0;
    const int n0 = ((tile0 - bz0 * tiles_pb) % n_tiles) * BN;
    for (int kb = 0; kb < pre_stages; ++kb) {
      mbar_expect_tx(full_bar(kb), a_res ? Cfg::kBBytes : Cfg::kStageBytes);
      if (ep.batch > 0) tma_load_3d(base + kb * Cfg::kStageBytes + kABytes, &tma_b, kb * kBK, n0, bz0, full_bar(kb));
      else tma_load_2d(a_res ? base + kBRingOff + kb * Cfg::kBBytes : base + kb * Cfg::kStageBytes + kABytes, &tma_b, kb * kBK, n0, full_bar(kb));
    }
  }
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int bz = ep.batch > 0 ? tile / tiles_pb : 0, tl = tile - bz * tiles_pb;
        const int m0 = (tl / n_tiles) * kTileM + row_off, n0 = (tl % n_tiles) * BN + (int)rank * (BN / CTAS);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = base + stage * Cfg::kStageBytes;
          if (CTAS == 2) {
            // both CTAs' bytes are expected by, and complete on, the leader's barrier
            const uint32_t lbar = map_to_cta(full_bar(stage), 0u);
            if (ep.dbg & 4) { if (rank == 0) mbar_arrive(full_bar(stage)); if (++stage == S) { stage = 0; phase ^= 1u; } continue; }
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * Cfg::kStageBytes);
            tma_load_2d_pair(sa, &tma_a, kb * kBK, m0, lbar);
            tma_load_2d_pair(sa + kABytes, &tma_b, kb * kBK, n0, lbar);
          } else {
            if (ep.dbg & 4) { mbar_arrive(full_bar(stage)); if (++stage == S) { stage = 0; phase ^= 1u; } continue; }
            if (a_res) {                                // only W travels through the ring; A is written by the epilogue warps
              if (pre_stages > 0) --pre_stages;
              else {
                mbar_expect_tx(full_bar(stage), Cfg::kBBytes);
                tma_load_2d(base + kBRingOff + stage * Cfg::kBBytes, &tma_b, kb * kBK, n0, full_bar(stage));
              }
            } else if (pre_stages > 0) {                // W tile and byte count of this stage were issued before the wait
              --pre_stages;
              if (ep.batch > 0) tma_load_3d(sa, &tma_a, kb * kBK, m0, bz, full_bar(stage));
              else tma_load_2d(sa, &tma_a, kb * kBK, m0, full_bar(stage));
            } else if (ep.batch > 0) {                  // batched problems: 3-D maps (k, row, problem), edges clipped per problem
              mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
              tma_load_3d(sa, &tma_a, kb * kBK, m0, bz, full_bar(stage));
              tma_load_3d(sa + kABytes, &tma_b, kb * kBK, n0, bz, full_bar(stage));
            } else {
              mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
              tma_load_2d(sa, &tma_a, kb * kBK, m0, full_bar(stage));
              tma_load_2d(sa + kABytes, &tma_b, kb * kBK, n0, full_bar(stage));
            }
          }
          if (++stage == (a_res ? kSB : S)) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc(kTileM, BN, ep.fp16 ? 0 : 1);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(buf), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * Cfg::kAccCols;
        if (a_res) { mbar_wait(afull_bar, 0u); tc_fence_after(); }      // the normalised A tile is in shared memory
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * Cfg::kStageBytes;
          const uint64_t adesc = make_sw128_desc(a_res ? base + kb * kABytes : sa);
          const uint64_t bdesc = make_sw128_desc(a_res ? base + kBRingOff + stage * Cfg::kBBytes : sa + kABytes);
          if (!(ep.dbg & 2)) {
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {  // +32 bytes (>>4 = 2) per UMMA_K step inside the swizzle atom
              if (CTAS == 2) umma_f16_pair(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)((kb | k) != 0));
              else umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (uint32_t)((kb | k) != 0));
            }
          }
          if (CTAS == 2) umma_commit_pair(empty_bar(stage));   // frees the smem slot in both CTAs when these MMAs retire
          else umma_commit(empty_bar(stage));
          if (++stage == (a_res ? kSB : S)) { stage = 0; phase ^= 1u; }
        }
        if (CTAS == 2) umma_commit_pair(tfull_bar(buf));       // accumulator complete -> both CTAs' epilogues
        else umma_commit(tfull_bar(buf));
      }
    }
  } else {
    if (a_res) {
      // ---- LayerNorm-on-load: epilogue warp e normalises rows 8e .. 8e+7 of the tile (four at a time, all their loads in
      // flight), rounds to the operand type and writes them into the resident A area in the 128-byte-swizzled K-major
      // layout the UMMA descriptor expects: k-block kb at kb * 16 KB, row r at r * 128 B, 16-byte chunk c at c ^ (r & 7).
      // Same per-lane column assignment and reduction order as layernorm_kernel, so the statistics are bit-identical.
      const int e = warp - 2;
      const int m0t = (tile0 / n_tiles) * kBM;
      const float inv_k = 1.0f / (float)K;
      for (int hb = 0; hb < 2; ++hb) {
        float4 v[4][4];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int row = m0t + e * 8 + hb * 4 + qq;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            v[qq][i] = row < M ? *reinterpret_cast<const float4*>(ep.a32 + (long)row * ep.lda32 + (i * 32 + lane) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          float mean = 0.f, rstd = 1.f;
          if (ep.ln_g) {
            float sm_ = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) sm_ += (v[qq][i].x + v[qq][i].y) + (v[qq][i].z + v[qq][i].w);
            mean = warp_sum(sm_) * inv_k;
            float qv = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float d0 = v[qq][i].x - mean, d1 = v[qq][i].y - mean, d2 = v[qq][i].z - mean, d3 = v[qq][i].w - mean;
              qv += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
            rstd = 1.0f / sqrtf(warp_sum(qv) * inv_k + 1e-5f);
          }
          const int r = e * 8 + hb * 4 + qq;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 o = v[qq][i];
            if (ep.ln_g) {
              const int c = (i * 32 + lane) * 4;
              const float4 ga = *reinterpret_cast<const float4*>(ep.ln_g + c), be = *reinterpret_cast<const float4*>(ep.ln_b + c);
              o.x = (o.x - mean) * rstd * ga.x + be.x; o.y = (o.y - mean) * rstd * ga.y + be.y;
              o.z = (o.z - mean) * rstd * ga.z + be.z; o.w = (o.w - mean) * rstd * ga.w + be.w;
            }
            uint2 u;
            if (ep.fp16) {
              __half2 a2 = __floats2half2_rn(o.x, o.y), b2 = __floats2half2_rn(o.z, o.w);
              u.x = *reinterpret_cast<uint32_t*>(&a2); u.y = *reinterpret_cast<uint32_t*>(&b2);
            } else {
              __nv_bfloat162 a2 = __floats2bfloat162_rn(o.x, o.y), b2 = __floats2bfloat162_rn(o.z, o.w);
              u.x = *reinterpret_cast<uint32_t*>(&a2); u.y = *reinterpret_cast<uint32_t*>(&b2);
            }
            const int kb = i * 2 + (lane >> 4);                 // column (i*32+lane)*4 lies in k-block ((i*32+lane)*4) / 64
            const int chunk = (lane & 15) >> 1;                 // 16-byte chunk inside the 128-byte row of that k-block
            *reinterpret_cast<uint2*>(gen_base + kb * kABytes + r * 128 + ((chunk ^ (r & 7)) << 4) + (lane & 1) * 8) = u;
          }
        }
      }
      fence_async_smem();                                        // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(afull_bar);
    }
    const int q = warp & 3;                        // TMEM lane quarter this warp may access (hardware: warp id % 4)
    const int hf = (warp - 2) >> 2;                // which quarter of the tile's columns this warp drains
    int it = 0;
    if (OUT == 1 && ep.use_tma_store) {
      // ---- 16-bit output through TMA stores.  Lane == accumulator row: bias / activation / conversion happen in that
      // layout, the 32 x 32 (rows x columns) 16-bit slab is written to smem in the 64-byte-swizzled box layout
      // (conflict-free 16-byte stores) and one elected lane hands it to the TMA, which also clips at the M / N edges.
      const int cq = hf;                                   // 64-column group of this warp
      uint8_t* slab_gen = reinterpret_cast<uint8_t*>(staging_gen) + (warp - 2) * 2048;
      const uint32_t slab_u32 = staging + (uint32_t)(warp - 2) * 2048u;
      float* bias_all = reinterpret_cast<float*>(gen_base + S * Cfg::kStageBytes + Cfg::kStagingBytes + Cfg::kBarBytes);
      const bool active = cq * 64 < BN;
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        const int m0 = (tile / n_tiles) * kTileM + row_off, n0 = (tile % n_tiles) * BN;
        const int col_base = n0 + cq * 64;
        // this warp's 64 bias values: fetched before the accumulator wait (latency hidden), published after it.  The
        // buffer is shared by the four warps of a column group (they write identical values) and double-buffered by
        // accumulator parity: once tile i's accumulator is full, every warp has left tile i-2.
        float b_lo = 0.f, b_hi = 0.f;
        if (active && ep.bias) {
          if (col_base + lane < N) b_lo = ep.bias[col_base + lane];
          if (col_base + lane + 32 < N) b_hi = ep.bias[col_base + lane + 32];
        }
        // folded LayerNorm: this lane's row scale 1/std (lane == accumulator row).  The weights were centred along K at load
        // time, so the row mean has already dropped out of the accumulator (kernels.h: TcGemmArgs::ln_stats).
        float ln_a = ep.scale;
        if (ep.ln_stats != nullptr && active) {
          const int row = m0 + q * 32 + lane;
          if (row < M) {
            const longlong2 st2 = *reinterpret_cast<const longlong2*>(reinterpret_cast<const long long*>(ep.ln_stats) + 2 * (long)row);
            const float mean = __ll2float_rn(st2.x) * (1.0f / 16777216.0f) * ep.ln_inv_k;
            ln_a = 1.0f / sqrtf(fmaxf(__ll2float_rn(st2.y) * (1.0f / 65536.0f) * ep.ln_inv_k - mean * mean, 0.f) + 1e-5f);
          }
        }
        float* bias_s = bias_all + buf * BN + cq * 64;
        mbar_wait(tfull_bar(buf), acc_phase);
        tc_fence_after();
        if (active) { bias_s[lane] = b_lo; bias_s[lane + 32] = b_hi; }
        __syncwarp();
        if (active && col_base < N) {
          const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + buf * Cfg::kAccCols + cq * 64;
          uint32_t v[2][32];
          tmem_ld32_nowait(t_base, v[0]);
          __syncwarp();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tmem_ld_wait();
            if (half == 0) tmem_ld32_nowait(t_base + 32, v[1]);
            if (ep.dbg & 1) continue;
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + half * 32 + c);
              float t0 = fmaf(__uint_as_float(v[half][c]), ln_a, b4.x), t1 = fmaf(__uint_as_float(v[half][c + 1]), ln_a, b4.y);
              float t2 = fmaf(__uint_as_float(v[half][c + 2]), ln_a, b4.z), t3 = fmaf(__uint_as_float(v[half][c + 3]), ln_a, b4.w);
              if (ACT == 1) { t0 = gelu_fast(t0); t1 = gelu_fast(t1); t2 = gelu_fast(t2); t3 = gelu_fast(t3); }
              else if (ACT == 2) { t0 = fmaxf(t0, 0.f); t1 = fmaxf(t1, 0.f); t2 = fmaxf(t2, 0.f); t3 = fmaxf(t3, 0.f); }
              if (ep.fp16) {
                __half2 a2 = __floats2half2_rn(t0, t1), b2 = __floats2half2_rn(t2, t3);
                pk[c >> 1] = *reinterpret_cast<uint32_t*>(&a2); pk[(c >> 1) + 1] = *reinterpret_cast<uint32_t*>(&b2);
              } else {
                __nv_bfloat162 a2 = __floats2bfloat162_rn(t0, t1), b2 = __floats2bfloat162_rn(t2, t3);
                pk[c >> 1] = *reinterpret_cast<uint32_t*>(&a2); pk[(c >> 1) + 1] = *reinterpret_cast<uint32_t*>(&b2);
              }
            }
            if (lane == 0) tma_store_wait_read();           // the previous slab has been read out of smem
            __syncwarp();
            // row = lane, 64 bytes per row; 16-byte chunk j lands at j ^ ((row >> 1) & 3)  (CU_TENSOR_MAP_SWIZZLE_64B)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(slab_gen + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                  make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            fence_async_smem();
            __syncwarp();
            if (lane == 0 && col_base + half * 32 < N) tma_store_2d(&tma_c, slab_u32, col_base + half * 32, m0 + q * 32);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CTAS == 2) mbar_arrive_cluster(map_to_cta(tempty_bar(buf), 0u)); else mbar_arrive(tempty_bar(buf)); }
      }
      if (lane == 0) tma_store_wait_all();
    } else if (OUT == 0 && ep.use_tma_store) {
      // ---- fp32 output (+ fp32 residual) through the TMA.  Eight warps (two per TMEM lane quarter, half of the columns
      // each) own two 2-KB slabs: the residual box (32 rows x 16 columns) of sub-step s+1 is TMA-loaded into one slab
      // while sub-step s is computed in place in the other (lane == row, 64 swizzled bytes per row) and TMA-stored.
      const int e = warp - 2;
      const bool active = e < 8;
      const int hf2 = (e >> 2) & 1;
      constexpr int kSub = (BN / 2) / 16;
      float* bias_all = reinterpret_cast<float*>(gen_base + S * Cfg::kStageBytes + Cfg::kStagingBytes + Cfg::kBarBytes);
      uint8_t* slab_gen[2] = {reinterpret_cast<uint8_t*>(staging_gen) + (e & 7) * 2048, reinterpret_cast<uint8_t*>(staging_gen) + ((e & 7) + 8) * 2048};
      const uint32_t slab_u32[2] = {staging + (uint32_t)(e & 7) * 2048u, staging + (uint32_t)((e & 7) + 8) * 2048u};
      const uint32_t rbar[2] = {bars + 8u * (2 * S + 5 + 2 * (e & 7)), bars + 8u * (2 * S + 5 + 2 * (e & 7) + 1)};
      uint32_t rphase[2] = {0u, 0u};
      const bool has_res = ep.res != nullptr;
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        const int m0 = (tile / n_tiles) * kTileM + row_off, n0 = (tile % n_tiles) * BN;
        const int col_base = n0 + hf2 * (BN / 2);
        const int row0 = m0 + q * 32;
        float bpre[4] = {0.f, 0.f, 0.f, 0.f};
        if (active) {
          if (ep.bias) {
#pragma unroll
            for (int j = 0; j < (BN / 2 + 31) / 32; ++j) {
              const int cc = lane + 32 * j;
              if (cc < BN / 2 && col_base + cc < N) bpre[j] = ep.bias[col_base + cc];
            }
          }
          if (has_res && lane == 0 && col_base < N) {                   // residual of sub-step 0, before the accumulator wait
            tma_store_wait_read();
            mbar_expect_tx(rbar[0], 2048);
            tma_load_2d(slab_u32[0], &tma_r, col_base, row0, rbar[0]);
          }
        }
        mbar_wait(tfull_bar(buf), acc_phase);
        tc_fence_after();
        if (active) {
          float* bias_s = bias_all + buf * BN + hf2 * (BN / 2);
#pragma unroll
          for (int j = 0; j < (BN / 2 + 31) / 32; ++j)
            if (lane + 32 * j < BN / 2) bias_s[lane + 32 * j] = bpre[j];
          __syncwarp();
          const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + buf * Cfg::kAccCols + hf2 * (BN / 2);
          uint32_t v[2][16];
          tmem_ld16_nowait(t_base, v[0]);
          // folded LayerNorm, producer side: this row's partial statistics.  Every aligned group of 16 columns is summed in
          // fp32 in a fixed order and then accumulated in 64-bit fixed point (2^-24 / 2^-16 units): integer adds commute, so
          // the statistics -- and with them every caption -- depend neither on the tile shape chosen for this M nor on the
          // order in which the tiles of a row finish
          long long st_sum = 0, st_sq = 0;
          const bool emit = ep.stats_out != nullptr;
          const int my_row = row0 + lane;
#pragma unroll
          for (int sub = 0; sub < kSub; ++sub) {
            const int cur = sub & 1;
            const int col = col_base + sub * 16;
            tmem_ld_wait();
            if (sub + 1 < kSub) tmem_ld16_nowait(t_base + (sub + 1) * 16, v[cur ^ 1]);
            if (col >= N) continue;                                       // warp-uniform
            if (lane == 0) {
              tma_store_wait_read();                                       // the other slab's last store has left smem
              if (has_res && sub + 1 < kSub && col + 16 < N) {
                mbar_expect_tx(rbar[cur ^ 1], 2048);
                tma_load_2d(slab_u32[cur ^ 1], &tma_r, col + 16, row0, rbar[cur ^ 1]);
              }
            }
            if (has_res) { mbar_wait(rbar[cur], rphase[cur]); rphase[cur] ^= 1u; }
            __syncwarp();
            if (!(ep.dbg & 1)) {
              uint32_t xpk[8];
              float g_sum = 0.f, g_sq = 0.f;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint8_t* ptr = slab_gen[cur] + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + sub * 16 + 4 * j);
                float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_res) r4 = *reinterpret_cast<const float4*>(ptr);
                float x0 = fmaf(__uint_as_float(v[cur][4 * j]), ep.scale, b4.x), x1 = fmaf(__uint_as_float(v[cur][4 * j + 1]), ep.scale, b4.y);
                float x2 = fmaf(__uint_as_float(v[cur][4 * j + 2]), ep.scale, b4.z), x3 = fmaf(__uint_as_float(v[cur][4 * j + 3]), ep.scale, b4.w);
                if (ACT == 1) { x0 = gelu_erf(x0); x1 = gelu_erf(x1); x2 = gelu_erf(x2); x3 = gelu_erf(x3); }
                else if (ACT == 2) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
                const float y0 = x0 + r4.x, y1 = x1 + r4.y, y2 = x2 + r4.z, y3 = x3 + r4.w;
                *reinterpret_cast<float4*>(ptr) = make_float4(y0, y1, y2, y3);
                if (emit) {
                  g_sum += (y0 + y1) + (y2 + y3);
                  g_sq = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, g_sq))));
                  if (ep.fp16) {
                    __half2 a2 = __floats2half2_rn(y0, y1), b2 = __floats2half2_rn(y2, y3);
                    xpk[2 * j] = *reinterpret_cast<uint32_t*>(&a2); xpk[2 * j + 1] = *reinterpret_cast<uint32_t*>(&b2);
                  } else {
                    __nv_bfloat162 a2 = __floats2bfloat162_rn(y0, y1), b2 = __floats2bfloat162_rn(y2, y3);
                    xpk[2 * j] = *reinterpret_cast<uint32_t*>(&a2); xpk[2 * j + 1] = *reinterpret_cast<uint32_t*>(&b2);
                  }
                }
              }
              if (emit) { st_sum += __float2ll_rn(g_sum * 16777216.0f); st_sq += __float2ll_rn(g_sq * 65536.0f); }
              if (emit && my_row < M && col + 15 < N) {     // the raw row rounded to the operand type (the next GEMM's A): one 32-byte sector
                uint4* xdst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(ep.x16_out) + (long)my_row * ep.ldx16 + col);
                xdst[0] = make_uint4(xpk[0], xpk[1], xpk[2], xpk[3]);
                xdst[1] = make_uint4(xpk[4], xpk[5], xpk[6], xpk[7]);
              }
              fence_async_smem();
              __syncwarp();
              if (lane == 0) tma_store_2d(&tma_c, slab_u32[cur], col, row0);
            }
          }
          if (emit && my_row < M && col_base < N) {
            unsigned long long* so = reinterpret_cast<unsigned long long*>(ep.stats_out) + 2 * (long)my_row;
            atomicAdd(so, (unsigned long long)st_sum);
            atomicAdd(so + 1, (unsigned long long)st_sq);
          }
          // The two statistics buffers of a Swin block alternate (proj -> fc1 uses one, fc2 -> next qkv the other); the
          // GEMM that runs after a buffer's consumer clears it for its next producer, so no memset node sits between the
          // kernels of a block (a memset also breaks the programmatic dependent launch chain).  First column tile only.
          if (ep.stats_zero && col_base == 0 && my_row < M)
            *reinterpret_cast<ulonglong2*>(reinterpret_cast<unsigned long long*>(ep.stats_zero) + 2 * (long)my_row) = make_ulonglong2(0ull, 0ull);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CTAS == 2) mbar_arrive_cluster(map_to_cta(tempty_bar(buf), 0u)); else mbar_arrive(tempty_bar(buf)); }
      }
      if (active && lane == 0) tma_store_wait_all();
    } else {
    // Transpose 16-column accumulator slices through smem so that global traffic is coalesced row segments
    // (direct lane==row stores were measured 30% slower: 32 sectors per store instruction saturate the LSU).
    float* tile_s = staging_gen + (warp - 2) * (32 * 16);
    const int sub_r = lane >> 2, c4 = (lane & 3) * 4;   // coalesced phase: 8 rows x 4 lanes x 4 columns per instruction
    constexpr int kColsPerWarp = BN / (kEpiWarps / 4);
    constexpr int kSteps = kColsPerWarp / 16;      // 16-column steps per warp
    const bool ld_vec = ((ep.ldc & 3) == 0) && (!ep.res || (ep.ldr & 3) == 0);
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
      const int buf = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int bz = ep.batch > 0 ? tile / tiles_pb : 0, tl = tile - bz * tiles_pb;
      const int m0 = (tl / n_tiles) * kTileM + row_off, n0 = (tl % n_tiles) * BN;
      const int row_base = m0 + q * 32;
      const int col_base = n0 + hf * kColsPerWarp;
      const float* res_b = ep.res ? ep.res + (long)bz * ep.sR : nullptr;
      // this lane's bias values for all steps of the tile, fetched before waiting for the accumulator (the L1 is
      // carved out for smem, so an in-loop bias load costs an exposed L2 round trip per step: measured 2.5k cycles/step)
      float4 bcol[kSteps];
#pragma unroll
      for (int sidx = 0; sidx < kSteps; ++sidx) {
        const int col = col_base + sidx * 16 + c4;
        bcol[sidx] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.bias && col < N) {
          if (col + 3 < N) bcol[sidx] = *reinterpret_cast<const float4*>(ep.bias + col);
          else {
            bcol[sidx].x = ep.bias[col];
            if (col + 1 < N) bcol[sidx].y = ep.bias[col + 1];
            if (col + 2 < N) bcol[sidx].z = ep.bias[col + 2];
          }
        }
      }
      float4 rres[2][4];
      auto load_res = [&](int step, float4 (&r)[4]) {
        const int col = col_base + step * 16 + c4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int row = row_base + i * 8 + sub_r;
          if (res_b && row < M && col < N) {
            const float* src = res_b + (long)row * ep.ldr + col;
            if (ld_vec && col + 3 < N) r[i] = *reinterpret_cast<const float4*>(src);
            else {
              r[i].x = src[0];
              if (col + 1 < N) r[i].y = src[1];
              if (col + 2 < N) r[i].z = src[2];
              if (col + 3 < N) r[i].w = src[3];
            }
          }
        }
      };
      load_res(0, rres[0]);                        // the residual does not depend on the accumulator either
      mbar_wait(tfull_bar(buf), acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + buf * Cfg::kAccCols + hf * kColsPerWarp;
      uint32_t v[2][16];
      tmem_ld16_nowait(t_base, v[0]);
#pragma unroll
      for (int sidx = 0; sidx < kSteps; ++sidx) {
        const int cur = sidx & 1;
        tmem_ld_wait();
        // lane == accumulator row.  Row stride 16 floats, 16-byte slot j stored at j ^ ((row >> 1) & 3): both the
        // row-wise stores and the 8-rows-x-4-slots loads below are bank-conflict free.
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(&tile_s[lane * 16 + 4 * (j ^ ((lane >> 1) & 3))]) =
              make_uint4(v[cur][4 * j], v[cur][4 * j + 1], v[cur][4 * j + 2], v[cur][4 * j + 3]);
        __syncwarp();
        if (sidx + 1 < kSteps) {                   // next step's accumulator slice and residual are in flight during this step's math
          tmem_ld16_nowait(t_base + (sidx + 1) * 16, v[cur ^ 1]);
          load_res(sidx + 1, rres[cur ^ 1]);
        }
        const int col = col_base + sidx * 16 + c4;
        if (!(ep.dbg & 1) && col < N) {
          const bool vec_ok = ld_vec && (col + 3 < N);
          const float bb[4] = {bcol[sidx].x, bcol[sidx].y, bcol[sidx].z, bcol[sidx].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rl = i * 8 + sub_r;
            const int row = row_base + rl;
            const float4 a = *reinterpret_cast<const float4*>(&tile_s[rl * 16 + 4 * ((lane & 3) ^ ((rl >> 1) & 3))]);
            float x[4] = {a.x, a.y, a.z, a.w};
            const float rr[4] = {rres[cur][i].x, rres[cur][i].y, rres[cur][i].z, rres[cur][i].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float t = fmaf(x[e], ep.scale, bb[e]);
              if (ACT == 1) t = (OUT == 1) ? gelu_fast(t) : gelu_erf(t);
              else if (ACT == 2) t = fmaxf(t, 0.f);
              x[e] = t + rr[e];
            }
            if ((ep.dbg & 8) && x[0] != 1234567.f) continue;     // timing experiment: all the math, no global stores
            if (row < M) {
              if (OUT == 0) {
                float* dst = ep.Cf + (long)bz * ep.sC + (long)row * ep.ldc + col;
                if (vec_ok) *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
                else for (int e = 0; e < 4; ++e) if (col + e < N) dst[e] = x[e];
              } else {
                uint32_t lo, hi;
                if (ep.fp16) {
                  __half2 a2 = __floats2half2_rn(x[0], x[1]), b2 = __floats2half2_rn(x[2], x[3]);
                  lo = *reinterpret_cast<uint32_t*>(&a2); hi = *reinterpret_cast<uint32_t*>(&b2);
                } else {
                  __nv_bfloat162 a2 = __floats2bfloat162_rn(x[0], x[1]), b2 = __floats2bfloat162_rn(x[2], x[3]);
                  lo = *reinterpret_cast<uint32_t*>(&a2); hi = *reinterpret_cast<uint32_t*>(&b2);
                }
                uint16_t* dst = reinterpret_cast<uint16_t*>(ep.Cb) + (long)bz * ep.sC + (long)row * ep.ldc + col;
                if (vec_ok) *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
                else {
                  const uint16_t h4[4] = {(uint16_t)(lo & 0xffffu), (uint16_t)(lo >> 16), (uint16_t)(hi & 0xffffu), (uint16_t)(hi >> 16)};
                  for (int e = 0; e < 4; ++e) if (col + e < N) dst[e] = h4[e];
                }
              }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      if (lane == 0) { if (CTAS == 2) mbar_arrive_cluster(map_to_cta(tempty_bar(buf), 0u)); else mbar_arrive(tempty_bar(buf)); }
    }
    }
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();      // neither CTA frees TMEM or exits while the pair still uses it
  else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static bool make_map(CUtensorMap* map, const void* ptr, long rows, long cols, long ld, int box_rows, int fp16,
                     int box_cols = kBK, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// (k, row, problem) view of `batch` stacked K-major matrices: rows x cols each, row pitch ld, problem pitch bs (elements)
static bool make_map3(CUtensorMap* map, const void* ptr, long rows, long cols, long ld, long bs, long batch, int box_rows, int fp16) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bs * 2};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int g_tc_debug = 0;
int g_pdl_enabled = 1;
void set_tc_debug(int v) { g_tc_debug = v; }

bool tc_gemm_supported(int M, int N, int K) { return M > 0 && N > 0 && K > 0 && (K % 8) == 0; }
// LayerNorm-on-load: K = 512 (8 resident k-blocks), 128-wide tiles, one tile per CTA
bool tc_gemm_ln_supported(int M, int N, int K) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return M > 0 && N > 0 && K == 512 && (long)((M + kBM - 1) / kBM) * ((N + 127) / 128) <= sms;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

extern int g_tc_debug;

static bool make_map32(CUtensorMap* map, const float* ptr, long rows, long cols, long ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {16u, 32u};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int g_tc_pair = 1;          // CTA-pair (cta_group::2) tiles for problems with at least a wave of 256-row tiles
void set_tc_pair(int v) { g_tc_pair = v; }

template <int BN, int OUT, int ACT, int CTAS>
static cudaError_t launch_tc(const TcGemmArgs& p, cudaStream_t st) {
  using Cfg = TcCfg<BN, CTAS>;
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(gemm_tc_kernel<BN, OUT, ACT, CTAS>, Cfg::kSmemBytes, smem_state)) return e;
  CUtensorMap ma, mb;
  if (p.batch > 0) {
    if (CTAS != 1 || p.a32 || p.stats_out || p.ln_stats || p.stats_zero) return cudaErrorInvalidValue;
    if (!make_map3(&ma, p.A, p.M, p.K, p.lda, p.sA, p.batch, kBM, p.fp16) || !make_map3(&mb, p.W, p.N, p.K, p.ldw, p.sW, p.batch, BN, p.fp16))
      return cudaErrorInvalidValue;
  } else if (!make_map(&ma, p.A, p.M, p.K, p.lda, kBM, p.fp16) || !make_map(&mb, p.W, p.N, p.K, p.ldw, BN / CTAS, p.fp16)) return cudaErrorInvalidValue;
  // 16-bit outputs without a residual leave through TMA stores: 32 x 32 boxes, 64-byte swizzle
  bool tma_c_ok = false;
  CUtensorMap mc = ma, mr = ma;
  if (OUT == 1) {
    tma_c_ok = p.batch <= 0 && !p.res && (p.ldc % 8) == 0 && (reinterpret_cast<uintptr_t>(p.Cb) & 15) == 0 && !(g_tc_debug & 16);
    if (tma_c_ok && !make_map(&mc, p.Cb, p.M, p.N, p.ldc, 32, p.fp16, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return cudaErrorInvalidValue;
  } else {
    // fp32 output (+ fp32 residual): 32-row x 16-column boxes (64 bytes per row), 64-byte swizzle
    tma_c_ok = p.batch <= 0 && (p.ldc % 4) == 0 && (reinterpret_cast<uintptr_t>(p.Cf) & 15) == 0 && !(g_tc_debug & 32) &&
               (!p.res || ((p.ldr % 4) == 0 && (reinterpret_cast<uintptr_t>(p.res) & 15) == 0));
    if (tma_c_ok) {
      if (!make_map32(&mc, p.Cf, p.M, p.N, p.ldc)) return cudaErrorInvalidValue;
      if (p.res && !make_map32(&mr, p.res, p.M, p.N, p.ldr)) return cudaErrorInvalidValue;
    }
  }
  // the folded-LayerNorm hooks live in the two TMA-store epilogues only
  if ((p.stats_out || p.ln_stats || p.stats_zero) && !tma_c_ok) return cudaErrorInvalidValue;
  if (p.stats_zero && OUT != 0) return cudaErrorInvalidValue;
  if (p.stats_out && (OUT != 0 || !p.x16_out || (p.ldx16 & 7) || (p.N & 15) || (reinterpret_cast<uintptr_t>(p.x16_out) & 15))) return cudaErrorInvalidValue;
  if (p.ln_stats && (OUT != 1 || p.ln_k <= 0 || p.div != 0.f)) return cudaErrorInvalidValue;
  TcEpilogue ep{p.Cf, p.Cb, p.ldc, p.fp16, p.bias, p.res, p.ldr, p.div != 0.f ? 1.0f / p.div : 1.0f, p.w_static, p.a32, p.lda32, p.ln_g, p.ln_b,
                p.stats_out, p.x16_out, p.ldx16, p.stats_zero, p.ln_stats, p.ln_k > 0 ? 1.0f / (float)p.ln_k : 0.f,
                tma_c_ok ? 1 : 0, g_tc_debug, p.batch > 0 ? p.batch : 0, p.sC, p.sR};
  const int tiles = ((p.M + kBM * CTAS - 1) / (kBM * CTAS)) * ((p.N + BN - 1) / BN) * (p.batch > 0 ? p.batch : 1);
  const int slots = sm_count() / CTAS;
  const int grid = CTAS * (tiles < slots ? tiles : slots);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CTAS == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_pdl_enabled) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, OUT, ACT, CTAS>, ma, mb, mc, mr, ep, p.M, p.N, p.K);
}

template <int BN, int CTAS>
static cudaError_t launch_tc_bn(const TcGemmArgs& p, cudaStream_t st) {
  const int out = p.Cf ? 0 : 1;
  switch (out * 3 + p.act) {
    case 0: return launch_tc<BN, 0, 0, CTAS>(p, st);
    case 1: return launch_tc<BN, 0, 1, CTAS>(p, st);
    case 2: return launch_tc<BN, 0, 2, CTAS>(p, st);
    case 3: return launch_tc<BN, 1, 0, CTAS>(p, st);
    case 4: return launch_tc<BN, 1, 1, CTAS>(p, st);
    case 5: return launch_tc<BN, 1, 2, CTAS>(p, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_gemm_tc(const TcGemmArgs& p, cudaStream_t st) {
  if (p.a32) {          // LayerNorm-on-load: fixed configuration (128-wide single-CTA tiles, one per CTA)
    if (!tc_gemm_ln_supported(p.M, p.N, p.K) || (p.ldw % 8) || (p.lda32 % 4) || ((p.Cf != nullptr) == (p.Cb != nullptr)) ||
        (reinterpret_cast<uintptr_t>(p.a32) & 15) || (reinterpret_cast<uintptr_t>(p.W) & 15) || p.act < 0 || p.act > 2)
      return cudaErrorInvalidValue;
    TcGemmArgs q = p;
    q.A = p.W; q.lda = p.ldw;                      // a valid tensor map is still built for the unused A operand
    return launch_tc_bn<128, 1>(q, st);
  }
  if (!tc_gemm_supported(p.M, p.N, p.K) || (p.lda % 8) || (p.ldw % 8) || ((p.Cf != nullptr) == (p.Cb != nullptr)))
    return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.W) & 15)) return cudaErrorInvalidValue;
  if (p.batch > 0) {
    // per-image contractions of the static-expansion block: independent problems of one shape, operands stacked with the
    // pitches sA / sW.  Tile width with the least padding (ties to the wider tile); single-CTA tiles, generic epilogue.
    if ((p.sA % 8) || (p.sW % 8) || p.act < 0 || p.act > 2) return cudaErrorInvalidValue;
    int bn = 256;
    long waste = -1;
    for (int c : {256, 192, 128}) {
      const long w = (long)((p.N + c - 1) / c) * c - p.N;
      if (waste < 0 || w < waste) { bn = c; waste = w; }
    }
    // few tiles in total (split-K slices of a decoder-step linear): 64-wide tiles, more CTAs sharing the operand stream
    if ((p.N % 64) == 0 && (long)((p.M + kBM - 1) / kBM) * (p.N / 64) * p.batch <= sm_count()) return launch_tc_bn<64, 1>(p, st);
    if (bn == 256) return launch_tc_bn<256, 1>(p, st);
    if (bn == 192) return launch_tc_bn<192, 1>(p, st);
    return launch_tc_bn<128, 1>(p, st);
  }
  // pick the tile width with the least padded work; ties go to the wider tile.  With only a few waves of tiles over the
  // persistent grid the wave count decides instead (a nearly empty second wave doubles the time): cost = waves x width.
  const int cands[3] = {256, 192, 128};
  int best = 256;
  long best_cost = -1;
  const long m_tiles = (p.M + kBM - 1) / kBM;
  const bool few = m_tiles * ((p.N + 255) / 256) < 2L * sm_count();
  for (int bn : cands) {
    const long n_tiles = (p.N + bn - 1) / bn;
    const long cost = few ? ((m_tiles * n_tiles + sm_count() - 1) / sm_count()) * bn : n_tiles * bn - p.N;
    if (best_cost < 0 || cost < best_cost) { best = bn; best_cost = cost; }
  }
  if (p.act < 0 || p.act > 2) return cudaErrorInvalidValue;
  // skinny problems (decoder steps, M = images x beam): few tiles, so prefer narrow tiles -- more CTAs in flight and
  // a shorter serial epilogue per CTA
  const long tiles256 = (long)((p.M + kBM - 1) / kBM) * ((p.N + 255) / 256);
  if (tiles256 * 2 <= sm_count()) best = 128;
  // CTA pairs where the main loop is the limit: at least one full wave of 256-row pair tiles and a long contraction.
  // Measured (profiles/): K = 3072 +8 %, K = 768 with 16-bit output +8 % (+6 % with the GELU epilogue once that became
  // cheap), 8192^3 +10 %; short-K / epilogue-bound shapes (stage 1-2, fp32+residual epilogues at K <= 768) lose 5-40 % to
  // the pair's coupled epilogues, so they stay on single-CTA tiles.
  const long pair_tiles = (long)((p.M + 2 * kBM - 1) / (2 * kBM)) * ((p.N + best - 1) / best);
  const bool long_k = p.K >= 1536 || (p.K >= 768 && p.Cb != nullptr);
  if (g_tc_pair && long_k && pair_tiles >= sm_count() / 2) {
    if (best == 256) return launch_tc_bn<256, 2>(p, st);
    if (best == 192) return launch_tc_bn<192, 2>(p, st);
    return launch_tc_bn<128, 2>(p, st);
  }
  if (best == 256) return launch_tc_bn<256, 1>(p, st);
  if (best == 192) return launch_tc_bn<192, 1>(p, st);
  // very few tiles (decoder-step projections onto d_model): 64-wide tiles double the CTAs that share the operand stream
  if ((long)((p.M + kBM - 1) / kBM) * ((p.N + 127) / 128) * 8 <= sm_count()) return launch_tc_bn<64, 1>(p, st);
  return launch_tc_bn<128, 1>(p, st);
}

}  // namespace xn
