// One decoder position for all rows in ONE persistent kernel (16-bit modes).
//
// The one-kernel-per-operation path (engine.cu: dec_step_t) runs ~33 dependent launches of a few microseconds per
// decoded position; at 64 images x beam 3 the 19 positions of a caption cost 5.2 ms for 1.2 % of the call's FLOPs.
// Here the whole position -- embedding, N_dec x (norm_1 + fused [cond|key|A|B|selector] projection, incremental dynamic
// expansion + norm_2, W_q, cross attention, W_o + residual, norm_3 + FF1 + ReLU, FF2 + residual), reduce group, final
// norm + vocabulary projection, and optionally log-softmax + top-k -- is a sequence of PHASES executed by one grid of
// co-resident CTAs (two per SM) with a grid barrier between dependent phases.  The decoder weights (34 MB in 16 bits)
// stay in the 126 MB L2 from one position to the next; activations move between phases through L2.
//
// Reference: models/End_ExpansionNet_v2.py:155-209 (forward_dec), models/layers.py:152-204 (DynamicExpansionBlock),
// :207-262 (DecoderLayer), :266-295 (MultiHeadAttention), models/captioning_model.py:302-317 (log-softmax + top-k).
//
// GEMM phases: C[M,N] = act(A[M,K] W[N,K]^T + b) + res on mma.sync m16n8k16 (M is 3..1536 rows: a 128-row tcgen05 tile
// would be mostly padding at the common sizes, and the phases are bound by L2 latency, not by the tensor pipe).  Tiles
// are 64 x {32,64}, K streamed in 64-column chunks through a 4-stage cp.async ring with a 128-byte XOR swizzle; where the
// operand is LayerNorm(x) of an fp32 row (K = 512) the CTA normalises its 64 rows straight into shared memory.
#include <algorithm>
#include "kernels.h"
#include "common.cuh"
#include "mma16_frag.cuh"
#include "tcgen05_ptx.cuh"
#include "decode_rows.cuh"
#include "beam_rows.cuh"

namespace xn {

int g_mega_coop = 1;

constexpr int kMegaThreads = 256;
constexpr int kBM = 32, kKI = 512;              // rows per tile; K per work item (longer K is split over CTAs)
constexpr int kPadK = kKI + 8;                  // elements per packed operand row
constexpr int kPitch = kPadK * 2;           // 1040 bytes per operand row: consecutive rows shift by 16 bytes -> ldmatrix is conflict-free
constexpr int kARegion = kBM * kPitch;          // 33 280 bytes
constexpr int kWRegion = 64 * kPitch;           // 66 560 bytes (BN <= 64)
constexpr int kMbarOff = kARegion + kWRegion;   // one mbarrier behind the operand regions
constexpr int kGemmSmem = kMbarOff + 64;
constexpr int kBarStride = 32;                  // unsigned per barrier word: one 128-byte line each
constexpr int kSplitCntOff = 2 * kBarStride;    // split-K tile counters follow the two barrier words
constexpr int kMaxSplitTiles = 4096;

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ---- grid barrier.  Arrivals are release atomic adds on one word (cooperative groups' trick: CTA 0 adds 2^31 - (n-1),
// the others 1, so bit 31 flips exactly on the last arrival and the word needs no reset between barriers or launches).
// The acquire side is one fence.acq_rel.gpu after the flip has been seen: it orders the data and invalidates the SM's L1
// (CCTL.IVALL), so plain loads after the barrier see what other CTAs wrote before it.  Measured alternatives (a separate
// flag line published by the last arriver; two-level arrival counters; __threadfence() = MEMBAR.SC on both sides) were
// all slower: the cost is the chain store-acknowledge -> atomic round trip -> poll round trip, not contention.
struct GridBar {
  unsigned* cnt;
  __device__ __forceinline__ void init(unsigned* b) { cnt = b; }
  __device__ __forceinline__ void sync(unsigned long long* dbg = nullptr, int* mark = nullptr) {
    __syncthreads();
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[(*mark)++] = gtimer();
    if (threadIdx.x == 0) {
      const unsigned inc = (blockIdx.x == 0) ? (0x80000000u - (gridDim.x - 1)) : 1u;
      unsigned old, v;
      asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(cnt), "r"(inc) : "memory");
      // everybody polls the arrival word itself (bit 31 flips on the last arrival) with RELAXED loads -- an acquire load
      // would carry an L1 invalidation per iteration -- and fences once afterwards
      while (true) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
        if ((v ^ old) & 0x80000000u) break;
      }
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[(*mark)++] = gtimer();
  }
};

struct GemmP {
  const void* A16;                                            // 16-bit A, slab-packed [K/512][M32][520] (see mega_pack), or
  const float* A32; long lda32; const float *ln_g, *ln_b;     // fp32 rows normalised on load (K == 512)
  const void* W; const float* bias;                           // 16-bit W, slab-packed [K/512][N][520]
  const float* res; long ldr;                                 // fp32 residual added after the activation (may alias Cf)
  float* Cf; long ldcf;                                       // fp32 output and / or
  void* Cb;                                                   // 16-bit output, slab-packed [N/512][M32][520]
  int M, N, K, act;                                           // act: 0 none, 2 ReLU;  K = 512 * ksplit
  int ksplit; float* scratch; unsigned* cnt;                  // split-K over CTAs: partial tiles + per-tile arrival counters
  // first decoder layer: the fp32 rows are the embedding itself, x0 = E[tok] sqrt(d) + P[p] (layers.py:16-17), computed in
  // the LayerNorm fill; the column-tile-0 items also store x0 (the residual input of the dynamic expansion)
  const float *emb, *pos_row; const int64_t* tok64; const int* tok32; long tok_stride; int tok_p; float* x0_out;
  int dbg_mode;
  unsigned long long* fine; int* fmark;                      // optional intra-phase timestamps (CTA 0, thread 0)
};

template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<f16>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<bf16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// per-tile hook of the vocabulary projection (fused log-softmax / top-k partials): sees the finished 32 x BN tile in
// shared memory (fp32, row pitch BN + 8)
struct NoTileHook {
  __device__ __forceinline__ void operator()(int, int, int, const float*) const {}
};

#ifdef XN_MEGA_FINE      // intra-phase timestamps (build with -DXN_MEGA_FINE; tools/mega_timeline.py prints them)
#define XN_FINE(g) do { if ((g).fine && blockIdx.x == 0 && threadIdx.x == 0 && *(g).fmark < 124) (g).fine[(*(g).fmark)++] = gtimer(); } while (0)
#else
#define XN_FINE(g) do { } while (0)
#endif

// One GEMM phase.  A work item is a 32 x BN output tile over K = 512 (longer K: ksplit items per tile).  Both operands
// of an item are brought in whole -- one 1 KB bulk copy (cp.async.bulk) per operand row, completing on an mbarrier: no
// per-thread copy instructions, no staging ring -- or, for a LayerNorm operand, normalised from the fp32 rows straight into
// shared memory.  The eight warps each take a 64-wide slice of K with a 32 x BN accumulator (8 or 16 independent MMAs per
// k-step, no barrier inside the item); the eight partial tiles meet in shared memory, where bias / activation / residual
// run with coalesced 16-byte accesses.  With ksplit > 1 the partial tile goes to `scratch` and the CTA that arrives last
// at the tile's counter adds the splits in order (deterministic) and finishes.
template <typename T, int BN, bool LN, bool STORE = true, typename Hook = NoTileHook>
__device__ __forceinline__ void gemm_phase_impl(const GemmP& g, char* smem, uint32_t& mphase, uint32_t& mphase2, bool& w_ready,
                                                const Hook& hook = Hook()) {
  constexpr int NB2 = BN / 16, NF = BN / 8, PITCH = BN + 8, C4 = BN / 4;
  static_assert(4 * kBM * PITCH * 4 <= kWRegion, "reduction tiles must fit in the W region");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = g.ksplit > 1 ? g.ksplit : 1;
  const int ntm = (g.M + kBM - 1) / kBM, ntn = (g.N + BN - 1) / BN, ntiles = ntm * ntn;
  const long m32 = (long)ntm * kBM;                               // rows of a packed activation slab
  char* a_sm = smem;
  char* w_sm = smem + kARegion;
  float* red = reinterpret_cast<float*>(smem + kARegion);        // [4][32][PITCH] after the MMAs (the A region survives: A-stationary tiles)
  const uint32_t mbar = smem_u32(smem + kMbarOff);
  __shared__ int s_last;
  const T* W = reinterpret_cast<const T*>(g.W);
  const T* A16 = reinterpret_cast<const T*>(g.A16);
  // A-stationary order for the LayerNorm operand: a CTA keeps one 32-row block and walks over column tiles, so the rows
  // are normalised once per CTA instead of once per item (the vocabulary projection has 157 column tiles)
  const bool astat = LN && S == 1 && (int)gridDim.x >= ntm;
  const int cpg = astat ? (int)gridDim.x / ntm : 1;
  const int my_tm = blockIdx.x % ntm, my_slot = blockIdx.x / ntm;
  const int it_end = astat ? ntn : ntiles * S, it_step = astat ? cpg : (int)gridDim.x;
  bool a_ready = false;
  for (int it = astat ? (my_slot < cpg ? my_slot : ntn) : (int)blockIdx.x; it < it_end; it += it_step) {
    const int item = astat ? it * ntm + my_tm : it;
    const int tile = item / S, sp = item - tile * S;
    const int tm = tile % ntm, tn = tile / ntm, r0 = tm * kBM, n0 = tn * BN;
    XN_FINE(g);                                          // 0: item start
    const bool do_load = !(g.dbg_mode & 2);
    if (do_load && tid == 0) {
      // Operands are stored with the 520-element row pitch of the shared-memory tile (weights re-packed at load time,
      // activations written that way by the producing phase), one K slab of 512 after the other: a whole operand tile is
      // ONE contiguous bulk copy.  The regions were last touched through the generic proxy (ldmatrix / reduction) by this
      // CTA, all before the __syncthreads that ended the previous item.  W and A complete on separate mbarriers: the W
      // tile of a phase's first item is usually already in flight (prefetch_w, issued before the preceding grid barrier).
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (!w_ready) {
        const int wrows = min(BN, g.N - n0);
        mbar_expect(mbar, (uint32_t)(wrows * kPitch));
        bulk_g2s(smem_u32(w_sm), W + ((long)sp * g.N + n0) * kPadK, (uint32_t)(wrows * kPitch), mbar);
      }
      if (!LN) {
        mbar_expect(mbar + 8, (uint32_t)(kBM * kPitch));
        bulk_g2s(smem_u32(a_sm), A16 + ((long)sp * m32 + r0) * kPadK, (uint32_t)(kBM * kPitch), mbar + 8);
      }
    }
    w_ready = false;
    if (LN && !a_ready && !(g.dbg_mode & 4)) {
      // LayerNorm(gamma, beta) of rows r0 .. r0+31 (K == 512): four rows per warp, all their loads in flight together; same
      // arithmetic and summation order as layernorm_kernel (elementwise.cu: ln_row)
      float4 v[4][4];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int grow = r0 + warp * 4 + rr, crow_ = min(grow, g.M - 1);
        if (g.emb) {
          const long tok = g.tok64 ? (long)g.tok64[crow_ * g.tok_stride + g.tok_p] : (long)g.tok32[crow_ * g.tok_stride + g.tok_p];
          const float sc = sqrtf(512.0f);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 e4 = *reinterpret_cast<const float4*>(g.emb + tok * 512 + j * 128 + lane * 4);
            const float4 p4 = *reinterpret_cast<const float4*>(g.pos_row + j * 128 + lane * 4);
            v[rr][j] = make_float4(e4.x * sc + p4.x, e4.y * sc + p4.y, e4.z * sc + p4.z, e4.w * sc + p4.w);
            if (tn == 0 && grow < g.M) *reinterpret_cast<float4*>(g.x0_out + (long)grow * 512 + j * 128 + lane * 4) = v[rr][j];
          }
        } else {
          const float* xr = g.A32 + (long)crow_ * g.lda32;
#pragma unroll
          for (int j = 0; j < 4; ++j) v[rr][j] = __ldcg(reinterpret_cast<const float4*>(xr + j * 128 + lane * 4));
        }
      }
      float4 gg[4], bb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        gg[j] = *reinterpret_cast<const float4*>(g.ln_g + j * 128 + lane * 4);
        bb[j] = *reinterpret_cast<const float4*>(g.ln_b + j * 128 + lane * 4);
      }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int row = warp * 4 + rr;
        float sm_ = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) sm_ += (v[rr][j].x + v[rr][j].y) + (v[rr][j].z + v[rr][j].w);
        const float mean = warp_sum(sm_) / 512.0f;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d0 = v[rr][j].x - mean, d1 = v[rr][j].y - mean, d2 = v[rr][j].z - mean, d3 = v[rr][j].w - mean;
          q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / 512.0f + 1e-5f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint2 u;
          u.x = pack2<T>((v[rr][j].x - mean) * rstd * gg[j].x + bb[j].x, (v[rr][j].y - mean) * rstd * gg[j].y + bb[j].y);
          u.y = pack2<T>((v[rr][j].z - mean) * rstd * gg[j].z + bb[j].z, (v[rr][j].w - mean) * rstd * gg[j].w + bb[j].w);
          *reinterpret_cast<uint2*>(a_sm + row * kPitch + (j * 128 + lane * 4) * 2) = u;
        }
      }
    }
    // bias / residual of this thread's outputs: in flight while the operands land
    float4 pb[C4 * kBM / kMegaThreads], pr[C4 * kBM / kMegaThreads];
#pragma unroll
    for (int u = 0; u < C4 * kBM / kMegaThreads; ++u) {
      const int idx = tid + u * kMegaThreads, row = idx / C4, c4 = idx % C4, grow = r0 + row, col = n0 + c4 * 4;
      pb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      pr[u] = pb[u];
      if (col < g.N) {
        if (g.bias) pb[u] = *reinterpret_cast<const float4*>(g.bias + col);
        if (STORE && g.res && grow < g.M && (S == 1 || true)) pr[u] = __ldcg(reinterpret_cast<const float4*>(g.res + (long)grow * g.ldr + col));
      }
    }
    XN_FINE(g);                                          // 1: copies issued, LayerNorm fill done
    if (do_load) {
      tc5::mbar_wait(mbar, mphase);
      mphase ^= 1u;
      if (!LN) {
        tc5::mbar_wait(mbar + 8, mphase2);
        mphase2 ^= 1u;
      }
    }
    if (LN && !a_ready) __syncthreads();                 // the normalised rows of all warps
    a_ready = astat;
    XN_FINE(g);                                          // 2: operands landed
    float acc[2][NF][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < NF; ++j) { acc[i][j][0] = 0.f; acc[i][j][1] = 0.f; acc[i][j][2] = 0.f; acc[i][j][3] = 0.f; }
    {
      const uint32_t a_s = smem_u32(a_sm) + warp * 128, w_s = smem_u32(w_sm) + warp * 128;       // this warp's 64 columns of K
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t af[2][4];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
          ldsm_x4(af[mb], a_s + (mb * 16 + (lane & 15)) * kPitch + ks * 32 + (lane >> 4) * 16);
#pragma unroll
        for (int nb = 0; nb < NB2; ++nb) {
          uint32_t bf_[4];
          ldsm_x4(bf_, w_s + (nb * 16 + (lane & 7) + ((lane >> 4) << 3)) * kPitch + ks * 32 + ((lane >> 3) & 1) * 16);
          if (!(g.dbg_mode & 1)) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
              Mma16<T>::mma(acc[mb][nb * 2], af[mb], bf_[0], bf_[1]);
              Mma16<T>::mma(acc[mb][nb * 2 + 1], af[mb], bf_[2], bf_[3]);
            }
          }
        }
      }
    }
    __syncthreads();                                     // every warp is done with the operands: they become the reduction tiles
    XN_FINE(g);                                          // 3: MMAs done
    // eight partial tiles -> four (warps 4-7 hand theirs to warps 0-3 in fragment layout) -> summed per output below
    auto frag_io = [&](float* rq, bool add) {
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
          const int row = mb * 16 + (lane >> 2), col = nf * 8 + (lane & 3) * 2;
          float2* p0 = reinterpret_cast<float2*>(rq + row * PITCH + col);
          float2* p1 = reinterpret_cast<float2*>(rq + (row + 8) * PITCH + col);
          if (add) {
            const float2 u0 = *p0, u1 = *p1;
            acc[mb][nf][0] += u0.x; acc[mb][nf][1] += u0.y; acc[mb][nf][2] += u1.x; acc[mb][nf][3] += u1.y;
          }
          if (!add || warp < 4) {
            *p0 = make_float2(acc[mb][nf][0], acc[mb][nf][1]);
            *p1 = make_float2(acc[mb][nf][2], acc[mb][nf][3]);
          }
        }
    };
    if (warp >= 4) frag_io(red + (size_t)(warp - 4) * kBM * PITCH, false);
    __syncthreads();
    if (warp < 4) frag_io(red + (size_t)warp * kBM * PITCH, true);          // same lanes read and rewrite the same words
    __syncthreads();
    auto sum8 = [&](int row, int c4) {
      float4 v = *reinterpret_cast<const float4*>(red + row * PITCH + c4 * 4);
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        const float4 u = *reinterpret_cast<const float4*>(red + (size_t)q * kBM * PITCH + row * PITCH + c4 * 4);
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
      }
      return v;
    };
    bool finish = true;
    if (S > 1) {
      // partial tile of this K split -> scratch; the last arriver of the tile sums the splits in order
      float* my = g.scratch + ((size_t)tile * S + sp) * (kBM * BN);
#pragma unroll
      for (int idx = tid; idx < kBM * C4; idx += kMegaThreads) {
        const int row = idx / C4, c4 = idx % C4;
        __stcg(reinterpret_cast<float4*>(my + row * BN + c4 * 4), sum8(row, c4));
      }
      __syncthreads();
      if (tid == 0) {
        unsigned old;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(g.cnt + tile), "r"(1u) : "memory");
        s_last = old == (unsigned)(S - 1);
        if (s_last) g.cnt[tile] = 0u;
      }
      __syncthreads();
      finish = s_last != 0;
    }
    if (finish) {
#pragma unroll
      for (int u = 0; u < C4 * kBM / kMegaThreads; ++u) {
        const int idx = tid + u * kMegaThreads, row = idx / C4, c4 = idx % C4, grow = r0 + row, col = n0 + c4 * 4;
        float4 v;
        if (S > 1) {
          const float* sc = g.scratch + (size_t)tile * S * (kBM * BN) + row * BN + c4 * 4;
          v = __ldcg(reinterpret_cast<const float4*>(sc));
          for (int q = 1; q < S; ++q) {
            const float4 u = __ldcg(reinterpret_cast<const float4*>(sc + (size_t)q * (kBM * BN)));
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
          }
        } else {
          v = sum8(row, c4);
        }
        if (col < g.N) {                                 // N is a multiple of 4
          v.x += pb[u].x; v.y += pb[u].y; v.z += pb[u].z; v.w += pb[u].w;
          if (g.act == 2) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          if (STORE && grow < g.M) {
            v.x += pr[u].x; v.y += pr[u].y; v.z += pr[u].z; v.w += pr[u].w;
            if (g.Cf) *reinterpret_cast<float4*>(g.Cf + (long)grow * g.ldcf + col) = v;
            if (g.Cb) {
              uint2 u;
              u.x = pack2<T>(v.x, v.y); u.y = pack2<T>(v.z, v.w);
              *reinterpret_cast<uint2*>(reinterpret_cast<T*>(g.Cb) + ((long)(col >> 9) * m32 + grow) * kPadK + (col & 511)) = u;
            }
          }
        } else {
          v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
        if constexpr (!STORE) {
          __syncwarp();
          *reinterpret_cast<float4*>(red + row * PITCH + c4 * 4) = v;       // finished tile, for the hook
        }
      }
      if constexpr (!STORE) {
        __syncthreads();
        hook(tn, r0, n0, red);
      }
    }
    __syncthreads();                                     // the regions are free: the next item's copies may land
    XN_FINE(g);                                          // 4: epilogue done
  }
}



// Issue the W tile of the first item this CTA will take in the NEXT GEMM phase (same item order as gemm_phase_impl).
// Weights do not depend on the grid barrier in between, so the copy flies during the barrier (and, for the output
// projection, during the whole cross-attention phase, which leaves the W region alone).  Returns whether a copy is in flight.
template <typename T, int BN, bool LN>
__device__ __forceinline__ bool prefetch_w(const void* Wv, int M, int N, int S, char* smem, int dbg_mode) {
  if (dbg_mode & (2 | 32)) return false;
  const int ntm = (M + kBM - 1) / kBM, ntn = (N + BN - 1) / BN;
  const bool astat = LN && S == 1 && (int)gridDim.x >= ntm;
  int tn, sp = 0;
  if (astat) {
    const int cpg = (int)gridDim.x / ntm, slot = blockIdx.x / ntm;
    if (slot >= cpg || slot >= ntn) return false;
    tn = slot;
  } else {
    const int item = blockIdx.x;
    if (item >= ntm * ntn * S) return false;
    const int tile = item / S;
    sp = item - tile * S;
    tn = tile / ntm;
  }
  if (threadIdx.x == 0) {
    const uint32_t mbar = smem_u32(smem + kMbarOff);
    const int n0 = tn * BN, wrows = min(BN, N - n0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect(mbar, (uint32_t)(wrows * kPitch));
    bulk_g2s(smem_u32(smem + kARegion), reinterpret_cast<const T*>(Wv) + ((long)sp * N + n0) * kPadK, (uint32_t)(wrows * kPitch), mbar);
  }
  return true;
}

// ---- fused log-softmax / top-k over the vocabulary (K6): every 64-column tile of the vocabulary projection leaves,
// per row, its maximum, sum of exp(x - max) and its best k (value, index) candidates; one warp per row then merges the
// tiles' partials: lse = M + log(sum_t s_t exp(m_t - M)), lp = (x - M) - lse.  The R x V logits are never stored.
// Partials are structure-of-arrays over (row, tile): mx | sum | v[kTopC] | i[kTopC], each R x ntn.
constexpr int kTopC = 8;
struct TopkParts {
  float* base; long n;                                            // n = R * ntn
  __device__ __forceinline__ float* mx() const { return base; }
  __device__ __forceinline__ float* sum() const { return base + n; }
  __device__ __forceinline__ float* v(int j) const { return base + (2 + j) * n; }
  __device__ __forceinline__ int* i(int j) const { return reinterpret_cast<int*>(base + (2 + kTopC + j) * n); }
};

__device__ __forceinline__ bool better_(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

template <int BN>
struct TopkHook {
  TopkParts parts; int ntn, M, N, k;
  // `tile` holds the finished logits (bias added; columns >= N are -inf), row pitch BN + 8.  Eight lanes own one row: each
  // scans BN/8 columns, then the group combines.
  __device__ __forceinline__ void operator()(int tn, int r0, int n0, const float* tile) const {
    const int row = threadIdx.x >> 3, sub = threadIdx.x & 7;      // 256 threads = 32 rows x 8 lanes
    const float* tr = tile + row * (BN + 8);
    float x[BN / 8];
#pragma unroll
    for (int j = 0; j < BN / 8; ++j) x[j] = tr[sub + 8 * j];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < BN / 8; ++j) mx = fmaxf(mx, x[j]);
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < BN / 8; ++j) sum += expf(x[j] - mx);      // exp(-inf) = 0 for the columns beyond N
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    // top-k of the row's BN values: k rounds of a group arg-max (descending, ties to the lower index)
    const long slot = (long)(r0 + row) * ntn + tn;
    const bool wr = sub == 0 && r0 + row < M;
    for (int round = 0; round < k; ++round) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < BN / 8; ++j)
        if (better_(x[j], n0 + sub + 8 * j, bv, bi)) { bv = x[j]; bi = n0 + sub + 8 * j; }
#pragma unroll
      for (int o = 1; o <= 4; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better_(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
#pragma unroll
      for (int j = 0; j < BN / 8; ++j)
        if (n0 + sub + 8 * j == bi) x[j] = -INFINITY;             // retire the winner (a retired -inf never beats index order again: see merge)
      if (wr) { parts.v(round)[slot] = bv; parts.i(round)[slot] = bi; }
    }
    if (wr) { parts.mx()[slot] = mx; parts.sum()[slot] = sum; }
  }
};

// merge of the per-tile partials: one warp per row.  Round r picks the best candidate that ranks strictly after round
// r-1's winner in the order (value descending, index ascending) -- no cursors, every lane rescans its tiles' k entries
// (L1 hits after the first round).
__device__ __forceinline__ void topk_merge_row(const TopkParts& parts, int ntn, int r, int k, float* top_val, int* top_idx, int lane) {
  constexpr int TPL = 8;                                        // tiles per lane (ntn <= 256)
  {
    const long base = (long)r * ntn;
    float m[TPL], sm_[TPL];
#pragma unroll
    for (int j = 0; j < TPL; ++j) {                             // all loads of the pass in flight together
      const int t = lane + 32 * j;
      m[j] = t < ntn ? parts.mx()[base + t] : -INFINITY;
      sm_[j] = t < ntn ? parts.sum()[base + t] : 0.f;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < TPL; ++j) mx = fmaxf(mx, m[j]);
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < TPL; ++j) sum += (lane + 32 * j < ntn) ? sm_[j] * expf(m[j] - mx) : 0.f;
    const float lse = logf(warp_sum(sum));
    float pv = INFINITY;
    int pi = -1;
    for (int round = 0; round < k; ++round) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int j = 0; j < k; ++j) {
        float cv[TPL];
        int ci[TPL];
#pragma unroll
        for (int u = 0; u < TPL; ++u) {
          const int t = lane + 32 * u;
          cv[u] = t < ntn ? parts.v(j)[base + t] : -INFINITY;
          ci[u] = t < ntn ? parts.i(j)[base + t] : 0x7fffffff;
        }
#pragma unroll
        for (int u = 0; u < TPL; ++u)
          if (ci[u] != 0x7fffffff && better_(pv, pi, cv[u], ci[u]) && better_(cv[u], ci[u], bv, bi)) { bv = cv[u]; bi = ci[u]; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better_(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
      pv = bv; pi = bi;
      if (lane == 0) {
        top_val[(long)r * k + round] = (bv - mx) - lse;
        top_idx[(long)r * k + round] = bi;
      }
    }
  }
}

__device__ __forceinline__ void topk_merge_phase(const TopkParts& parts, int ntn, int R, int k, float* top_val, int* top_idx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = blockIdx.x * (kMegaThreads / 32) + warp; r < R; r += gridDim.x * (kMegaThreads / 32))
    topk_merge_row(parts, ntn, r, k, top_val, top_idx, lane);
}

// ---- dynamic expansion of one row with the history rows STAGED in shared memory by bulk copies (np <= 20 positions,
// 16 expansion vectors, d = 512).  Same mathematics as dyn_exp_row (decode_rows.cuh), reorganised so that every round of
// dependent L2 reads becomes one wave of 2 KB bulk copies: wave 1 = cond c_i and key K_j rows (the 2p+1 new dot products
// then run from shared memory; q_e.K_p runs from global meanwhile), wave 2 = class-A rows over the keys (lands during the
// scalar phase), wave 3 = class-B rows over the cond rows once their contribution is taken.
constexpr int kDxMaxPos = 20, kDxRowB = 2048;
constexpr int kDxScal = 13312;                                   // bytes reserved for the scalar arrays (3036 floats at P = 20)
constexpr int kDxCs = kDxScal, kDxXs = kDxScal + kDxMaxPos * kDxRowB;
static_assert(kDxXs + kDxMaxPos * kDxRowB <= kMbarOff, "staged dynamic expansion must fit under the mbarrier");

template <typename T>
__device__ __forceinline__ void dyn_exp_row_staged(const DecState& s, int layer, int p, const float* __restrict__ qexp,
                                                   const float* __restrict__ bexp, const int* row_len, const float* x_in,
                                                   long ldxi, float* x_out, long ldxo, const float* __restrict__ ln_g,
                                                   const float* __restrict__ ln_b, T* ln_out, long ldn, int r, char* smem,
                                                   uint32_t& mphase, uint32_t& mphase2, unsigned long long* fine = nullptr,
                                                   int* fmark = nullptr) {
  constexpr int d = 512, n_exp = 16;
#ifdef XN_MEGA_FINE
  auto stamp = [&]() { if (fine && blockIdx.x == 0 && threadIdx.x == 0 && *fmark < 124) fine[(*fmark)++] = gtimer(); };
#else
  auto stamp = []() {};
#endif
  stamp();
  float* sm = reinterpret_cast<float*>(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = s.P, np = p + 1;
  // two mbarriers: wave 3 is issued while wave 2 may still be in flight
  const uint32_t mbar = smem_u32(smem + kMbarOff), mbar2 = mbar + 8;
  auto ln_tail = [&](float v0, float v1, float* red) {       // columns tid and tid + 256
    float v[2] = {v0, v1};
    float mean, rstd;
    block_ln_stats<2>(v, d, red, mean, rstd);
    ln_out[(long)r * ldn + tid] = from_f32<T>((v0 - mean) * rstd * ln_g[tid] + ln_b[tid]);
    ln_out[(long)r * ldn + tid + 256] = from_f32<T>((v1 - mean) * rstd * ln_g[tid + 256] + ln_b[tid + 256]);
  };
  if (row_len && p >= row_len[r]) {            // padded position: the block contributes 0 (all-zero mask rows)
    for (int c = tid; c < d; c += kMegaThreads) x_out[(long)r * ldxo + c] = x_in[(long)r * ldxi + c];
    ln_tail(x_in[(long)r * ldxi + tid], x_in[(long)r * ldxi + tid + 256], sm);
    return;
  }
  const int P4 = (P + 3) & ~3, E4 = 16;
  int* slot = reinterpret_cast<int*>(sm);      // [P4]
  float* ck_row = sm + P4;                     // [P4]  c_p . K_j
  float* ck_col = ck_row + P4;                 // [P4]  c_i . K_p
  float* qkp = ck_col + P4;                    // [E4]
  float* af = qkp + E4;                        // [P][n_exp] forward weights of the new row-block (A), key-major
  float* bf = af + P4 * E4;
  float* ab = bf + P4 * E4;                    // [P][n_exp] backward weights (A)
  float* bb = ab + P4 * E4;
  float* wA = bb + P4 * E4;                    // [P4]
  float* wB = wA + P4;
  float* tA = wB + P4;
  float* tB = tA + P4;
  float* sA = tB + P4;                         // [E4]
  float* sB = sA + E4;
  float* red = sB + E4;                        // [32]
  float* part = red + 32;                      // [2][P][P]; first the history q_e.K_j table, later the mix scratch
  const float* Cs = reinterpret_cast<const float*>(smem + kDxCs);
  const float* Xs = reinterpret_cast<const float*>(smem + kDxXs);

  for (int i = tid; i < np; i += kMegaThreads) slot[i] = (i == p || !s.anc) ? r : s.anc[(long)r * P + i];
  __syncthreads();
  auto crow = [&](int i) { return s.cache + (((long)layer * P + i) * s.R + slot[i]) * s.cw; };
  const float* cp = crow(p);
  stamp();                                                       // slots known
  // ---- wave 1: cond rows -> Cs, key rows -> Xs
  if (warp == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane == 0) mbar_expect(mbar, (uint32_t)(2 * np * kDxRowB));
    __syncwarp();
    for (int i = lane; i < 2 * np; i += 32) {
      const int j = i >> 1, which = i & 1;
      bulk_g2s(smem_u32(smem + (which ? kDxXs : kDxCs) + j * kDxRowB), crow(j) + (which ? d : 0), kDxRowB, mbar);
    }
  }
  // history q_e.K_j of this row's ancestry (one 64-byte read per position), selector / residual of the output
  float4* qkh = reinterpret_cast<float4*>(part);                 // [np][16]
  if (tid < np * 4 && (tid >> 2) != p)
    qkh[tid] = *reinterpret_cast<const float4*>(s.qk + (((long)layer * P + (tid >> 2)) * s.R + slot[tid >> 2]) * n_exp + (tid & 3) * 4);
  // q_e . K_p from global while the rows land: warp w takes e = w and w + 8
  {
    const float* Kp = cp + d;
    float4 k4[4], q0[4], q1[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int c = it * 128 + lane * 4;
      k4[it] = *reinterpret_cast<const float4*>(Kp + c);
      q0[it] = *reinterpret_cast<const float4*>(qexp + (long)warp * d + c);
      q1[it] = *reinterpret_cast<const float4*>(qexp + (long)(warp + 8) * d + c);
    }
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      a0 = fmaf(q0[it].x, k4[it].x, a0); a0 = fmaf(q0[it].y, k4[it].y, a0); a0 = fmaf(q0[it].z, k4[it].z, a0); a0 = fmaf(q0[it].w, k4[it].w, a0);
      a1 = fmaf(q1[it].x, k4[it].x, a1); a1 = fmaf(q1[it].y, k4[it].y, a1); a1 = fmaf(q1[it].z, k4[it].z, a1); a1 = fmaf(q1[it].w, k4[it].w, a1);
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) { qkp[warp] = a0; qkp[warp + 8] = a1; }
  }
  stamp();                                                       // wave 1 issued, q_e.K_p done
  tc5::mbar_wait(mbar, mphase);
  mphase ^= 1u;
  stamp();                                                       // wave 1 landed
  // ---- phase A from shared memory: c_p . K_j (j <= p) and c_i . K_p (i < p), one warp per dot product
  for (int t = warp; t < np + p; t += kMegaThreads / 32) {
    const float* u = t < np ? Cs + p * d : Cs + (t - np) * d;
    const float* v = t < np ? Xs + t * d : Xs + p * d;
    float a = 0.f;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const float4 x4 = *reinterpret_cast<const float4*>(u + it * 128 + lane * 4);
      const float4 y4 = *reinterpret_cast<const float4*>(v + it * 128 + lane * 4);
      a = fmaf(x4.x, y4.x, a); a = fmaf(x4.y, y4.y, a); a = fmaf(x4.z, y4.z, a); a = fmaf(x4.w, y4.w, a);
    }
    a = warp_sum(a);
    if (lane == 0) { if (t < np) ck_row[t] = a; else ck_col[t - np] = a; }
  }
  __syncthreads();
  // ---- wave 2: class-A rows over the keys (every warp is past its last key read)
  if (warp == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane == 0) mbar_expect(mbar2, (uint32_t)(np * kDxRowB));
    __syncwarp();
    for (int j = lane; j < np; j += 32) bulk_g2s(smem_u32(smem + kDxXs + j * kDxRowB), crow(j) + 2 * d, kDxRowB, mbar2);
  }
  if (tid == 0) ck_col[p] = ck_row[p];
  float* qk_out = s.qk + (((long)layer * P + p) * s.R + r) * n_exp;
  if (tid < n_exp) qk_out[tid] = qkp[tid];
  __syncthreads();

  stamp();                                                       // phase A done
  // ---- phase B: scalar work
  const float sq = sqrtf((float)d);
  float* fw_out = s.fw + (((long)layer * P + p) * s.R + r) * (2L * n_exp * P);
  {
    const float* qkh_f = reinterpret_cast<const float*>(qkh);
    for (int e = warp; e < n_exp; e += kMegaThreads / 32) {      // forward weights of the new row-block: one warp per e (np <= 32)
      const int j = lane;
      float z = 0.f, sa = 0.f, sb = 0.f;
      if (j < np) {
        const float qk = (j == p) ? qkp[e] : qkh_f[j * 16 + e];
        z = (qk + ck_row[j]) / sq;
        sa = fmaxf(z, 0.f);
        sb = fmaxf(-z, 0.f);
      }
      sa = warp_sum(sa) + kExpEps;
      sb = warp_sum(sb) + kExpEps;
      if (j < np) {
        const float a = fmaxf(z, 0.f) / sa, b = fmaxf(-z, 0.f) / sb;
        af[j * n_exp + e] = a; bf[j * n_exp + e] = b;
        fw_out[j * n_exp + e] = a; fw_out[(long)n_exp * P + j * n_exp + e] = b;
      }
    }
  }
  float la = 0.f, lb = 0.f;
  for (int i = tid; i < np * n_exp; i += kMegaThreads) {         // backward weights for output position p (column p of z)
    const int pi = i / n_exp, e = i % n_exp;
    const float z = (qkp[e] + ck_col[pi]) / sq;
    const float a = fmaxf(z, 0.f), b = fmaxf(-z, 0.f);
    ab[i] = a; bb[i] = b;
    la += a; lb += b;
  }
  const float ta = block_sum(la, red) + kExpEps;
  const float tb = block_sum(lb, red) + kExpEps;
  __syncthreads();
  for (int i = tid; i < np * n_exp; i += kMegaThreads) { ab[i] = ab[i] / ta; bb[i] = bb[i] / tb; }
  __syncthreads();
  // wA[j] = sum_{i>=j} sum_e ab[(i,e)] * Af_i[e][j]: one task per (which, i, j<=i); a thread's (<= 4) tasks load first
  {
    constexpr int TPT = 2;                                       // 2 * 20 * 20 = 800 <= 2 batches x 2 x 256
#pragma unroll 1
    for (int b0 = 0; b0 < 2; ++b0) {
      float4 f4[TPT][4];
      int ti[TPT], tj[TPT], tw[TPT];
#pragma unroll
      for (int u = 0; u < TPT; ++u) {
        const int t = tid + (b0 * TPT + u) * kMegaThreads;
        const int which = t / (np * np), rem = t % (np * np), i = rem / np, j = rem % np;
        const bool ok = t < 2 * np * np && j <= i;
        ti[u] = ok ? i : -1; tj[u] = j; tw[u] = which;
        if (ok) {
          const float* f = (i == p) ? (which ? bf : af)
                                    : s.fw + (((long)layer * P + i) * s.R + slot[i]) * (2L * n_exp * P) + (which ? (long)n_exp * P : 0);
          const float4* fj = reinterpret_cast<const float4*>(f + j * n_exp);
          f4[u][0] = fj[0]; f4[u][1] = fj[1]; f4[u][2] = fj[2]; f4[u][3] = fj[3];
        }
      }
#pragma unroll
      for (int u = 0; u < TPT; ++u) {
        if (ti[u] >= 0) {
          const float* wi = (tw[u] ? bb : ab) + ti[u] * n_exp;
          float acc = 0.f;
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            acc = fmaf(wi[e4 * 4], f4[u][e4].x, acc); acc = fmaf(wi[e4 * 4 + 1], f4[u][e4].y, acc);
            acc = fmaf(wi[e4 * 4 + 2], f4[u][e4].z, acc); acc = fmaf(wi[e4 * 4 + 3], f4[u][e4].w, acc);
          }
          part[(tw[u] * P + ti[u]) * P + tj[u]] = acc;
        }
      }
    }
  }
  __syncthreads();
  for (int t = tid; t < 2 * np; t += kMegaThreads) {
    const int j = t >> 1, which = t & 1;
    float acc = 0.f;
    for (int i = j; i < np; ++i) acc += part[(which * P + i) * P + j];
    (which ? wB : wA)[j] = acc;
  }
  for (int t = tid; t < 2 * np; t += kMegaThreads) {             // tA[i] = sum_e ab[(i,e)]
    const int i = t >> 1, which = t & 1;
    const float* wsrc = which ? bb : ab;
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < n_exp; ++e) acc += wsrc[i * n_exp + e];
    (which ? tB : tA)[i] = acc;
  }
  for (int t = tid; t < 2 * n_exp; t += kMegaThreads) {          // sA[e] = sum_i ab[(i,e)]
    const int e = t >> 1, which = t & 1;
    const float* wsrc = which ? bb : ab;
    float acc = 0.f;
    for (int i = 0; i < np; ++i) acc += wsrc[i * n_exp + e];
    (which ? sB : sA)[e] = acc;
  }
  __syncthreads();

  stamp();                                                       // phase B done
  // ---- phase C: the d-wide mixes.  Thread t owns columns 4 (t % 128) .. +3 and one half of the history (t / 128)
  const int c4 = (tid & 127) * 4, half = tid >> 7;
  const int j0 = half ? (np + 1) / 2 : 0, j1 = half ? np : (np + 1) / 2;
  float4 sl4 = make_float4(0.f, 0.f, 0.f, 0.f), xi = sl4;
  if (!half) {                                                   // selector and residual: in flight during the mixes
    sl4 = *reinterpret_cast<const float4*>(cp + 4 * d + c4);
    xi = *reinterpret_cast<const float4*>(x_in + (long)r * ldxi + c4);
  }
  float4 oa = make_float4(0.f, 0.f, 0.f, 0.f), ob = oa;
  for (int j = j0; j < j1; ++j) {                                // cond rows
    const float4 cj = *reinterpret_cast<const float4*>(Cs + j * d + c4);
    const float ta_ = tA[j], tb_ = tB[j];
    oa.x = fmaf(ta_, cj.x, oa.x); oa.y = fmaf(ta_, cj.y, oa.y); oa.z = fmaf(ta_, cj.z, oa.z); oa.w = fmaf(ta_, cj.w, oa.w);
    ob.x = fmaf(tb_, cj.x, ob.x); ob.y = fmaf(tb_, cj.y, ob.y); ob.z = fmaf(tb_, cj.z, ob.z); ob.w = fmaf(tb_, cj.w, ob.w);
  }
  __syncthreads();                                               // the cond rows are done with
  if (warp == 0) {                                               // ---- wave 3: class-B rows over the cond rows
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane == 0) mbar_expect(mbar, (uint32_t)(np * kDxRowB));
    __syncwarp();
    for (int j = lane; j < np; j += 32) bulk_g2s(smem_u32(smem + kDxCs + j * kDxRowB), crow(j) + 3 * d, kDxRowB, mbar);
  }
  if (half) {                                                    // expansion-bias term (weights: L2 / L1), all loads first
#pragma unroll 1
    for (int e0 = 0; e0 < 16; e0 += 8) {
      float4 be[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) be[e] = *reinterpret_cast<const float4*>(bexp + (long)(e0 + e) * d + c4);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float wa = sA[e0 + e], wb = sB[e0 + e];
        oa.x = fmaf(wa, be[e].x, oa.x); oa.y = fmaf(wa, be[e].y, oa.y); oa.z = fmaf(wa, be[e].z, oa.z); oa.w = fmaf(wa, be[e].w, oa.w);
        ob.x = fmaf(wb, be[e].x, ob.x); ob.y = fmaf(wb, be[e].y, ob.y); ob.z = fmaf(wb, be[e].z, ob.z); ob.w = fmaf(wb, be[e].w, ob.w);
      }
    }
  }
  tc5::mbar_wait(mbar2, mphase2);                                // wave 2 (class-A rows)
  mphase2 ^= 1u;
  for (int j = j0; j < j1; ++j) {
    const float4 aj = *reinterpret_cast<const float4*>(Xs + j * d + c4);
    const float wa = wA[j];
    oa.x = fmaf(wa, aj.x, oa.x); oa.y = fmaf(wa, aj.y, oa.y); oa.z = fmaf(wa, aj.z, oa.z); oa.w = fmaf(wa, aj.w, oa.w);
  }
  tc5::mbar_wait(mbar, mphase);                                  // wave 3 (class-B rows)
  mphase ^= 1u;
  for (int j = j0; j < j1; ++j) {
    const float4 bj = *reinterpret_cast<const float4*>(Cs + j * d + c4);
    const float wb = wB[j];
    ob.x = fmaf(wb, bj.x, ob.x); ob.y = fmaf(wb, bj.y, ob.y); ob.z = fmaf(wb, bj.z, ob.z); ob.w = fmaf(wb, bj.w, ob.w);
  }
  stamp();                                                       // mixes done
  float4* mix = reinterpret_cast<float4*>(part);                 // [2 (a|b)][128] partial sums of the upper half
  if (half) {
    mix[tid & 127] = oa;
    mix[128 + (tid & 127)] = ob;
  }
  __syncthreads();
  float* xo_s = reinterpret_cast<float*>(mix + 256);             // the row, for the LayerNorm tail's column assignment
  if (!half) {
    const float4 ua = mix[tid], ub = mix[128 + tid];
    oa.x += ua.x; oa.y += ua.y; oa.z += ua.z; oa.w += ua.w;
    ob.x += ub.x; ob.y += ub.y; ob.z += ub.z; ob.w += ub.w;
    const float s0 = sigmoidf_(sl4.x), s1 = sigmoidf_(sl4.y), s2 = sigmoidf_(sl4.z), s3 = sigmoidf_(sl4.w);
    float4 xo;
    xo.x = xi.x + (s0 * oa.x + (1.0f - s0) * ob.x);
    xo.y = xi.y + (s1 * oa.y + (1.0f - s1) * ob.y);
    xo.z = xi.z + (s2 * oa.z + (1.0f - s2) * ob.z);
    xo.w = xi.w + (s3 * oa.w + (1.0f - s3) * ob.w);
    *reinterpret_cast<float4*>(x_out + (long)r * ldxo + c4) = xo;
    *reinterpret_cast<float4*>(xo_s + c4) = xo;
  }
  __syncthreads();
  ln_tail(xo_s[tid], xo_s[tid + 256], red);
  stamp();                                                       // row stored, norm_2 written
}

// One decoder position: all phases up to the vocabulary projection (and, with merge_rows, the row-wise top-k merge).
// p, the token table and the ancestry table are arguments: the whole-search kernel calls this once per time step.
template <typename T>
__device__ __forceinline__ void mega_position(const MegaArgs& a, const DecState& st, int p, const int64_t* tok64, const int* tok32,
                                              bool merge_rows, GridBar& bar, char* smem, uint32_t& mphase, uint32_t& mphase2, int& mark,
                                              int& fmark) {
  const int tid = threadIdx.x;
  const int d = 512, R = a.R, nd = a.n_layers;
  const long ldc = (long)d * nd;
  const long m32 = (long)((R + kBM - 1) / kBM) * kBM;        // rows of a packed 16-bit activation slab
  T* xn = reinterpret_cast<T*>(a.xn);
  T* att = reinterpret_cast<T*>(a.att);
  T* hid = reinterpret_cast<T*>(a.hid);
  T* ycat16 = reinterpret_cast<T*>(a.ycat16);
  const T* kv = reinterpret_cast<const T*>(a.kv);
  (void)tid;
  // Every GEMM phase is its own inlined, specialised copy of gemm_phase (null pointers and constant shapes folded away).
  // One shared copy per tile variant driven by a run-time descriptor was measured: 17 us slower per position, although the
  // kernel's code is fetched cold at every position (ncu: 18 % of the non-barrier warp samples are "no instruction").
#define XN_GEMM_INIT(g)                                                                                   \
  GemmP g{};                                                                                              \
  g.dbg_mode = a.dbg_mode; g.fine = (a.dbg && l == 0 && !(a.dbg_mode & 16)) ? a.dbg : nullptr; g.fmark = &fmark; \
  g.M = R; g.N = d; g.K = d;
  bool w_ready = false;                          // the next GEMM phase's first W tile is already in flight
  int l = 0;
  for (; l < nd; ++l) {
    const MegaLayer& W = a.L[l];
    const float* xin = l == 0 ? a.x0 : a.ycat + (size_t)(l - 1) * d;
    const long ldi = l == 0 ? d : ldc;
    float* xout = a.ycat + (size_t)l * d;
    {                                            // [cond | key | A | B | selector] = LN1(x) W5^T + b    (layers.py:152-170, 226-229)
      XN_GEMM_INIT(g)
      g.A32 = xin; g.lda32 = ldi; g.ln_g = W.n1g; g.ln_b = W.n1b;
      if (l == 0) {                              // the embedding is computed in the fill (no separate phase, no barrier)
        g.emb = a.emb; g.pos_row = a.pos + (long)p * d; g.tok64 = tok64; g.tok32 = tok32; g.tok_stride = a.tok_stride;
        g.tok_p = p; g.x0_out = a.x0;
      }
      g.W = W.w_dyn5; g.bias = W.b_dyn5; g.Cf = st.cache + (((size_t)l * st.P + p) * R) * st.cw; g.ldcf = st.cw; g.N = 5 * d;
      gemm_phase_impl<T, 64, true>(g, smem, mphase, mphase2, w_ready);
    }
    bar.sync(a.dbg, &mark);
    // incremental dynamic expansion of position p + residual, then norm_2 -> xn
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
      dyn_exp_row_staged<T>(st, l, p, W.qexp, W.bexp, a.row_len, xin, ldi, xout, ldc, W.n2g, W.n2b, xn, kPadK, r, smem, mphase,
                            mphase2, (a.dbg_mode & 16) && l == 0 ? a.dbg : nullptr, &fmark);
      __syncthreads();
    }
    w_ready = prefetch_w<T, 32, false>(W.w_wq, R, d, 1, smem, a.dbg_mode);
    bar.sync(a.dbg, &mark);
    {                                            // q = xn Wq^T + b
      XN_GEMM_INIT(g)
      g.A16 = xn; g.W = W.w_wq; g.bias = W.b_wq; g.Cf = a.q; g.ldcf = d;
      gemm_phase_impl<T, 32, false>(g, smem, mphase, mphase2, w_ready);
    }
    w_ready = prefetch_w<T, 32, false>(W.w_wo, R, d, 1, smem, a.dbg_mode);      // lands during the cross attention
    bar.sync(a.dbg, &mark);
    {                                            // cross attention, one item per (image, head, group of <= 4 beam rows)   (layers.py:266-295)
      const int rpi = a.rows_per_image, n_img = R / rpi, ngrp = (rpi + 3) / 4, heads = a.heads;
      const int k_off = l * 2 * d, v_off = l * 2 * d + d;
      for (int it = blockIdx.x; it < n_img * heads * ngrp; it += gridDim.x) {
        const int gq = it % ngrp, hh = (it / ngrp) % heads, b = it / (ngrp * heads);
        const int row0 = b * rpi + gq * 4, cnt = min(4, rpi - gq * 4);
        float* smf = reinterpret_cast<float*>(smem);
        switch (cnt) {
          case 1: cross_attn16_item<T, 1>(a.q, d, kv, a.ldkv, k_off, v_off, att, kPadK, a.n_keys, a.n_valid, a.row_len, p, b, hh, row0, smf); break;
          case 2: cross_attn16_item<T, 2>(a.q, d, kv, a.ldkv, k_off, v_off, att, kPadK, a.n_keys, a.n_valid, a.row_len, p, b, hh, row0, smf); break;
          case 3: cross_attn16_item<T, 3>(a.q, d, kv, a.ldkv, k_off, v_off, att, kPadK, a.n_keys, a.n_valid, a.row_len, p, b, hh, row0, smf); break;
          default: cross_attn16_item<T, 4>(a.q, d, kv, a.ldkv, k_off, v_off, att, kPadK, a.n_keys, a.n_valid, a.row_len, p, b, hh, row0, smf); break;
        }
        __syncthreads();
      }
    }
    bar.sync(a.dbg, &mark);
    {                                            // x = x + att Wo^T + b
      XN_GEMM_INIT(g)
      g.A16 = att; g.W = W.w_wo; g.bias = W.b_wo; g.res = xout; g.ldr = ldc; g.Cf = xout; g.ldcf = ldc;
      gemm_phase_impl<T, 32, false>(g, smem, mphase, mphase2, w_ready);
    }
    w_ready = prefetch_w<T, 64, true>(W.w_ff1, R, a.ff, 1, smem, a.dbg_mode);
    bar.sync(a.dbg, &mark);
    {                                            // hid = relu(LN3(x) W1^T + b)
      XN_GEMM_INIT(g)
      g.A32 = xout; g.lda32 = ldc; g.ln_g = W.n3g; g.ln_b = W.n3b;
      g.W = W.w_ff1; g.bias = W.b_ff1; g.Cb = hid; g.N = a.ff; g.act = 2;
      gemm_phase_impl<T, 64, true>(g, smem, mphase, mphase2, w_ready);
    }
    w_ready = prefetch_w<T, 64, false>(W.w_ff2, R, d, a.ksplit_ff2 > 1 ? a.ksplit_ff2 : 1, smem, a.dbg_mode);
    bar.sync(a.dbg, &mark);
    {                                            // x = x + hid W2^T + b   (also kept in 16 bits: operand of the reduce group)
      XN_GEMM_INIT(g)
      g.A16 = hid; g.W = W.w_ff2; g.bias = W.b_ff2; g.res = xout; g.ldr = ldc; g.Cf = xout; g.ldcf = ldc;
      g.Cb = ycat16 + (size_t)l * m32 * kPadK; g.K = a.ff;
      g.ksplit = a.ksplit_ff2; g.scratch = a.scratch; g.cnt = a.bar + kSplitCntOff;
      gemm_phase_impl<T, 64, false>(g, smem, mphase, mphase2, w_ready);
    }
    if (l + 1 < nd) w_ready = prefetch_w<T, 64, true>(a.L[l + 1].w_dyn5, R, 5 * d, 1, smem, a.dbg_mode);
    else w_ready = prefetch_w<T, 64, false>(a.w_reduce, R, d, a.ksplit_red > 1 ? a.ksplit_red : 1, smem, a.dbg_mode);
    bar.sync(a.dbg, &mark);
  }
  l = nd - 1;
  {                                              // reduce group: pre = x_last + [y_1 | .. | y_n] Wr^T + b   (End_ExpansionNet_v2.py:196-199)
    XN_GEMM_INIT(g)
    g.A16 = ycat16; g.W = a.w_reduce; g.bias = a.b_reduce; g.res = a.ycat + (size_t)(nd - 1) * d; g.ldr = ldc;
    g.Cf = a.pre; g.ldcf = d; g.K = d * nd;
    g.ksplit = a.ksplit_red; g.scratch = a.scratch; g.cnt = a.bar + kSplitCntOff;
    gemm_phase_impl<T, 64, false>(g, smem, mphase, mphase2, w_ready);
  }
  w_ready = prefetch_w<T, 64, true>(a.w_vocab, R, a.vocab, 1, smem, a.dbg_mode);
  bar.sync(a.dbg, &mark);
  {                                              // logits = LN(pre) Wv^T + b    (End_ExpansionNet_v2.py:200-204)
    XN_GEMM_INIT(g)
    g.A32 = a.pre; g.lda32 = d; g.ln_g = a.ng; g.ln_b = a.nb;
    g.W = a.w_vocab; g.bias = a.b_vocab; g.N = a.vocab;
    if (a.topk > 0) {
      const int ntn_v = (a.vocab + 63) / 64;
      TopkHook<64> hook{TopkParts{reinterpret_cast<float*>(a.parts), (long)R * ntn_v}, ntn_v, R, a.vocab, a.topk};
      gemm_phase_impl<T, 64, true, false, TopkHook<64>>(g, smem, mphase, mphase2, w_ready, hook);
      bar.sync(a.dbg, &mark);
      if (merge_rows) topk_merge_phase(TopkParts{reinterpret_cast<float*>(a.parts), (long)R * ntn_v}, ntn_v, R, a.topk, a.top_val, a.top_idx);
    } else {
      g.Cf = a.logits; g.ldcf = a.ldl;
      gemm_phase_impl<T, 64, true>(g, smem, mphase, mphase2, w_ready);
    }
  }
#undef XN_GEMM_INIT
}

template <typename T>
__global__ void __launch_bounds__(kMegaThreads, 2) dec_step_mega_kernel(const __grid_constant__ MegaArgs a) {
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x;
  int mark = 1, fmark = 64;
  if (a.dbg && blockIdx.x == 0 && tid == 0) { a.dbg[0] = gtimer(); a.dbg[124] = (unsigned long long)clock64(); }
  GridBar bar;
  bar.init(a.bar);
  uint32_t mphase = 0u, mphase2 = 0u;            // parity of the mbarriers' next completion
  if (tid == 0) {
    tc5::mbar_init(smem_u32(smem + kMbarOff), 1);
    tc5::mbar_init(smem_u32(smem + kMbarOff) + 8, 1);
    tc5::fence_mbar_init();
  }
  __syncthreads();
  mega_position<T>(a, a.s, a.p, a.tok64, a.tok32, true, bar, smem, mphase, mphase2, mark, fmark);
  if (a.dbg && blockIdx.x == 0 && tid == 0) { a.dbg[mark++] = gtimer(); a.dbg[127] = (unsigned long long)mark; a.dbg[126] = (unsigned long long)fmark; a.dbg[125] = (unsigned long long)clock64(); }
}

// ---- the whole beam search ('max' branch) in ONE launch: time steps loop inside the kernel, the bookkeeping of a step
// (captioning_model.py:295-397) runs per image right after the image's rows have merged their top-k, and the reference's
// early `break` (:397, "no beam was extended") is a grid-uniform branch on a flag read after the step's barrier.  Against
// one launch per position this removes the launch gaps (20-40 us per step in a graph) and the conditional graph nodes --
// but MEASURED SLOWER (B200, 64 images x beam 3: 307 vs 265 us per position, 5.8 vs 5.4 ms per search): inside the
// time-step loop the compiler hoists the position's loop-invariant address arithmetic across the loop and spills it in the
// hot phases; as an out-of-line function (ABI register limits: 1.4 KB of spills) it is 391 us.  Kept behind the option
// "mega_search" (default off), under test.
template <typename T>
__global__ void __launch_bounds__(kMegaThreads, 2) dec_search_mega_kernel(const __grid_constant__ MegaArgs a,
                                                                           const __grid_constant__ MegaSearch q) {
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int mark = 1, fmark = 64;
  if (a.dbg && blockIdx.x == 0 && tid == 0) { a.dbg[0] = gtimer(); a.dbg[124] = (unsigned long long)clock64(); }
  GridBar bar;
  bar.init(a.bar);
  uint32_t mphase = 0u, mphase2 = 0u;
  if (tid == 0) {
    tc5::mbar_init(smem_u32(smem + kMbarOff), 1);
    tc5::mbar_init(smem_u32(smem + kMbarOff) + 8, 1);
    tc5::fence_mbar_init();
  }
  __syncthreads();
  const int beam = q.beam, L = q.L, B = a.R / beam, ntn_v = (a.vocab + 63) / 64;
  const TopkParts parts{reinterpret_cast<float*>(a.parts), (long)a.R * ntn_v};
  int src = 0;
  for (int t = 1; t < L; ++t) {                  // choosing token t: decode position t - 1
    DecState st = a.s;
    st.anc = q.bb.anc[src];
    if (a.dbg && blockIdx.x == 0 && tid == 0) a.dbg[0] = gtimer();       // the timeline keeps the last executed step
    mark = 1;
    mega_position<T>(a, st, t - 1, nullptr, q.bb.tokens[src], false, bar, smem, mphase, mphase2, mark, fmark);
    // per image: its beam rows merge their top-k (one warp per row), then one warp runs the step's bookkeeping
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
      if (warp < beam) topk_merge_row(parts, ntn_v, b * beam + warp, a.topk, a.top_val, a.top_idx, lane);
      __syncthreads();
      if (t == 1) { if (tid < beam) beam_first_row(q.bb, a.top_val, a.top_idx, beam, L, q.eos, b * beam + tid); }
      else if (warp == 0) beam_step_image(q.bb, src, a.top_val, a.top_idx, beam, L, t, q.eos, b, lane);
      __syncthreads();
    }
    bar.sync(a.dbg, &mark);
    if (t > 1) {
      src ^= 1;
      if (q.early_exit && *reinterpret_cast<volatile int*>(q.bb.grew + t) == 0) break;     // every beam had ended (:397)
    }
  }
  for (int b = blockIdx.x * kMegaThreads + tid; b < B; b += gridDim.x * kMegaThreads)
    beam_finalize_image(q.bb, beam, L, L, q.how_many, q.r_tok, q.r_len, q.r_lp, b);
  if (a.dbg && blockIdx.x == 0 && tid == 0) { a.dbg[mark++] = gtimer(); a.dbg[127] = (unsigned long long)mark; a.dbg[126] = (unsigned long long)fmark; a.dbg[125] = (unsigned long long)clock64(); }
}

// fp32 (N x K) -> 16-bit slabs [K/512][N][520] (pad columns zero): the layout gemm_phase copies whole tiles from
template <typename T>
__global__ void mega_pack_kernel(const float* __restrict__ w, T* __restrict__ out, int N, int K) {
  const long total = (long)(K / kKI) * N * kPadK;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % kPadK);
    const long rn = i / kPadK;
    const int n = (int)(rn % N), sp = (int)(rn / N);
    out[i] = from_f32<T>(c < kKI ? w[(long)n * K + sp * kKI + c] : 0.f);
  }
}
size_t mega_packed_bytes(int N, int K) { return (size_t)(K / kKI) * N * kPadK * 2; }
cudaError_t launch_mega_pack_weight(const float* w, void* out, int N, int K, int fp16, cudaStream_t st) {
  if (K % kKI) return cudaErrorInvalidValue;
  if (fp16) mega_pack_kernel<f16><<<592, 256, 0, st>>>(w, reinterpret_cast<f16*>(out), N, K);
  else mega_pack_kernel<bf16><<<592, 256, 0, st>>>(w, reinterpret_cast<bf16*>(out), N, K);
  return cudaGetLastError();
}
// packed 16-bit activations of one decoder position: xn | att | hid (ff/512 slabs) | layer outputs (n_layers slabs)
size_t mega_act_bytes(int R, int ff, int n_layers) {
  const size_t m32 = (size_t)((R + kBM - 1) / kBM) * kBM;
  return (size_t)(2 + ff / kKI + n_layers) * m32 * kPitch;
}
size_t mega_scratch_bytes(int R) { return (size_t)((R + kBM - 1) / kBM) * (512 / 64) * 4 * (kBM * 64 * sizeof(float)); }
size_t mega_parts_bytes(int R, int vocab) { return (size_t)R * ((vocab + 63) / 64) * (2 + 2 * kTopC) * sizeof(float); }

bool mega_supported(const MegaArgs& a) {
  if (a.d != 512 || a.heads * 64 != a.d || a.n_keys > kCaMaxKeys || a.n_layers < 1 || a.n_layers > kMegaMaxLayers) return false;
  if (a.ff % 128 || a.vocab % 4 || a.s.P > kDxMaxPos || a.n_exp != 16 || a.rows_per_image < 1 || a.R % a.rows_per_image) return false;
  if (((a.R + kBM - 1) / kBM) * (a.d / 64) > kMaxSplitTiles && (a.ksplit_ff2 > 1 || a.ksplit_red > 1)) return false;
  if (a.ksplit_ff2 > 4 || a.ksplit_red > 4 || a.ff != kKI * std::max(1, a.ksplit_ff2) || a.d * a.n_layers != kKI * std::max(1, a.ksplit_red)) return false;
  if (a.s.cw != 5 * a.d || (a.ldkv & 7) || (a.vocab + 63) / 64 > 256 || a.topk > kTopC) return false;
  return true;
}

template <typename K>
static cudaError_t mega_launch_config(K kernel, const MegaArgs& a, cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int which) {
  struct DevInfo { int grid = 0; size_t smem = 0; int sms = 0; };
  static DevInfo info[2][32];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 32) dev = 0;
  const size_t smem_rows = dyn_exp_smem_floats(a.s.P, a.n_exp) * sizeof(float);
  const size_t smem = std::max((size_t)kGemmSmem, smem_rows);
  DevInfo& di = info[which][dev];
  if (di.grid == 0 || di.smem < smem) {
    if (cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) return e;
    int per_sm = 0, sms = 0;
    if (cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kMegaThreads, smem)) return e;
    if (cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    di.grid = std::min(per_sm, 2) * sms;
    di.smem = smem;
    di.sms = sms;
  }
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(a.max_ctas_per_sm == 1 ? di.sms : di.grid);
  cfg.blockDim = dim3(kMegaThreads);
  cfg.dynamicSmemBytes = di.smem;
  attr[0].id = cudaLaunchAttributeCooperative;          // all CTAs co-resident or the launch fails: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_mega_coop ? 1 : 0;
  return cudaSuccess;
}

template <typename T>
cudaError_t launch_dec_step_mega(const MegaArgs& a, cudaStream_t st) {
  if (!mega_supported(a)) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  if (cudaError_t e = mega_launch_config(dec_step_mega_kernel<T>, a, cfg, attr, std::is_same<T, f16>::value ? 0 : 1)) return e;
  cfg.stream = st;
  return cudaLaunchKernelEx(&cfg, dec_step_mega_kernel<T>, a);
}
template <typename T>
cudaError_t launch_dec_search_mega(const MegaArgs& a, const MegaSearch& q, cudaStream_t st) {
  if (!mega_supported(a) || a.topk != q.beam || q.beam < 1 || q.beam > kMaxBeam || a.rows_per_image != q.beam) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  static_assert(kMegaThreads / 32 >= kMaxBeam, "one warp per beam row of an image");
  if (cudaError_t e = mega_launch_config(dec_search_mega_kernel<T>, a, cfg, attr, std::is_same<T, f16>::value ? 0 : 1)) return e;
  cfg.stream = st;
  return cudaLaunchKernelEx(&cfg, dec_search_mega_kernel<T>, a, q);
}
template cudaError_t launch_dec_search_mega<f16>(const MegaArgs&, const MegaSearch&, cudaStream_t);
template cudaError_t launch_dec_search_mega<bf16>(const MegaArgs&, const MegaSearch&, cudaStream_t);
template cudaError_t launch_dec_step_mega<f16>(const MegaArgs&, cudaStream_t);
template cudaError_t launch_dec_step_mega<bf16>(const MegaArgs&, cudaStream_t);

}  // namespace xn
