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

namespace xn {

int g_mega_coop = 1;

constexpr int kMegaThreads = 256;
constexpr int kBM = 32, kKI = 512;              // rows per tile; K per work item (longer K is split over CTAs)
constexpr int kPitch = (kKI + 8) * 2;           // 1040 bytes per operand row: consecutive rows shift by 16 bytes -> ldmatrix is conflict-free
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

// ---- grid barrier.  Arrivals are atomic adds on one word (cooperative groups' trick: CTA 0 adds 2^31 - (n-1), the
// others 1, so bit 31 flips exactly on the last arrival and the word needs no reset); the last arriver publishes a
// generation number in a SEPARATE line, which is what everybody polls -- the pollers' loads do not queue in front of the
// arrivals at the L2 slice.  __threadfence() (gpu scope) orders the data and invalidates the SM's L1 (CCTL.IVALL), so plain
// loads after the barrier see what other CTAs wrote before it; fence.proxy.async extends that to the bulk-copy engine.
struct GridBar {
  unsigned* cnt; unsigned* flag; unsigned gen;
  __device__ __forceinline__ void init(unsigned* b) {
    cnt = b;
    flag = b + kBarStride;
    gen = 0u;
    if (threadIdx.x == 0) gen = *reinterpret_cast<volatile unsigned*>(flag);           // read before this CTA's first arrival: cannot have advanced yet
  }
  __device__ __forceinline__ void sync(unsigned long long* dbg = nullptr, int* mark = nullptr) {
    __syncthreads();
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[(*mark)++] = gtimer();
    if (threadIdx.x == 0) {
      // acquire / release instead of __threadfence(): one MEMBAR.ALL.GPU at the arrival (two for the last arriver) in place
      // of three MEMBAR.SC.GPU; the acquire load carries the L1 invalidation
      const unsigned inc = (blockIdx.x == 0) ? (0x80000000u - (gridDim.x - 1)) : 1u;
      unsigned old, v;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(cnt), "r"(inc) : "memory");
      if ((old ^ (old + inc)) & 0x80000000u)     // last arrival
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(gen + 1u) : "memory");
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      } while (v == gen);
      gen += 1u;
    }
    __syncthreads();
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[(*mark)++] = gtimer();
  }
};

struct GemmP {
  const void* A16; long lda;                                  // 16-bit A (M x K), or
  const float* A32; long lda32; const float *ln_g, *ln_b;     // fp32 rows normalised on load (K == 512)
  const void* W; const float* bias;                           // (N x K) 16-bit, K contiguous
  const float* res; long ldr;                                 // fp32 residual added after the activation (may alias Cf)
  float* Cf; long ldcf;                                       // fp32 output and / or
  void* Cb; long ldcb;                                        // 16-bit output
  int M, N, K, act;                                           // act: 0 none, 2 ReLU;  K = 512 * ksplit
  int ksplit; float* scratch; unsigned* cnt;                  // split-K over CTAs: partial tiles + per-tile arrival counters
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

#define XN_FINE(g) do { if ((g).fine && blockIdx.x == 0 && threadIdx.x == 0 && *(g).fmark < 126) (g).fine[(*(g).fmark)++] = gtimer(); } while (0)

// One GEMM phase.  A work item is a 32 x BN output tile over K = 512 (longer K: ksplit items per tile).  Both operands
// of an item are brought in whole -- one 1 KB bulk copy (cp.async.bulk) per operand row, completing on an mbarrier: no
// per-thread copy instructions, no staging ring -- or, for a LayerNorm operand, normalised from the fp32 rows straight into
// shared memory.  The eight warps each take a 64-wide slice of K with a 32 x BN accumulator (8 or 16 independent MMAs per
// k-step, no barrier inside the item); the eight partial tiles meet in shared memory, where bias / activation / residual
// run with coalesced 16-byte accesses.  With ksplit > 1 the partial tile goes to `scratch` and the CTA that arrives last
// at the tile's counter adds the splits in order (deterministic) and finishes.
template <typename T, int BN, bool LN, bool STORE = true, typename Hook = NoTileHook>
__device__ __forceinline__ void gemm_phase(const GemmP& g, char* smem, uint32_t& mphase, const Hook& hook = Hook()) {
  constexpr int NB2 = BN / 16, NF = BN / 8, PITCH = BN + 8, C4 = BN / 4;
  static_assert(8 * kBM * PITCH * 4 <= kMbarOff, "reduction tiles must fit in the operand regions");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = g.ksplit > 1 ? g.ksplit : 1;
  const int ntm = (g.M + kBM - 1) / kBM, ntn = (g.N + BN - 1) / BN, ntiles = ntm * ntn;
  char* a_sm = smem;
  char* w_sm = smem + kARegion;
  float* red = reinterpret_cast<float*>(smem);                   // [8][32][PITCH] after the MMAs
  const uint32_t mbar = smem_u32(smem + kMbarOff);
  __shared__ int s_last;
  const T* W = reinterpret_cast<const T*>(g.W);
  const T* A16 = reinterpret_cast<const T*>(g.A16);
  for (int item = blockIdx.x; item < ntiles * S; item += gridDim.x) {
    const int tile = item / S, sp = item - tile * S;
    const int tm = tile % ntm, tn = tile / ntm, r0 = tm * kBM, n0 = tn * BN, k0 = sp * kKI;
    XN_FINE(g);                                          // 0: item start
    const bool do_load = !(g.dbg_mode & 2);
    if (do_load && warp == 0) {
      // the regions were last touched through the generic proxy (ldmatrix / reduction) by this CTA, all before the
      // __syncthreads that ended the previous item
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (lane == 0) mbar_expect(mbar, (uint32_t)((BN + (LN ? 0 : kBM)) * kKI * 2));
      __syncwarp();
#pragma unroll
      for (int i = lane; i < BN + (LN ? 0 : kBM); i += 32) {
        if (i < BN) {
          const int nrow = min(n0 + i, g.N - 1);
          bulk_g2s(smem_u32(w_sm + i * kPitch), W + (long)nrow * g.K + k0, kKI * 2, mbar);
        } else {
          const int row = i - BN, arow = min(r0 + row, g.M - 1);
          bulk_g2s(smem_u32(a_sm + row * kPitch), A16 + (long)arow * g.lda + k0, kKI * 2, mbar);
        }
      }
    }
    if (LN && !(g.dbg_mode & 4)) {
      // LayerNorm(gamma, beta) of rows r0 .. r0+31 (K == 512): four rows per warp, all their loads in flight together; same
      // arithmetic and summation order as layernorm_kernel (elementwise.cu: ln_row)
      float4 v[4][4];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int grow = r0 + warp * 4 + rr;
        const float* xr = g.A32 + (long)min(grow, g.M - 1) * g.lda32;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[rr][j] = __ldcg(reinterpret_cast<const float4*>(xr + j * 128 + lane * 4));
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
    }
    if (LN) __syncthreads();                             // the normalised rows of all warps
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
    {
      float* rq = red + (size_t)warp * kBM * PITCH;
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
          const int row = mb * 16 + (lane >> 2), col = nf * 8 + (lane & 3) * 2;
          *reinterpret_cast<float2*>(rq + row * PITCH + col) = make_float2(acc[mb][nf][0], acc[mb][nf][1]);
          *reinterpret_cast<float2*>(rq + (row + 8) * PITCH + col) = make_float2(acc[mb][nf][2], acc[mb][nf][3]);
        }
    }
    __syncthreads();
    auto sum8 = [&](int row, int c4) {
      float4 v = *reinterpret_cast<const float4*>(red + row * PITCH + c4 * 4);
#pragma unroll
      for (int q = 1; q < 8; ++q) {
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
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const unsigned old = atomicAdd(g.cnt + tile, 1u);
        s_last = old == (unsigned)(S - 1);
        if (s_last) { g.cnt[tile] = 0u; __threadfence(); }
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
              *reinterpret_cast<uint2*>(reinterpret_cast<T*>(g.Cb) + (long)grow * g.ldcb + col) = u;
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
__device__ __forceinline__ void topk_merge_phase(const TopkParts& parts, int ntn, int R, int k, float* top_val, int* top_idx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int TPL = 8;                                        // tiles per lane (ntn <= 256)
  for (int r = blockIdx.x * (kMegaThreads / 32) + warp; r < R; r += gridDim.x * (kMegaThreads / 32)) {
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

template <typename T>
__global__ void __launch_bounds__(kMegaThreads, 2) dec_step_mega_kernel(const __grid_constant__ MegaArgs a) {
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = 512, R = a.R, nd = a.n_layers, p = a.p;
  const long ldc = (long)d * nd;
  T* xn = reinterpret_cast<T*>(a.xn);
  T* att = reinterpret_cast<T*>(a.att);
  T* hid = reinterpret_cast<T*>(a.hid);
  T* ycat16 = reinterpret_cast<T*>(a.ycat16);
  const T* kv = reinterpret_cast<const T*>(a.kv);
  int mark = 1, fmark = 64;
  if (a.dbg && blockIdx.x == 0 && tid == 0) a.dbg[0] = gtimer();
  GridBar bar;
  bar.init(a.bar);
  uint32_t mphase = 0u;                          // parity of the operand mbarrier's next completion
  if (tid == 0) {
    tc5::mbar_init(smem_u32(smem + kMbarOff), 1);
    tc5::fence_mbar_init();
  }
  __syncthreads();

  // ---- phase 0: x0 = E[tok] sqrt(d) + P[p]   (layers.py:16-17); one warp per row
  {
    const float sc = sqrtf((float)d);
    for (int r = blockIdx.x * 8 + warp; r < R; r += gridDim.x * 8) {
      const long tok = a.tok64 ? (long)a.tok64[r * a.tok_stride + p] : (long)a.tok32[r * a.tok_stride + p];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = j * 128 + lane * 4;
        const float4 e4 = *reinterpret_cast<const float4*>(a.emb + tok * d + c);
        const float4 p4 = *reinterpret_cast<const float4*>(a.pos + (long)p * d + c);
        *reinterpret_cast<float4*>(a.x0 + (long)r * d + c) =
            make_float4(e4.x * sc + p4.x, e4.y * sc + p4.y, e4.z * sc + p4.z, e4.w * sc + p4.w);
      }
    }
  }
  bar.sync(a.dbg, &mark);

  for (int l = 0; l < nd; ++l) {
    const MegaLayer& W = a.L[l];
    const float* xin = l == 0 ? a.x0 : a.ycat + (size_t)(l - 1) * d;
    const long ldi = l == 0 ? d : ldc;
    float* xout = a.ycat + (size_t)l * d;
    float* crow = a.s.cache + (((size_t)l * a.s.P + p) * R) * a.s.cw;
    // P1: [cond | key | A | B | selector] = LN1(x) W5^T + b    (layers.py:152-170, 226-229)
    {
      GemmP g{};
    g.dbg_mode = a.dbg_mode; g.fine = a.dbg; g.fmark = &fmark;
      g.dbg_mode = a.dbg_mode; g.fine = (a.dbg && l == 0) ? a.dbg : nullptr; g.fmark = &fmark;
      g.A32 = xin; g.lda32 = ldi; g.ln_g = W.n1g; g.ln_b = W.n1b;
      g.W = W.w_dyn5; g.bias = W.b_dyn5; g.Cf = crow; g.ldcf = a.s.cw; g.M = R; g.N = 5 * d; g.K = d;
      gemm_phase<T, 64, true>(g, smem, mphase);
    }
    bar.sync(a.dbg, &mark);
    // P2: incremental dynamic expansion of position p + residual, then norm_2 -> xn
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
      dyn_exp_row<T>(a.s, l, p, W.qexp, W.bexp, a.n_exp, a.row_len, xin, ldi, xout, ldc, d, W.n2g, W.n2b, xn, d, r,
                     reinterpret_cast<float*>(smem));
      __syncthreads();
    }
    bar.sync(a.dbg, &mark);
    // P3: q = xn Wq^T + b
    {
      GemmP g{};
    g.dbg_mode = a.dbg_mode; g.fine = a.dbg; g.fmark = &fmark;
      g.dbg_mode = a.dbg_mode; g.fine = (a.dbg && l == 0) ? a.dbg : nullptr; g.fmark = &fmark;
      g.A16 = xn; g.lda = d; g.W = W.w_wq; g.bias = W.b_wq; g.Cf = a.q; g.ldcf = d; g.M = R; g.N = d; g.K = d;
      gemm_phase<T, 32, false>(g, smem, mphase);
    }
    bar.sync(a.dbg, &mark);
    // P4: cross attention, one item per (image, head, group of <= 4 beam rows)   (layers.py:266-295)
    {
      const int rpi = a.rows_per_image, n_img = R / rpi, ngrp = (rpi + 3) / 4, heads = a.heads;
      const int k_off = l * 2 * d, v_off = l * 2 * d + d;
      for (int it = blockIdx.x; it < n_img * heads * ngrp; it += gridDim.x) {
        const int gq = it % ngrp, hh = (it / ngrp) % heads, b = it / (ngrp * heads);
        const int row0 = b * rpi + gq * 4, cnt = min(4, rpi - gq * 4);
        float* smf = reinterpret_cast<float*>(smem);
        switch (cnt) {
          case 1: cross_attn16_item<T, 1>(a.q, d, kv, a.ldkv, k_off, v_off, att, d, a.n_keys, a.n_valid, a.row_len, p, b, hh, row0, smf); break;
          case 2: cross_attn16_item<T, 2>(a.q, d, kv, a.ldkv, k_off, v_off, att, d, a.n_keys, a.n_valid, a.row_len, p, b, hh, row0, smf); break;
          case 3: cross_attn16_item<T, 3>(a.q, d, kv, a.ldkv, k_off, v_off, att, d, a.n_keys, a.n_valid, a.row_len, p, b, hh, row0, smf); break;
          default: cross_attn16_item<T, 4>(a.q, d, kv, a.ldkv, k_off, v_off, att, d, a.n_keys, a.n_valid, a.row_len, p, b, hh, row0, smf); break;
        }
        __syncthreads();
      }
    }
    bar.sync(a.dbg, &mark);
    // P5: x = x + att Wo^T + b
    {
      GemmP g{};
    g.dbg_mode = a.dbg_mode; g.fine = a.dbg; g.fmark = &fmark;
      g.dbg_mode = a.dbg_mode; g.fine = (a.dbg && l == 0) ? a.dbg : nullptr; g.fmark = &fmark;
      g.A16 = att; g.lda = d; g.W = W.w_wo; g.bias = W.b_wo; g.res = xout; g.ldr = ldc; g.Cf = xout; g.ldcf = ldc;
      g.M = R; g.N = d; g.K = d;
      gemm_phase<T, 32, false>(g, smem, mphase);
    }
    bar.sync(a.dbg, &mark);
    // P6: hid = relu(LN3(x) W1^T + b)
    {
      GemmP g{};
    g.dbg_mode = a.dbg_mode; g.fine = a.dbg; g.fmark = &fmark;
      g.dbg_mode = a.dbg_mode; g.fine = (a.dbg && l == 0) ? a.dbg : nullptr; g.fmark = &fmark;
      g.A32 = xout; g.lda32 = ldc; g.ln_g = W.n3g; g.ln_b = W.n3b;
      g.W = W.w_ff1; g.bias = W.b_ff1; g.Cb = hid; g.ldcb = a.ff; g.M = R; g.N = a.ff; g.K = d; g.act = 2;
      gemm_phase<T, 64, true>(g, smem, mphase);
    }
    bar.sync(a.dbg, &mark);
    // P7: x = x + hid W2^T + b   (also kept in 16 bits: operand of the reduce group)
    {
      GemmP g{};
    g.dbg_mode = a.dbg_mode; g.fine = a.dbg; g.fmark = &fmark;
      g.dbg_mode = a.dbg_mode; g.fine = (a.dbg && l == 0) ? a.dbg : nullptr; g.fmark = &fmark;
      g.A16 = hid; g.lda = a.ff; g.W = W.w_ff2; g.bias = W.b_ff2; g.res = xout; g.ldr = ldc; g.Cf = xout; g.ldcf = ldc;
      g.Cb = ycat16 + (size_t)l * d; g.ldcb = ldc; g.M = R; g.N = d; g.K = a.ff;
      g.ksplit = a.ksplit_ff2; g.scratch = a.scratch; g.cnt = a.bar + kSplitCntOff;
      gemm_phase<T, 64, false>(g, smem, mphase);
    }
    bar.sync(a.dbg, &mark);
  }
  // P8: reduce group: pre = x_last + [y_1 | .. | y_n] Wr^T + b   (End_ExpansionNet_v2.py:196-199)
  {
    GemmP g{};
    g.dbg_mode = a.dbg_mode; g.fine = a.dbg; g.fmark = &fmark;
    g.A16 = ycat16; g.lda = ldc; g.W = a.w_reduce; g.bias = a.b_reduce; g.res = a.ycat + (size_t)(nd - 1) * d; g.ldr = ldc;
    g.Cf = a.pre; g.ldcf = d; g.M = R; g.N = d; g.K = d * nd;
    g.ksplit = a.ksplit_red; g.scratch = a.scratch; g.cnt = a.bar + kSplitCntOff;
    gemm_phase<T, 64, false>(g, smem, mphase);
  }
  bar.sync(a.dbg, &mark);
  // P9: logits = LN(pre) Wv^T + b    (End_ExpansionNet_v2.py:200-204)
  {
    GemmP g{};
    g.dbg_mode = a.dbg_mode; g.fine = a.dbg; g.fmark = &fmark;
    g.A32 = a.pre; g.lda32 = d; g.ln_g = a.ng; g.ln_b = a.nb;
    g.W = a.w_vocab; g.bias = a.b_vocab; g.M = R; g.N = a.vocab; g.K = d;
    if (a.topk > 0) {
      const int ntn = (a.vocab + 63) / 64;
      TopkHook<64> hook{TopkParts{reinterpret_cast<float*>(a.parts), (long)R * ntn}, ntn, R, a.vocab, a.topk};
      gemm_phase<T, 64, true, false>(g, smem, mphase, hook);
      bar.sync(a.dbg, &mark);
      topk_merge_phase(TopkParts{reinterpret_cast<float*>(a.parts), (long)R * ntn}, ntn, R, a.topk, a.top_val, a.top_idx);
    } else {
      g.Cf = a.logits; g.ldcf = a.ldl;
      gemm_phase<T, 64, true>(g, smem, mphase);
    }
  }
  if (a.dbg && blockIdx.x == 0 && tid == 0) { a.dbg[mark++] = gtimer(); a.dbg[127] = (unsigned long long)mark; a.dbg[126] = (unsigned long long)fmark; }
}

size_t mega_scratch_bytes(int R) { return (size_t)((R + kBM - 1) / kBM) * (512 / 64) * 4 * (kBM * 64 * sizeof(float)); }
size_t mega_parts_bytes(int R, int vocab) { return (size_t)R * ((vocab + 63) / 64) * (2 + 2 * kTopC) * sizeof(float); }

bool mega_supported(const MegaArgs& a) {
  if (a.d != 512 || a.heads * 64 != a.d || a.n_keys > kCaMaxKeys || a.n_layers < 1 || a.n_layers > kMegaMaxLayers) return false;
  if (a.ff % 128 || a.vocab % 4 || a.s.P > 128 || a.n_exp > 64 || a.rows_per_image < 1 || a.R % a.rows_per_image) return false;
  if (((a.R + kBM - 1) / kBM) * (a.d / 64) > kMaxSplitTiles && (a.ksplit_ff2 > 1 || a.ksplit_red > 1)) return false;
  if (a.ksplit_ff2 > 4 || a.ksplit_red > 4 || a.ff != kKI * std::max(1, a.ksplit_ff2) || a.d * a.n_layers != kKI * std::max(1, a.ksplit_red)) return false;
  if (a.s.cw != 5 * a.d || (a.ldkv & 7) || (a.vocab + 63) / 64 > 256 || a.topk > kTopC) return false;
  return true;
}

template <typename T>
cudaError_t launch_dec_step_mega(const MegaArgs& a, cudaStream_t st) {
  if (!mega_supported(a)) return cudaErrorInvalidValue;
  struct DevInfo { int grid = 0; size_t smem = 0; };
  static DevInfo info[32];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 32) dev = 0;
  const size_t smem_gemm = kGemmSmem;
  const size_t smem_rows = dyn_exp_smem_floats(a.s.P, a.n_exp) * sizeof(float);
  const size_t smem = std::max(smem_gemm, smem_rows);
  if (info[dev].grid == 0 || info[dev].smem < smem) {
    if (cudaError_t e = cudaFuncSetAttribute(dec_step_mega_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) return e;
    int per_sm = 0, sms = 0;
    if (cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dec_step_mega_kernel<T>, kMegaThreads, smem)) return e;
    if (cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    info[dev].grid = std::min(per_sm, 2) * sms;
    info[dev].smem = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(info[dev].grid);
  cfg.blockDim = dim3(kMegaThreads);
  cfg.dynamicSmemBytes = info[dev].smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;          // all CTAs co-resident or the launch fails: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_mega_coop ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, dec_step_mega_kernel<T>, a);
}
template cudaError_t launch_dec_step_mega<f16>(const MegaArgs&, cudaStream_t);
template cudaError_t launch_dec_step_mega<bf16>(const MegaArgs&, cudaStream_t);

}  // namespace xn
