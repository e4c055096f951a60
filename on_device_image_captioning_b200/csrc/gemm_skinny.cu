// Latency-oriented GEMM for the decoder-step linears (reference models/layers.py:222-248, 266-308: the
// DynamicExpansion projections, MHA Wq / out_linear, FeedForward, dec_reduce_group), where the row count is
// images x beam (a few hundred) and every launch is bounded by how quickly its operands arrive from L2, not by
// tensor throughput:
//
//   C (M x N) = act( LN?(A) (M x K) . W^T (N x K) + bias ) + res
//
// * tiles of 64 rows x BN columns (BN = 32 / 64), 8 warps (4 along M x 2 along N), mma.sync.m16n8k16 with fp32 accumulation;
//   the point of the small tiles is the CTA count (48 .. 240 CTAs for M = 192): every SM pulls < 200 KB
// * both operands of a CTA are fully resident in shared memory and ALL their loads are issued up front (cp.async in
//   four K-quarter groups; compute starts when the first quarter has landed)
// * A may be fp32 (K slice <= 512) with an optional LayerNorm over the row (gamma, beta; K == row width): each warp
//   converts / normalises 8 rows on the way into shared memory with all of their loads in flight at once
//
// Measured in-graph on B200 (tools/dec_gemm_graph_bench.py, profiles/): for the few-row case (batch-1 latency, M <= 64)
// the N = 512 projections run in 6-8 us here against 7-13.5 us on the 128-row tcgen05 tiles; at M = 192 the 16-byte
// cp.async streams of 100+ KB per CTA are slower than TMA (15 us vs 7 us), so the engine selects this kernel only for
// M <= 64 and N <= 2048 and keeps the tcgen05 kernel otherwise.
// * split-K across a thread-block cluster (grid.z = cluster size 2..4): each CTA contracts one K slice, the partial
//   tiles are pushed into the leader's shared memory over DSMEM and summed there in rank order (deterministic)
#include <cuda_fp16.h>
#include "kernels.h"
#include "common.cuh"

namespace xn {

namespace {

constexpr int kSkBM = 64;
constexpr int kSkThreads = 256;
constexpr int kSkGroups = 4;

template <typename T> struct SkMma;
template <> struct SkMma<bf16> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};
template <> struct SkMma<f16> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

__device__ __forceinline__ void sk_ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void sk_cp16z(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void sk_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void sk_wait(int pending) {
  switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
  }
}
__device__ __forceinline__ uint32_t sk_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void sk_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void sk_st_remote_f2(uint32_t local_addr, uint32_t rank, float a, float b) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(ra), "f"(a), "f"(b) : "memory");
}

}  // namespace

template <typename T, int VPL, int RIF>
__device__ __forceinline__ void skinny_ln_rows(const SkinnyArgs& p, T* As, int PA, int m0, int k0, int Kc, int vpl, int warp, int lane) {
  const float inv_k = 1.0f / (float)Kc;
#pragma unroll 1
  for (int rb = 0; rb < 8; rb += RIF) {
    float4 v[RIF][VPL];
#pragma unroll
    for (int q = 0; q < RIF; ++q) {
      const int row = m0 + warp * 8 + rb + q;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        v[q][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < vpl && row < p.M) v[q][i] = *reinterpret_cast<const float4*>(p.A32 + (long)row * p.lda + k0 + (i * 32 + lane) * 4);
      }
    }
#pragma unroll
    for (int q = 0; q < RIF; ++q) {
      float mean = 0.f, rstd = 1.f;
      if (p.ln_g) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) s += (v[q][i].x + v[q][i].y) + (v[q][i].z + v[q][i].w);   // unused slots hold zeros
        mean = warp_sum(s) * inv_k;
        float qq = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          if (i < vpl) {
            const float d0 = v[q][i].x - mean, d1 = v[q][i].y - mean, d2 = v[q][i].z - mean, d3 = v[q][i].w - mean;
            qq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
          }
        rstd = 1.0f / sqrtf(warp_sum(qq) * inv_k + 1e-5f);
      }
      T* dst = As + (warp * 8 + rb + q) * PA;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if (i < vpl) {
          const int c = (i * 32 + lane) * 4;
          float4 o = v[q][i];
          if (p.ln_g) {
            const float4 ga = *reinterpret_cast<const float4*>(p.ln_g + c), be = *reinterpret_cast<const float4*>(p.ln_b + c);
            o.x = (o.x - mean) * rstd * ga.x + be.x; o.y = (o.y - mean) * rstd * ga.y + be.y;
            o.z = (o.z - mean) * rstd * ga.z + be.z; o.w = (o.w - mean) * rstd * ga.w + be.w;
          }
          uint2 u;
          u.x = SkMma<T>::pack(o.x, o.y);
          u.y = SkMma<T>::pack(o.z, o.w);
          *reinterpret_cast<uint2*>(dst + c) = u;
        }
    }
  }
}

// grid = (ceil(N / BN), ceil(M / 64), split); cluster = (1, 1, split).  Kc = K / split columns per CTA.
template <typename T, int BN>
__global__ void __launch_bounds__(kSkThreads) gemm_skinny_kernel(SkinnyArgs p, int Kc, int split) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char smraw[];
  const int PA = Kc + 8;                               // smem row pitch in elements: (2 Kc + 16) / 16 is odd -> conflict-free ldmatrix
  T* As = reinterpret_cast<T*>(smraw);                 // [64][PA]
  T* Ws = As + kSkBM * PA;                             // [BN][PA]
  float* part = reinterpret_cast<float*>(Ws + BN * PA);   // [split - 1][64][BN]   (leader only)
  const uint32_t as_u = (uint32_t)__cvta_generic_to_shared(As), ws_u = (uint32_t)__cvta_generic_to_shared(Ws);
  const uint32_t part_u = (uint32_t)__cvta_generic_to_shared(part);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * kSkBM;
  const uint32_t rank = split > 1 ? sk_cluster_rank() : 0u;
  const int k0 = (int)rank * Kc;
  const int cpr = Kc / 8;                              // 16-byte chunks per row
  const int gq = cpr / kSkGroups;                      // chunks per row per group (Kc % 64 == 0 -> whole k16 steps per group)

  // ---- issue every load of this CTA: K-quarter g of W (and of a 16-bit A) forms cp.async group g
  const T* W = reinterpret_cast<const T*>(p.W);
  const T* A16 = reinterpret_cast<const T*>(p.A16);
  for (int g = 0; g < kSkGroups; ++g) {
    for (int i = tid; i < BN * gq; i += kSkThreads) {
      const int r = i / gq, ch = g * gq + i % gq;
      const bool ok = n0 + r < p.N;
      sk_cp16z(ws_u + (uint32_t)((r * PA + ch * 8) * 2), ok ? (const void*)(W + (long)(n0 + r) * p.ldw + k0 + ch * 8) : (const void*)W, ok ? 16 : 0);
    }
    if (A16) {
      for (int i = tid; i < kSkBM * gq; i += kSkThreads) {
        const int r = i / gq, ch = g * gq + i % gq;
        const bool ok = m0 + r < p.M;
        sk_cp16z(as_u + (uint32_t)((r * PA + ch * 8) * 2), ok ? (const void*)(A16 + (long)(m0 + r) * p.lda + k0 + ch * 8) : (const void*)A16, ok ? 16 : 0);
      }
    }
    sk_commit();
  }

  // ---- fp32 A: (LayerNorm and) conversion; warp w takes rows 8w .. 8w+7, every load of a batch of rows issued before
  // the first reduction (32 float4 registers per lane: 8 rows of a K slice <= 512)
  if (!A16) {
    skinny_ln_rows<T, 4, 8>(p, As, PA, m0, k0, Kc, Kc / 128, warp, lane);     // <= 4 float4 per lane per row (K slice <= 512)
  }

  // ---- contraction: warp (wm, wn) owns rows 16 wm .. +15 and columns wn * BN/2 .. + BN/2 - 1
  constexpr int NT = BN / 16;
  const int wm = warp >> 1, wn = warp & 1;
  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  const uint32_t a_addr = as_u + (uint32_t)(((wm * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PA + (lane >> 4) * 8) * 2);
  const uint32_t w_addr = ws_u + (uint32_t)(((wn * (BN / 2) + ((lane >> 4) & 1) * 8 + (lane & 7)) * PA + ((lane >> 3) & 1) * 8) * 2);
  const int steps_per_group = Kc / 16 / kSkGroups;
  for (int g = 0; g < kSkGroups; ++g) {
    sk_wait(kSkGroups - 1 - g);
    __syncthreads();
#pragma unroll 4
    for (int s = 0; s < steps_per_group; ++s) {
      const int ks = g * steps_per_group + s;
      uint32_t af[4];
      sk_ldsm4(af, a_addr + ks * 32);
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t bq[4];
        sk_ldsm4(bq, w_addr + (uint32_t)(np * 16 * PA * 2) + ks * 32);
        SkMma<T>::mma(acc[2 * np], af, bq[0], bq[1]);
        SkMma<T>::mma(acc[2 * np + 1], af, bq[2], bq[3]);
      }
    }
  }

  // ---- split-K: partial tiles to the leader over DSMEM, summed in rank order
  const int rl0 = wm * 16 + (lane >> 2);               // local rows rl0, rl0 + 8; local cols wn*BN/2 + nt*8 + (lane&3)*2
  if (split > 1) {
    if (rank != 0) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int cl = wn * (BN / 2) + nt * 8 + (lane & 3) * 2;
        sk_st_remote_f2(part_u + (uint32_t)((((rank - 1) * kSkBM + rl0) * BN + cl) * 4), 0u, acc[nt][0], acc[nt][1]);
        sk_st_remote_f2(part_u + (uint32_t)((((rank - 1) * kSkBM + rl0 + 8) * BN + cl) * 4), 0u, acc[nt][2], acc[nt][3]);
      }
    }
    sk_cluster_sync();
    if (rank != 0) return;
    for (int r = 1; r < split; ++r)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int cl = wn * (BN / 2) + nt * 8 + (lane & 3) * 2;
        const float2 a = *reinterpret_cast<const float2*>(part + ((r - 1) * kSkBM + rl0) * BN + cl);
        const float2 b = *reinterpret_cast<const float2*>(part + ((r - 1) * kSkBM + rl0 + 8) * BN + cl);
        acc[nt][0] += a.x; acc[nt][1] += a.y; acc[nt][2] += b.x; acc[nt][3] += b.y;
      }
  }

  // ---- epilogue: bias, activation, fp32 residual, fp32 or 16-bit store
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int col = n0 + wn * (BN / 2) + nt * 8 + (lane & 3) * 2;
    if (col >= p.N) continue;
    float b0 = 0.f, b1 = 0.f;
    if (p.bias) { b0 = p.bias[col]; b1 = col + 1 < p.N ? p.bias[col + 1] : 0.f; }
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int row = m0 + rl0 + hr * 8;
      if (row >= p.M) continue;
      float v0 = acc[nt][hr * 2] + b0, v1 = acc[nt][hr * 2 + 1] + b1;
      if (p.act == 2) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
      else if (p.act == 1) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
      if (p.res) {
        const float2 rv = *reinterpret_cast<const float2*>(p.res + (long)row * p.ldr + col);
        v0 += rv.x; v1 += rv.y;
      }
      if (p.Cf) *reinterpret_cast<float2*>(p.Cf + (long)row * p.ldc + col) = make_float2(v0, v1);
      else *reinterpret_cast<uint32_t*>(reinterpret_cast<T*>(p.Cb) + (long)row * p.ldc + col) = SkMma<T>::pack(v0, v1);
    }
  }
}

bool skinny_gemm_supported(const SkinnyArgs& p) {
  if (p.M <= 0 || p.M > 512 || p.N <= 0 || (p.N & 1) || p.K <= 0 || (p.K % 64)) return false;
  if ((p.ldw & 7) || (p.ldc & 1) || (p.res && (p.ldr & 1))) return false;
  if ((p.A16 != nullptr) == (p.A32 != nullptr) || (p.Cf != nullptr) == (p.Cb != nullptr)) return false;
  if (p.A16 && ((p.lda & 7) || (reinterpret_cast<uintptr_t>(p.A16) & 15))) return false;
  if (p.A32 && ((p.lda & 3) || (reinterpret_cast<uintptr_t>(p.A32) & 15))) return false;
  if (reinterpret_cast<uintptr_t>(p.W) & 15) return false;
  if (p.ln_g && (!p.A32 || p.K > 512 || (p.K % 128))) return false;
  if (p.res && (reinterpret_cast<uintptr_t>(p.res) & 7)) return false;
  if (p.Cf && (reinterpret_cast<uintptr_t>(p.Cf) & 7)) return false;
  return true;
}

template <typename T, int BN>
static cudaError_t launch_skinny_bn(const SkinnyArgs& p, int split, cudaStream_t st) {
  const int Kc = p.K / split;
  const size_t smem = (size_t)(kSkBM + BN) * (Kc + 8) * 2 + (split > 1 ? (size_t)(split - 1) * kSkBM * BN * 4 : 0);
  static DynSmemState smem_state;
  if (cudaError_t e = ensure_dyn_smem(gemm_skinny_kernel<T, BN>, smem, smem_state)) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((p.N + BN - 1) / BN, (p.M + kSkBM - 1) / kSkBM, split);
  cfg.blockDim = dim3(kSkThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (split > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = split;
    ++na;
  }
  if (g_pdl_enabled) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<T, BN>, p, Kc, split);
}

template <typename T>
cudaError_t launch_gemm_skinny(const SkinnyArgs& p, cudaStream_t st) {
  if (!skinny_gemm_supported(p)) return cudaErrorInvalidValue;
  const int m_tiles = (p.M + kSkBM - 1) / kSkBM;
  // wide outputs take 64-column tiles; narrow ones 32-column tiles and, when K is long and there is no LayerNorm over
  // the row, a K split across a cluster so that ~100+ CTAs share the operand traffic
  const int bn = (long)m_tiles * ((p.N + 63) / 64) >= 96 ? 64 : 32;
  int split = 1;
  if (!p.ln_g) {
    const long ctas = (long)m_tiles * ((p.N + bn - 1) / bn);
    for (int s = 2; s <= 4; ++s) {
      if (ctas * split >= 128 && !(p.A32 && p.K / split > 512)) break;
      if (p.K % (64 * s) == 0 && p.K / s >= 256 && (!p.A32 || (p.K / s) % 128 == 0)) split = s;
    }
  }
  if (p.A32 && ((p.K / split) % 128 || p.K / split > 512)) return cudaErrorInvalidValue;
  return bn == 64 ? launch_skinny_bn<T, 64>(p, split, st) : launch_skinny_bn<T, 32>(p, split, st);
}
template cudaError_t launch_gemm_skinny<bf16>(const SkinnyArgs&, cudaStream_t);
template cudaError_t launch_gemm_skinny<f16>(const SkinnyArgs&, cudaStream_t);

}  // namespace xn
