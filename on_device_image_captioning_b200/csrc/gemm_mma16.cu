// Batched 16-bit tensor-core GEMM (mma.sync.m16n8k16, fp32 accumulate) for the per-image contractions of the
// static-expansion block (reference models/layers.py:52,62-63,82-83):
//     z[b]     = Q . key[b]^T / sqrt(d)            (992 x 512) . (512 x 144)
//     class[b] = fw[b] . A[b] + bias_exp           (992 x 144) . (144 x 512)
//     out[b]   = bw[b] . class[b] / n_groups       (144 x 992) . (992 x 512)
// These are block-diagonal over images (a different right-hand operand per image) with 144-wide dimensions and, for
// two of the three, a row-major (K x N) right operand -- shapes that do not fit the TMA/tcgen05 linear-layer kernel
// (gemm_tcgen05.cu: one weight matrix, both operands K-major, 128-row UMMA tiles).  They are 2.6 % of the path's
// FLOPs; this kernel exists so that they stop costing 10 % of its time on the fp32 CUDA-core path.
//
//   C[b] (M x N) = scale * A[b] (M x K, K contiguous) . op(B[b]) + res[b],   B (N x K) K-contiguous or (K x N) N-contiguous
// Tiles BM x 128 x 32, 8 warps, 3-stage cp.async ring, ldmatrix(.trans) fragments.
#include <cuda_fp16.h>
#include "kernels.h"
#include "common.cuh"

namespace xn {

namespace {

constexpr int kBN = 128, kBKm = 32, kMmaStages = 3;
constexpr int kApad = kBKm + 8;        // A / B(NK) smem row: 40 halves = 80 B  (conflict-free ldmatrix)
constexpr int kBpadKN = kBN + 8;       // B(KN) smem row: 136 halves = 272 B

template <typename T> struct MmaOp;
template <> struct MmaOp<bf16> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
template <> struct MmaOp<f16> {
  static __device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// 16-byte async copy; src_bytes == 0 zero-fills the destination (out-of-range rows / K tail)
__device__ __forceinline__ void cp16z(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <typename T> __device__ __forceinline__ void store2(T* p, float a, float b);
template <> __device__ __forceinline__ void store2<float>(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
template <> __device__ __forceinline__ void store2<bf16>(bf16* p, float a, float b) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b); }
template <> __device__ __forceinline__ void store2<f16>(f16* p, float a, float b) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(a, b); }

}  // namespace

// BM = 128: warps 2 (M) x 4 (N), warp tile 64 x 32.   BM = 64: warps 2 x 4, warp tile 32 x 32.
template <typename T, typename OutT, int BM, bool BKN>
__global__ void __launch_bounds__(256) gemm_mma16_kernel(Mma16Args p) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int kAStage = BM * kApad;
  constexpr int kBStage = BKN ? kBKm * kBpadKN : kBN * kApad;
  T* As = reinterpret_cast<T*>(smraw);
  T* Bs = As + kMmaStages * kAStage;
  const uint32_t as_u = (uint32_t)__cvta_generic_to_shared(As), bs_u = (uint32_t)__cvta_generic_to_shared(Bs);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;                    // warp grid 2 x 4
  constexpr int WM = BM / 2, MT = WM / 16;                    // warp rows, m16 tiles per warp
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * kBN, bz = blockIdx.z;
  const T* A = reinterpret_cast<const T*>(p.A) + (long)bz * p.sA;
  const T* B = reinterpret_cast<const T*>(p.B) + (long)bz * p.sB;

  float acc[MT][4][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

  const int nk = (p.K + kBKm - 1) / kBKm;
  auto load_stage = [&](int kt, int st) {
    const int k0 = kt * kBKm;
    // A tile: BM rows x 4 chunks of 8 halves
    for (int i = tid; i < BM * 4; i += 256) {
      const int r = i >> 2, ch = (i & 3) * 8;
      const bool ok = (m0 + r < p.M) && (k0 + ch < p.K);
      cp16z(as_u + (uint32_t)((st * kAStage + r * kApad + ch) * 2), ok ? (const void*)(A + (long)(m0 + r) * p.lda + k0 + ch) : (const void*)A, ok ? 16 : 0);
    }
    if (!BKN) {
      for (int i = tid; i < kBN * 4; i += 256) {
        const int r = i >> 2, ch = (i & 3) * 8;
        const bool ok = (n0 + r < p.N) && (k0 + ch < p.K);
        cp16z(bs_u + (uint32_t)((st * kBStage + r * kApad + ch) * 2), ok ? (const void*)(B + (long)(n0 + r) * p.ldb + k0 + ch) : (const void*)B, ok ? 16 : 0);
      }
    } else {
      for (int i = tid; i < kBKm * (kBN / 8); i += 256) {
        const int r = i / (kBN / 8), ch = (i % (kBN / 8)) * 8;
        const bool ok = (k0 + r < p.K) && (n0 + ch < p.N);
        cp16z(bs_u + (uint32_t)((st * kBStage + r * kBpadKN + ch) * 2), ok ? (const void*)(B + (long)(k0 + r) * p.ldb + n0 + ch) : (const void*)B, ok ? 16 : 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  for (int s = 0; s < kMmaStages - 1; ++s) {
    if (s < nk) load_stage(s, s);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kt = 0; kt < nk; ++kt) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kMmaStages - 2) : "memory");
    __syncthreads();
    if (kt + kMmaStages - 1 < nk) load_stage(kt + kMmaStages - 1, (kt + kMmaStages - 1) % kMmaStages);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    const int st = kt % kMmaStages;
#pragma unroll
    for (int ks = 0; ks < kBKm / 16; ++ks) {
      uint32_t af[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int row = wm * WM + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int col = ks * 16 + (lane >> 4) * 8;
        ldsm4(af[mt], as_u + (uint32_t)((st * kAStage + row * kApad + col) * 2));
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {                          // pairs of n8 tiles
        uint32_t bq[4];
        if (!BKN) {
          const int nrow = wn * 32 + np * 16 + ((lane >> 4) & 1) * 8 + (lane & 7);
          const int kcol = ks * 16 + ((lane >> 3) & 1) * 8;
          ldsm4(bq, bs_u + (uint32_t)((st * kBStage + nrow * kApad + kcol) * 2));
        } else {
          const int krow = ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
          const int ncol = wn * 32 + np * 16 + ((lane >> 4) & 1) * 8;
          ldsm4t(bq, bs_u + (uint32_t)((st * kBStage + krow * kBpadKN + ncol) * 2));
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          MmaOp<T>::mma(acc[mt][np * 2 + 0], af[mt], bq[0], bq[1]);
          MmaOp<T>::mma(acc[mt][np * 2 + 1], af[mt], bq[2], bq[3]);
        }
      }
    }
  }

  OutT* C = reinterpret_cast<OutT*>(p.C) + (long)bz * p.sC;
  const float* R = p.res ? p.res + (long)bz * p.sR : nullptr;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = n0 + wn * 32 + nt * 8 + (lane & 3) * 2;
      if (col >= p.N) continue;
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int row = m0 + wm * WM + mt * 16 + (lane >> 2) + hrow * 8;
        if (row >= p.M) continue;
        float v0 = acc[mt][nt][hrow * 2] * p.scale, v1 = acc[mt][nt][hrow * 2 + 1] * p.scale;
        if (R) { v0 += R[(long)row * p.ldr + col]; v1 += R[(long)row * p.ldr + col + 1]; }
        store2<OutT>(C + (long)row * p.ldc + col, v0, v1);
      }
    }
}

template <typename T, typename OutT>
cudaError_t launch_gemm_mma16(const Mma16Args& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return cudaSuccess;
  if ((p.K & 7) || (p.lda & 7) || (p.ldb & 7) || (p.N & 1) || (p.ldc & 1) || (p.b_kn && (p.N & 7))) return cudaErrorInvalidValue;
  const bool small_m = p.M <= 192;
  const int BM = small_m ? 64 : 128;
  const size_t a_st = (size_t)BM * kApad, b_st = p.b_kn ? (size_t)kBKm * kBpadKN : (size_t)kBN * kApad;
  const size_t smem = kMmaStages * (a_st + b_st) * 2;
  dim3 grid((p.N + kBN - 1) / kBN, (p.M + BM - 1) / BM, p.batch);
#define XN_LAUNCH_MMA(BMV, KNV)                                                                                         \
  do {                                                                                                                   \
    static DynSmemState smem_state;                                                                                      \
    if (cudaError_t e = ensure_dyn_smem(gemm_mma16_kernel<T, OutT, BMV, KNV>, smem, smem_state)) return e;               \
    launch_k(gemm_mma16_kernel<T, OutT, BMV, KNV>, dim3(grid), dim3(256), smem, st, p);                                                   \
  } while (0)
  if (small_m) { if (p.b_kn) XN_LAUNCH_MMA(64, true); else XN_LAUNCH_MMA(64, false); }
  else         { if (p.b_kn) XN_LAUNCH_MMA(128, true); else XN_LAUNCH_MMA(128, false); }
#undef XN_LAUNCH_MMA
  return cudaGetLastError();
}
template cudaError_t launch_gemm_mma16<bf16, float>(const Mma16Args&, cudaStream_t);
template cudaError_t launch_gemm_mma16<bf16, bf16>(const Mma16Args&, cudaStream_t);
template cudaError_t launch_gemm_mma16<f16, float>(const Mma16Args&, cudaStream_t);
template cudaError_t launch_gemm_mma16<f16, f16>(const Mma16Args&, cudaStream_t);

}  // namespace xn
