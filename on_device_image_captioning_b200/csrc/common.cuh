// Shared device helpers for the xnv2_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <utility>

namespace xn {

constexpr int kWin = 12;            // Swin window edge (swin_window_size=12 at every reference call site)
constexpr int kWinTok = kWin * kWin; // 144 tokens per window
constexpr int kHeadDim = 32;        // C / heads == 32 in every Swin-L stage (reference swin:189-190)

// ---- programmatic dependent launch (PDL).  Every kernel of the 16-bit production path is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and starts with pdl_wait() (griddepcontrol.wait: blocks until the
// preceding kernels in the stream have completed and flushed) followed by pdl_trigger() (lets the next kernel's CTAs
// be scheduled as soon as every CTA of this one is running).  Launch latency, block scheduling and the per-kernel
// prologue then overlap the tail of the previous kernel instead of following it -- the decode loop is ~700 dependent
// launches of a few microseconds each.  Nothing before pdl_wait() may touch global memory written by earlier kernels.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: remember the largest size configured
// on each device (call sites keep one `static DynSmemState` per kernel instantiation)
struct DynSmemState { size_t bytes[32] = {}; };
template <typename F>
inline cudaError_t ensure_dyn_smem(F func, size_t bytes, DynSmemState& st) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 32) dev = 0;
  if (bytes <= st.bytes[dev] || bytes <= 48 * 1024) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) st.bytes[dev] = bytes;
  return e;
}

extern int g_pdl_enabled;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); `red` is >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = (l < nw) ? red[l] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = (l < nw) ? red[l] : -INFINITY;
  r = warp_max(r);
  return r;
}

__device__ __forceinline__ float gelu_erf(float x) {   // nn.GELU() default (exact erf), swin:97-120
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

}  // namespace xn
