// Shared device/host helpers for libmurcl_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/murcl_b200.h"

namespace murcl {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MURCL_ECUDA;
  }
  return MURCL_OK;
}

#define MURCL_REQUIRE(cond, ...)         \
  do {                                   \
    if (!(cond)) {                       \
      ::murcl::set_error(__VA_ARGS__);   \
      return MURCL_EINVAL;               \
    }                                    \
  } while (0)

#define MURCL_CUDA(expr)                                                        \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      ::murcl::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));       \
      return MURCL_ECUDA;                                                       \
    }                                                                           \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }
int sm_count();
// murcl_set_row_order: 1 = the row-tile walk of the next launches (dense layers, fused pooling) starts at the LAST rows
int row_order_descending();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE attribute: one flag per device ordinal, so that a
// process driving several GPUs (nn.DataParallel-style) configures the kernel on each of them.
struct PerDeviceOnce {
  std::atomic<bool> done[64];
  PerDeviceOnce() { for (auto& d : done) d.store(false); }
  // returns the device slot to configure (>= 0) when the current device has not been configured yet, else -1
  int pending() const {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;   // unknown ordinal: always (re)configure
    return done[dev].load(std::memory_order_acquire) ? -1 : dev;
  }
  void mark(int dev) { if (dev >= 0 && dev < 64) done[dev].store(true, std::memory_order_release); }
};

// ---- storage type helpers -----------------------------------------------------------------
template <typename T>
struct Store;
template <>
struct Store<float> {
  static __device__ __forceinline__ float load(const float* p) { return *p; }
  static __device__ __forceinline__ void store(float* p, float v) { *p = v; }
};
template <>
struct Store<__nv_bfloat16> {
  static __device__ __forceinline__ float load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// 4 consecutive elements as fp32 (16 B fp32 / 8 B bf16 vector access; pointer must be aligned).
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
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

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) v = warp_sum(v);
  if (threadIdx.x == 0) red[0] = v;
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  v = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) v = warp_max(v);
  if (threadIdx.x == 0) red[0] = v;
  __syncthreads();
  return red[0];
}

// Activation selector used by GEMM epilogues.  tanhf/expf are the accurate libm versions: the
// fp32 path has a 1e-5 relative budget against the reference.
__device__ __forceinline__ float apply_act(float v, int act, int col, int n_cols) {
  switch (act) {
    case MURCL_ACT_RELU: return fmaxf(v, 0.f);
    case MURCL_ACT_TANH: return tanhf(v);
    case MURCL_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case MURCL_ACT_TANH_SIGMOID: return (col < (n_cols >> 1)) ? tanhf(v) : 1.f / (1.f + expf(-v));
    default: return v;
  }
}

}  // namespace murcl
