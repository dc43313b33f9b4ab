// Shared host/device helpers for the bscnav_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/bscnav_b200.h"

namespace bscnav {

extern thread_local char g_last_error[256];
extern std::atomic<long long> g_launches;

inline int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
  return BSCNAV_ECUDA;
}

#define BSC_CUDA(call)                                                     \
  do {                                                                     \
    cudaError_t _e = (call);                                               \
    if (_e != cudaSuccess) return ::bscnav::cuda_fail(_e, #call);          \
  } while (0)

// every kernel launch goes through this so bench.py can report gpu_launches
#define BSC_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);            \
    ::bscnav::g_launches.fetch_add(1, std::memory_order_relaxed);          \
    cudaError_t _e = cudaGetLastError();                                   \
    if (_e != cudaSuccess) return ::bscnav::cuda_fail(_e, #kernel);        \
  } while (0)

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

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

// streaming 128-bit load that does not allocate in L1
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

}  // namespace bscnav
