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

// Programmatic dependent launch: the kernel may start while its predecessor in the stream is still draining; it must execute
// pdl_wait() before touching global memory.  Hides launch latency and the prologue (barrier init, TMEM allocation, descriptor
// prefetch) of the ~170 back-to-back kernels of a ViT forward.  BSCNAV_PDL=0 launches them fully serialised (A/B, debugging).
bool pdl_enabled();
#define BSC_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                                   \
  do {                                                                                           \
    cudaLaunchConfig_t _cfg = {};                                                                \
    _cfg.gridDim = (grid); _cfg.blockDim = (block); _cfg.dynamicSmemBytes = (smem); _cfg.stream = (stream); \
    cudaLaunchAttribute _at[1];                                                                  \
    _at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                              \
    _at[0].val.programmaticStreamSerializationAllowed = ::bscnav::pdl_enabled() ? 1 : 0;         \
    _cfg.attrs = _at; _cfg.numAttrs = 1;                                                         \
    cudaError_t _e = cudaLaunchKernelEx(&_cfg, kernel, __VA_ARGS__);                             \
    ::bscnav::g_launches.fetch_add(1, std::memory_order_relaxed);                                \
    if (_e != cudaSuccess) return ::bscnav::cuda_fail(_e, #kernel);                              \
  } while (0)

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
