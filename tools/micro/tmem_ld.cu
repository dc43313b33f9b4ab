// Micro-benchmark: cost of tcgen05.ld round trips (32x32b shapes) as the softmax warps of k_attention_ws2 use them:
// cycles per (load(s) + tcgen05.wait::ld) for x16 / x32 / x64 / x128 columns per load, 1 or 2 loads per wait,
// with 1, 2 (same scheduler), 4 and 8 warps reading concurrently.  Answers: is a chunk's ~300 cycles in the row-max pass a
// fixed round-trip cost (then wider loads help) or TMEM read bandwidth (then nothing does)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../bsc-nav_b200/csrc tmem_ld.cu -o tmem_ld
#include <cstdio>
#include <cuda.h>
#include "tc_ptx.cuh"
using namespace bscnav;
namespace bscnav { thread_local char g_last_error[256]; std::atomic<long long> g_launches{0}; bool pdl_enabled() { return false; } }

template <int X>
__device__ __forceinline__ void ld_x(uint32_t taddr, uint32_t* r);
template <> __device__ __forceinline__ void ld_x<16>(uint32_t t, uint32_t* r) { tmem_ld_32x16(t, r); }
template <> __device__ __forceinline__ void ld_x<32>(uint32_t t, uint32_t* r) { tmem_ld_32x32(t, r); }
template <> __device__ __forceinline__ void ld_x<64>(uint32_t t, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, "
      "%29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, "
      "%56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]),
        "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]),
        "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(t));
}

// MODE 0: one load per wait; MODE 1: two loads per wait
// mma: 0 none, 1 = a ninth warp issues tcgen05.mma (A and B in shared memory, N = 128) back to back while the others load,
// 2 = the same with A in tensor memory and N = 64 (the O = P V shape of the attention kernel)
template <int X, int MODE>
__global__ void __launch_bounds__(288, 1) k_ld(int iters, int mask, long long* out, float* sink, int mma) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned long long bar;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) { stop = 0; mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // warps w and w + 4 share a scheduler and a TMEM lane quarter (as softmax groups a / b do)
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  float acc = 0.f;
  if ((mask >> warp) & 1) {
    uint32_t va[X], vb[MODE ? X : 1];
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t c = (uint32_t)((i * (MODE ? 2 : 1) * X) & 127);
      ld_x<X>(lane_addr + c, va);
      if (MODE) ld_x<X>(lane_addr + ((c + X) & 127), vb);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < X; j += 8) acc += __uint_as_float(va[j]);
      if (MODE) {
#pragma unroll
        for (int j = 0; j < X; j += 8) acc += __uint_as_float(vb[j]);
      }
    }
    const long long t1 = clock64();
    if (lane == 0) out[warp] = t1 - t0;
    if (warp == 0 && lane == 0) stop = 1;
  } else if (warp == 8 && lane == 0 && mma) {
    const uint32_t sbase = smem_u32(smem);
    const uint32_t idesc = umma_idesc_f16(mma == 2 ? 64 : 128, 128, mma == 2 ? 1 : 0);
    const uint64_t a = umma_desc_sw128(sbase), b = umma_desc_sw128(sbase + 32768);
    long long n = 0;
    while (!stop) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        // accumulators in columns [384, 512): outside the columns the loading warps read ([0, 256 + 128))
        if (mma == 2) {
          asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }"
                       ::"r"(tmem + 448), "r"(tmem + 384 + 8 * (k & 3)), "l"(b + 128 * k), "r"(idesc), "r"(1u) : "memory");
        } else {
          tcgen05_mma_f16(tmem + 384, a + 2 * (k & 3), b + 2 * (k & 3), idesc, 1u);
        }
      }
      n += 8;
    }
    tcgen05_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    out[8] = n;
  }
  if (acc == 123.456f) sink[0] = acc;
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

template <int X, int MODE>
static void run(const char* name, long long* d_out, float* sink, int mma = 0) {
  const int iters = 2000;
  cudaFuncSetAttribute(k_ld<X, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int warps : {1, 2, 4, 8}) {
    const int mask = warps == 1 ? 0x01 : warps == 2 ? 0x11 : warps == 4 ? 0x0F : 0xFF;   // 2: warps 0 and 4 share a scheduler
    cudaMemset(d_out, 0, 9 * sizeof(long long));
    k_ld<X, MODE><<<1, 288, 100 * 1024>>>(iters, mask, d_out, sink, mma);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[9];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 8; ++i) mx = h[i] > mx ? h[i] : mx;
    const double cyc = (double)mx / iters;
    const double bytes = (double)(MODE ? 2 : 1) * X * 4 * 32;
    printf("%-28s warps %d: %7.1f cycles per wait, %6.1f B/cycle/warp", name, warps, cyc, bytes / cyc);
    if (mma) printf("   | concurrent MMAs (%s): %lld, one per %.1f cycles", mma == 2 ? "A in TMEM, N=64" : "A in smem, N=128", h[8], (double)mx / (double)(h[8] ? h[8] : 1));
    printf("  (%s)\n", cudaGetErrorString(e));
  }
}

int main() {
  long long* d_out; float* sink;
  cudaMalloc(&d_out, 128); cudaMalloc(&sink, 4);
  run<16, 0>("x16, one load per wait", d_out, sink);
  run<32, 0>("x32, one load per wait", d_out, sink);
  run<32, 1>("x32, two loads per wait", d_out, sink);
  run<64, 0>("x64, one load per wait", d_out, sink);
  run<64, 1>("x64, two loads per wait", d_out, sink);
  printf("---- with a ninth warp issuing tcgen05.mma continuously\n");
  run<32, 0>("x32, one load per wait", d_out, sink, 1);
  run<32, 0>("x32, one load per wait", d_out, sink, 2);
  run<32, 1>("x32, two loads per wait", d_out, sink, 2);
  run<64, 0>("x64, one load per wait", d_out, sink, 2);
  return 0;
}
