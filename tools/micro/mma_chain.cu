// Micro-benchmark: issue rate of small tcgen05.mma instructions (M=128, K=16) as a function of N and of how many independent TMEM
// accumulators the back-to-back instructions rotate over.  Answers: is the ~110 cycles per P V step seen in k_attention_ws2 a
// fixed issue cost, or the latency of accumulating into the SAME accumulator?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../bsc-nav_b200/csrc mma_chain.cu -o mma_chain
#include <cstdio>
#include <cuda.h>
#include "tc_ptx.cuh"
using namespace bscnav;
namespace bscnav { thread_local char g_last_error[256]; std::atomic<long long> g_launches{0}; bool pdl_enabled() { return false; } }

__global__ void __launch_bounds__(128, 1) k_chain(int n_mma, int N, int n_acc, int ts, long long* out, int issuers) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ unsigned long long bar, bar2;
  const uint32_t sbase = smem_u32(smem);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;   // fp16 1.0
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < issuers) {
    const int me = threadIdx.x >> 5;
    n_mma /= issuers;
    const uint32_t idesc = umma_idesc_f16(N, 128, ts ? 1 : 0);
    const uint64_t a = umma_desc_sw128(sbase), b = umma_desc_sw128(sbase + 32768);
    // operands precomputed, eight instructions per loop trip: nothing but the MMAs between the clock reads
    uint32_t d8[8], at8[8];
    uint64_t a8[8], b8[8];
    for (int k = 0; k < 8; ++k) {
      d8[k] = tmem + 256 + (uint32_t)(((k % n_acc) + me * n_acc) * (N > 64 ? N : 64));
      at8[k] = tmem + (uint32_t)(8 * k);
      a8[k] = a + 2 * (k & 3);
      b8[k] = ts ? b + 128 * k : b + 2 * (k & 3);
    }
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (ts) {
          asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }"
                       ::"r"(d8[k]), "r"(at8[k]), "l"(b8[k]), "r"(idesc), "r"(1u) : "memory");
        } else {
          tcgen05_mma_f16(d8[k], a8[k], b8[k], idesc, 1u);
        }
      }
    }
    const long long t1 = clock64();
    tcgen05_commit(smem_u32(me ? &bar2 : &bar));
    mbar_wait(smem_u32(me ? &bar2 : &bar), 0);
    const long long t2 = clock64();
    out[2 * me] = t1 - t0; out[2 * me + 1] = t2 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 64);
  cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int n = 2048;
  printf("tcgen05.mma M=128 K=16, %d instructions from one thread; cycles per instruction (issue loop / until complete)\n", n);
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {16, 64, 128, 256})
      for (int n_acc : {1, 2, 4}) {

        for (int issuers = 1; issuers <= 2; ++issuers) {
          if ((N > 64 ? N : 64) * n_acc * issuers > 256) continue;          // accumulators in TMEM columns [256, 512)
          for (int rep = 0; rep < 2; ++rep) k_chain<<<1, 128, 100 * 1024>>>(n, N, n_acc, ts, out, issuers);
          long long h[4];
          cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
          cudaError_t e = cudaGetLastError();
          // cycles per instruction of the whole CTA: the slower issuer's span over ALL n instructions
          const double span = issuers == 2 ? (double)(h[1] > h[3] ? h[1] : h[3]) : (double)h[1];
          printf("%s N=%3d accumulators=%d issuing warps=%d: %.1f cycles per instruction (floor %d)%s\n", ts ? "A in TMEM" : "A in smem", N,
                 n_acc, issuers, span / n, 128 * N / 256, e ? cudaGetErrorString(e) : "");
        }
      }
  return 0;
}
