// micro-benchmark: legacy mma.sync.m16n8k16 fp16 peak on sm_100a (is an mma.sync attention kernel worth tuning?)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
  float c[8][4] = {};
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  for (int warps : {4, 8, 16, 32}) {
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148 * 2, warps * 32 / 2>>>(out, 100);
    cudaEventRecord(e0);
    k<<<148 * 2, warps * 32 / 2>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 16 * 8 * 16 * 8 * iters * (double)(148 * warps);
    printf("warps/SM=%d: %.3f ms  %.1f TFLOP/s\n", warps, ms, flops / ms / 1e9);
  }
  return 0;
}
