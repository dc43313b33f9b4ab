// Support kernels of the wide-range ("bf16x3") ViT path: every GEMM operand travels as a pair of bf16 numbers per fp32 value
// (hi = bf16(v), lo = bf16(v - hi); row layout [hi (K columns) | lo (K columns)]), the tcgen05 GEMM accumulates
// a_hi w_hi + a_lo w_hi + a_hi w_lo in fp32 (gemm_tcgen05.cu, split mode), attention runs in fp32 on the CUDA cores.
// fp32 exponent range and ~2^-17 relative error per product instead of fp16's 2^-11 and 65504 ceiling: the parity mode for
// trained checkpoints with outlier channels (the reference runs its ViT in fp32, memory_2.py:737-739), at ~1/3 of the GEMM rate.
#include <cuda_bf16.h>
#include "vit_kernels.cuh"

namespace bscnav {

__device__ __forceinline__ void split_bf16(float v, unsigned short& hi, unsigned short& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
}

// fp32 [rows, K] -> bf16 [rows, 2K] (hi | lo)
__global__ void __launch_bounds__(256)
k_split_rows(const float* __restrict__ src, long long rows, int K, unsigned short* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * K) return;
  const long long r = idx / K;
  const int c = (int)(idx - r * K);
  unsigned short hi, lo;
  split_bf16(src[idx], hi, lo);
  dst[r * 2 * K + c] = hi;
  dst[r * 2 * K + K + c] = lo;
}

// already-normalised float images [B,3,S,S] -> split bf16 patch rows [B*Np, 2*k_pad]
__global__ void __launch_bounds__(256)
k_im2col_images_split(const float* __restrict__ img, int batch, int S, int P, int k_pad, unsigned short* __restrict__ a0) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)batch * 3 * S * S;
  if (idx >= total) return;
  const int ox = (int)(idx % S);
  const int oy = (int)((idx / S) % S);
  const int c = (int)((idx / ((long long)S * S)) % 3);
  const int b = (int)(idx / ((long long)3 * S * S));
  const int G = S / P;
  const int py = oy / P, ky = oy - py * P, px = ox / P, kx = ox - px * P;
  unsigned short hi, lo;
  split_bf16(img[idx], hi, lo);
  const long long o = ((long long)b * G * G + py * G + px) * 2 * k_pad + c * P * P + ky * P + kx;
  a0[o] = hi;
  a0[o + k_pad] = lo;
}

// LayerNorm (two passes in registers as k_layernorm) -> split bf16 [rows, 2D]
template <int NV>
__global__ void __launch_bounds__(256)
k_layernorm_split(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                  long long rows, unsigned short* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  pdl_wait();
  pdl_launch_dependents();
  if (row >= rows) return;
  constexpr int D = NV * 128;
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = g4[lane + 32 * i], bb = b4[lane + 32 * i];
    const float o[4] = {(v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y,
                        (v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w};
    unsigned short hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_bf16(o[e], hi[e], lo[e]);
    const uint2 uh = make_uint2((uint32_t)hi[0] | ((uint32_t)hi[1] << 16), (uint32_t)hi[2] | ((uint32_t)hi[3] << 16));
    const uint2 ul = make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 16), (uint32_t)lo[2] | ((uint32_t)lo[3] << 16));
    reinterpret_cast<uint2*>(out + row * 2 * D)[lane + 32 * i] = uh;
    reinterpret_cast<uint2*>(out + row * 2 * D + D)[lane + 32 * i] = ul;
  }
}

// fp32 self-attention on the CUDA cores, head dim 64: one CTA = one (frame, head); K (row stride 65 words: conflict-free when
// lanes read different keys) and V staged in shared memory; every warp owns query rows round robin, two at a time (the K / V
// reads are shared by the pair).  qkv fp32 [batch*T, 3D] (q pre-scaled) -> split bf16 [batch*T, 2D].
static constexpr int kF32Warps = 8;
__global__ void __launch_bounds__(kF32Warps * 32)
k_attention_f32(const float* __restrict__ qkv, unsigned short* __restrict__ out, int T, int heads, int causal) {
  extern __shared__ float sm[];
  float* sK = sm;                               // [T][65]
  float* sV = sK + (((size_t)T * 65 + 3) & ~(size_t)3);   // [T][64], 16-byte aligned
  float* sP = sV + (size_t)T * 64;              // [warps][2][T]
  float* sQ = sP + (size_t)kF32Warps * 2 * T;   // [warps][2][64]
  const int head = blockIdx.x, b = blockIdx.y, D = heads * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  pdl_launch_dependents();
  const float* base = qkv + (long long)b * T * 3 * D + head * 64;
  for (int i = threadIdx.x; i < T * 16; i += blockDim.x) {
    const int j = i >> 4, c = i & 15;
    const float4 k4 = *reinterpret_cast<const float4*>(base + (long long)j * 3 * D + D + 4 * c);
    const float4 v4 = *reinterpret_cast<const float4*>(base + (long long)j * 3 * D + 2 * D + 4 * c);
    float* kd = sK + j * 65 + 4 * c;
    kd[0] = k4.x; kd[1] = k4.y; kd[2] = k4.z; kd[3] = k4.w;
    *reinterpret_cast<float4*>(sV + j * 64 + 4 * c) = v4;
  }
  __syncthreads();
  const float kLog2e = 1.4426950408889634f;
  float* p0 = sP + (warp * 2) * T;
  float* p1 = p0 + T;
  float* q0 = sQ + warp * 128;
  float* q1 = q0 + 64;
  const int n_groups = (T + 31) >> 5;
  for (int r0 = 2 * warp; r0 < T; r0 += 2 * kF32Warps) {
    const int r1 = r0 + 1 < T ? r0 + 1 : r0;
    q0[lane] = base[(long long)r0 * 3 * D + lane]; q0[lane + 32] = base[(long long)r0 * 3 * D + lane + 32];
    q1[lane] = base[(long long)r1 * 3 * D + lane]; q1[lane + 32] = base[(long long)r1 * 3 * D + lane + 32];
    __syncwarp();
    float m0 = -INFINITY, m1 = -INFINITY;
    for (int g = 0; g < n_groups; ++g) {
      const int j = lane + 32 * g;
      float a0 = -INFINITY, a1 = -INFINITY;
      if (j < T) {
        const float* kr = sK + j * 65;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll 16
        for (int d = 0; d < 64; ++d) {
          const float kv = kr[d];
          s0 = fmaf(q0[d], kv, s0);
          s1 = fmaf(q1[d], kv, s1);
        }
        a0 = (causal && j > r0) ? -INFINITY : s0;
        a1 = (causal && j > r1) ? -INFINITY : s1;
        p0[j] = a0; p1[j] = a1;
      }
      m0 = fmaxf(m0, a0); m1 = fmaxf(m1, a1);
    }
    m0 = warp_max(m0); m1 = warp_max(m1);
    float l0 = 0.f, l1 = 0.f;
    for (int g = 0; g < n_groups; ++g) {
      const int j = lane + 32 * g;
      if (j < T) {
        const float e0 = exp2f((p0[j] - m0) * kLog2e), e1 = exp2f((p1[j] - m1) * kLog2e);
        p0[j] = e0; p1[j] = e1;
        l0 += e0; l1 += e1;
      }
    }
    l0 = warp_sum(l0); l1 = warp_sum(l1);
    __syncwarp();
    float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;      // dims 2 lane, 2 lane + 1 of both rows
    for (int j = 0; j < T; ++j) {
      const float2 vv = *reinterpret_cast<const float2*>(sV + j * 64 + 2 * lane);
      const float w0 = p0[j], w1 = p1[j];
      o00 = fmaf(w0, vv.x, o00); o01 = fmaf(w0, vv.y, o01);
      o10 = fmaf(w1, vv.x, o10); o11 = fmaf(w1, vv.y, o11);
    }
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    unsigned short h[2], l[2];
    split_bf16(o00 * i0, h[0], l[0]); split_bf16(o01 * i0, h[1], l[1]);
    unsigned short* d0 = out + ((long long)b * T + r0) * 2 * D + head * 64 + 2 * lane;
    *reinterpret_cast<uint32_t*>(d0) = (uint32_t)h[0] | ((uint32_t)h[1] << 16);
    *reinterpret_cast<uint32_t*>(d0 + D) = (uint32_t)l[0] | ((uint32_t)l[1] << 16);
    if (r1 != r0) {
      split_bf16(o10 * i1, h[0], l[0]); split_bf16(o11 * i1, h[1], l[1]);
      unsigned short* d1 = out + ((long long)b * T + r1) * 2 * D + head * 64 + 2 * lane;
      *reinterpret_cast<uint32_t*>(d1) = (uint32_t)h[0] | ((uint32_t)h[1] << 16);
      *reinterpret_cast<uint32_t*>(d1 + D) = (uint32_t)l[0] | ((uint32_t)l[1] << 16);
    }
    __syncwarp();
  }
}

int split_rows(const float* src, long long rows, int K, void* dst, cudaStream_t stream) {
  if (!src || !dst || rows <= 0 || K <= 0) return BSCNAV_EINVAL;
  BSC_LAUNCH(k_split_rows, (unsigned)((rows * K + 255) / 256), 256, 0, stream, src, rows, K, static_cast<unsigned short*>(dst));
  return BSCNAV_OK;
}

int im2col_images_split(const float* img, int batch, int S, int P, int k_pad, void* a0, cudaStream_t stream) {
  const long long total = (long long)batch * 3 * S * S;
  BSC_LAUNCH(k_im2col_images_split, (unsigned)((total + 255) / 256), 256, 0, stream, img, batch, S, P, k_pad, static_cast<unsigned short*>(a0));
  return BSCNAV_OK;
}

int layernorm_split(const float* x, const float* g, const float* b, float eps, long long rows, int D, void* out, cudaStream_t stream) {
  if (D % 128) return BSCNAV_EINVAL;
  const int grid = (int)((rows + 7) / 8);
  unsigned short* o = static_cast<unsigned short*>(out);
#define BSC_LNS_CASE(NV) case NV: BSC_LAUNCH_PDL((k_layernorm_split<NV>), grid, 256, 0, stream, x, g, b, eps, rows, o); break;
  switch (D / 128) {
    BSC_LNS_CASE(1) BSC_LNS_CASE(2) BSC_LNS_CASE(3) BSC_LNS_CASE(4) BSC_LNS_CASE(5) BSC_LNS_CASE(6) BSC_LNS_CASE(7) BSC_LNS_CASE(8)
    BSC_LNS_CASE(10) BSC_LNS_CASE(12) BSC_LNS_CASE(16)
    default: return BSCNAV_EINVAL;
  }
#undef BSC_LNS_CASE
  return BSCNAV_OK;
}

int attention_f32_split(const float* qkv, void* out, int batch, int T, int heads, bool causal, cudaStream_t stream) {
  if (!qkv || !out || batch <= 0 || T <= 0 || heads <= 0) return BSCNAV_EINVAL;
  const size_t smem = ((((size_t)T * 65 + 3) & ~(size_t)3) + (size_t)T * 64 + (size_t)kF32Warps * 2 * T + kF32Warps * 128) * sizeof(float);
  if (smem > 220 * 1024) return BSCNAV_EINVAL;
  static size_t attr = 0;
  if (smem > attr) {
    BSC_CUDA(cudaFuncSetAttribute(k_attention_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  BSC_LAUNCH_PDL(k_attention_f32, dim3((unsigned)heads, (unsigned)batch), kF32Warps * 32, smem, stream, qkv,
                 static_cast<unsigned short*>(out), T, heads, causal ? 1 : 0);
  return BSCNAV_OK;
}

}  // namespace bscnav
