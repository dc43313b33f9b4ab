// ViT support kernels (sm_100a): fused pre-processing + im2col, LayerNorm, attention, token assembly,
// Gaussian query pooling.  The dense contractions live in gemm_tcgen05.cu.
#include "vit_kernels.cuh"

namespace bscnav {

// ---------------------------------------------------------------------------------------------------------
// Pre-processing fused with im2col (memory_2.py:71-74,733-735): uint8 HWC frame -> /255 -> antialiased bilinear
// resize (ATen upsample_bilinear2d_aa: triangle filter, support = max(scale,1), weights normalised, horizontal
// pass then vertical) -> (x-mean)/std -> fp16 patch rows [frame*Np + patch, c*P*P + ky*P + kx].
static constexpr int kMaxTaps = 34;

template <int MAXT = kMaxTaps>
__device__ __forceinline__ int aa_taps(int i, float scale, int in_size, float* w, int* start) {
  const float support = scale >= 1.f ? scale : 1.f;
  const float invscale = scale >= 1.f ? 1.f / scale : 1.f;
  // explicit roundings (no FMA contraction): a tap boundary such as center - support + 0.5 == 68.0 must fall the way ATen's does
  const float center = __fmul_rn(scale, __fadd_rn((float)i, 0.5f));
  int lo = (int)__fadd_rn(__fsub_rn(center, support), 0.5f);
  if (lo < 0) lo = 0;
  int hi = (int)__fadd_rn(__fadd_rn(center, support), 0.5f);
  if (hi > in_size) hi = in_size;
  int n = hi - lo;
  if (n > MAXT) n = MAXT;
  float total = 0.f;
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    if (j < n) {
      float x = __fmul_rn(__fadd_rn(__fsub_rn((float)(j + lo), center), 0.5f), invscale);
      x = fabsf(x);
      const float v = x < 1.f ? 1.f - x : 0.f;
      w[j] = v;
      total += v;
    }
  }
  if (total != 0.f) {
#pragma unroll
    for (int j = 0; j < MAXT; ++j)
      if (j < n) w[j] /= total;
  }
  *start = lo;
  return n;
}

__global__ void __launch_bounds__(256)
k_preprocess_frames(const uint8_t* __restrict__ rgb, int batch, int H, int W, int ch, int S, int P, int k_pad,
                    __half* __restrict__ a0, float* __restrict__ out_pre) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)batch * S * S;
  if (idx >= total) return;
  const int ox = (int)(idx % S);
  const int oy = (int)((idx / S) % S);
  const int b = (int)(idx / ((long long)S * S));
  float wx[kMaxTaps], wy[kMaxTaps];
  int x0, y0;
  const int nx = aa_taps(ox, (float)W / (float)S, W, wx, &x0);
  const int ny = aa_taps(oy, (float)H / (float)S, H, wy, &y0);
  float acc[3] = {0.f, 0.f, 0.f};
  const uint8_t* frame = rgb + (long long)b * H * W * ch;
  for (int jy = 0; jy < ny; ++jy) {
    const uint8_t* rowp = frame + ((long long)(y0 + jy) * W + x0) * ch;
    float r[3] = {0.f, 0.f, 0.f};
    for (int jx = 0; jx < nx; ++jx) {
      const float w = wx[jx];
      r[0] += w * ((float)rowp[jx * ch + 0] / 255.f);
      r[1] += w * ((float)rowp[jx * ch + 1] / 255.f);
      r[2] += w * ((float)rowp[jx * ch + 2] / 255.f);
    }
    acc[0] += wy[jy] * r[0]; acc[1] += wy[jy] * r[1]; acc[2] += wy[jy] * r[2];
  }
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  const int G = S / P;
  const int py = oy / P, ky = oy - py * P, px = ox / P, kx = ox - px * P;
  const long long rowi = (long long)b * G * G + py * G + px;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = (acc[c] - mean[c]) / stdv[c];
    if (out_pre) out_pre[((long long)(b * 3 + c) * S + oy) * S + ox] = v;
    a0[rowi * k_pad + c * P * P + ky * P + kx] = __float2half_rn(v);
  }
}

// Tiled form of the same computation for scale factors <= 4 (640x480, 680x680 -> 224): a block produces a 32 x 8 tile of output
// pixels from a source window staged once in shared memory with coalesced loads (the per-pixel kernel above issues ~130 byte
// loads and as many divisions per thread: 276 us for 54 frames, 0.24 TB/s).  Same weights and summation order; the /255 is
// folded into the horizontal weights.
static constexpr int kPreTX = 32, kPreTY = 8, kPreTaps = 10, kPreWinW = 140, kPreWinH = 44;
static constexpr int kPreTabStride = kPreTaps + 3;       // per output coordinate: 10 weights, tap count, first source index; odd stride: lanes read their own row conflict-free

// tap tables of one (H, W) -> S resize: tab[0 .. S) for x, tab[S .. 2S) for y (the per-thread tap computation, with its
// float divisions, was 3/4 of the tiled kernel's time: ncu SM 76 %, 189 us per 54 frames)
__global__ void k_aa_tables(int H, int W, int S, float* __restrict__ tab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * S) return;
  const bool isy = i >= S;
  const int o = isy ? i - S : i;
  float w[kPreTaps];
  int start;
  const int n = aa_taps<kPreTaps>(o, isy ? (float)H / (float)S : (float)W / (float)S, isy ? H : W, w, &start);
  float* t = tab + (long long)i * kPreTabStride;
  // x weights carry the /255 of memory_2.py:733 (w * (b / 255) == (w / 255) * b up to one rounding, far inside the 2e-6 parity bound)
#pragma unroll
  for (int j = 0; j < kPreTaps; ++j) t[j] = j < n ? (isy ? w[j] : w[j] / 255.f) : 0.f;
  t[kPreTaps] = __int_as_float(n);
  t[kPreTaps + 1] = __int_as_float(start);
}

__global__ void __launch_bounds__(kPreTX * kPreTY)
k_preprocess_tiled(const uint8_t* __restrict__ rgb, int batch, int H, int W, int ch, int S, int P, int k_pad,
                   const float* __restrict__ tab, __half* __restrict__ a0, float* __restrict__ out_pre) {
  __shared__ uint32_t win[kPreWinH * kPreWinW];
  __shared__ float s_tx[kPreTX * kPreTabStride], s_ty[kPreTY * kPreTabStride];
  const int tid = threadIdx.x, tx = tid & (kPreTX - 1), ty = tid / kPreTX;
  const int b = blockIdx.z, ox0 = blockIdx.x * kPreTX, oy0 = blockIdx.y * kPreTY;
  for (int i = tid; i < kPreTX * kPreTabStride; i += kPreTX * kPreTY) {
    const int o = ox0 + i / kPreTabStride;
    s_tx[i] = o < S ? tab[(long long)o * kPreTabStride + i % kPreTabStride] : 0.f;
  }
  if (tid < kPreTY * kPreTabStride) {
    const int o = oy0 + tid / kPreTabStride;
    s_ty[tid] = o < S ? tab[(long long)(S + o) * kPreTabStride + tid % kPreTabStride] : 0.f;
  }
  __syncthreads();
  const int lx = min(ox0 + kPreTX - 1, S - 1) - ox0, ly = min(oy0 + kPreTY - 1, S - 1) - oy0;
  const int wx0 = __float_as_int(s_tx[kPreTaps + 1]), wy0 = __float_as_int(s_ty[kPreTaps + 1]);
  const int ww = __float_as_int(s_tx[lx * kPreTabStride + kPreTaps + 1]) + __float_as_int(s_tx[lx * kPreTabStride + kPreTaps]) - wx0;
  const int wh = __float_as_int(s_ty[ly * kPreTabStride + kPreTaps + 1]) + __float_as_int(s_ty[ly * kPreTabStride + kPreTaps]) - wy0;
  const uint8_t* frame = rgb + (long long)b * H * W * ch;
  for (int i = tid; i < ww * wh; i += kPreTX * kPreTY) {
    const int y = i / ww, x = i - y * ww;
    const uint8_t* px = frame + ((long long)(wy0 + y) * W + wx0 + x) * ch;
    uint32_t v;
    if (ch == 4) v = *reinterpret_cast<const uint32_t*>(px);
    else v = (uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16);
    win[y * kPreWinW + x] = v;
  }
  __syncthreads();
  const int ox = ox0 + tx, oy = oy0 + ty;
  if (ox >= S || oy >= S) return;
  const float* wxs = s_tx + tx * kPreTabStride;
  const float* wys = s_ty + ty * kPreTabStride;
  const int nx = __float_as_int(wxs[kPreTaps]), x0 = __float_as_int(wxs[kPreTaps + 1]);
  const int ny = __float_as_int(wys[kPreTaps]), y0 = __float_as_int(wys[kPreTaps + 1]);
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int jy = 0; jy < kPreTaps; ++jy) {
    if (jy < ny) {
      const uint32_t* rowp = win + (y0 + jy - wy0) * kPreWinW + (x0 - wx0);
      float r[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int jx = 0; jx < kPreTaps; ++jx) {
        if (jx < nx) {
          const float w = wxs[jx];
          const uint32_t v = rowp[jx];
          // byte -> float on the FMA pipe: 0x4B000000 | b is the float 2^23 + b (I2F runs on the XU pipe, a 256-entry table
          // in shared memory bank-conflicts: both measured slower)
          r[0] = fmaf(w, __uint_as_float(0x4B000000u | (v & 255u)) - 8388608.f, r[0]);
          r[1] = fmaf(w, __uint_as_float(0x4B000000u | ((v >> 8) & 255u)) - 8388608.f, r[1]);
          r[2] = fmaf(w, __uint_as_float(0x4B000000u | ((v >> 16) & 255u)) - 8388608.f, r[2]);
        }
      }
      acc[0] += wys[jy] * r[0]; acc[1] += wys[jy] * r[1]; acc[2] += wys[jy] * r[2];
    }
  }
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  const int G = S / P;
  const int py = oy / P, ky = oy - py * P, pxx = ox / P, kx = ox - pxx * P;
  const long long rowi = (long long)b * G * G + py * G + pxx;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = (acc[c] - mean[c]) / stdv[c];
    if (out_pre) out_pre[((long long)(b * 3 + c) * S + oy) * S + ox] = v;
    a0[rowi * k_pad + c * P * P + ky * P + kx] = __float2half_rn(v);
  }
}

// already-normalised float images [B,3,S,S] -> fp16 patch rows (query path, memory_2.py:577-580)
__global__ void __launch_bounds__(256)
k_im2col_images(const float* __restrict__ img, int batch, int S, int P, int k_pad, __half* __restrict__ a0) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)batch * 3 * S * S;
  if (idx >= total) return;
  const int ox = (int)(idx % S);
  const int oy = (int)((idx / S) % S);
  const int c = (int)((idx / ((long long)S * S)) % 3);
  const int b = (int)(idx / ((long long)3 * S * S));
  const int G = S / P;
  const int py = oy / P, ky = oy - py * P, px = ox / P, kx = ox - px * P;
  a0[((long long)b * G * G + py * G + px) * k_pad + c * P * P + ky * P + kx] = __float2half_rn(img[idx]);
}

// cls (+pos[0]) and register tokens of every frame
__global__ void k_fill_prefix_tokens(float* __restrict__ x, const float* __restrict__ cls_pos, const float* __restrict__ reg,
                                     int batch, int tokens, int n_reg, int D) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)batch * (1 + n_reg) * D;
  if (idx >= total) return;
  const int c = (int)(idx % D);
  const int t = (int)((idx / D) % (1 + n_reg));
  const int b = (int)(idx / ((long long)D * (1 + n_reg)));
  x[((long long)b * tokens + t) * D + c] = t == 0 ? cls_pos[c] : reg[(long long)(t - 1) * D + c];
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm over the last dimension, one warp per row, two passes in registers (mean, then centred variance).
//   MODE 0: fp16 out (GEMM operand)   MODE 1: fp32 in place   MODE 2: fp32 out, patch rows only (drop cls/registers)
template <int NV, int MODE>
__global__ void __launch_bounds__(256)
k_layernorm(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
            long long rows, void* __restrict__ out, int tokens, int skip) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  pdl_wait();                      // x comes from the previous kernel in the stream (programmatic dependent launch)
  pdl_launch_dependents();
  if (row >= rows) return;
  constexpr int D = NV * 128;
  long long in_row = row;
  if (MODE == 2) {
    const int np = tokens - skip;
    const long long f = row / np;
    in_row = f * tokens + skip + (row - f * np);
  }
  const float4* xr = reinterpret_cast<const float4*>(x + in_row * D);
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
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + bb.x;
    o.y = (v[i].y - mean) * rstd * g.y + bb.y;
    o.z = (v[i].z - mean) * rstd * g.z + bb.z;
    o.w = (v[i].w - mean) * rstd * g.w + bb.w;
    if (MODE == 0) {
      __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      reinterpret_cast<uint2*>(static_cast<__half*>(out) + row * D)[lane + 32 * i] = u;
    } else {
      reinterpret_cast<float4*>(static_cast<float*>(out) + row * D)[lane + 32 * i] = o;
    }
  }
}

template <int MODE>
static int layernorm_dispatch(const float* x, const float* g, const float* b, float eps, long long rows, int D, void* out,
                              int tokens, int skip, cudaStream_t stream) {
  const int grid = (int)((rows + 7) / 8);
#define BSC_LN_CASE(NV) case NV: BSC_LAUNCH_PDL((k_layernorm<NV, MODE>), grid, 256, 0, stream, x, g, b, eps, rows, out, tokens, skip); break;
  switch (D / 128) {
    BSC_LN_CASE(1) BSC_LN_CASE(2) BSC_LN_CASE(3) BSC_LN_CASE(4) BSC_LN_CASE(5) BSC_LN_CASE(6) BSC_LN_CASE(7) BSC_LN_CASE(8)
    BSC_LN_CASE(10) BSC_LN_CASE(12) BSC_LN_CASE(16)
    default: return BSCNAV_EINVAL;
  }
#undef BSC_LN_CASE
  return BSCNAV_OK;
}

int layernorm_half(const float* x, const float* g, const float* b, float eps, long long rows, int D, __half* out, cudaStream_t s) {
  if (D % 128) return BSCNAV_EINVAL;
  return layernorm_dispatch<0>(x, g, b, eps, rows, D, out, 0, 0, s);
}
int layernorm_inplace(float* x, const float* g, const float* b, float eps, long long rows, int D, cudaStream_t s) {
  if (D % 128) return BSCNAV_EINVAL;
  return layernorm_dispatch<1>(x, g, b, eps, rows, D, x, 0, 0, s);
}
int layernorm_patch_tokens(const float* x, const float* g, const float* b, float eps, int batch, int tokens, int skip, int D,
                           float* out, cudaStream_t s) {
  if (D % 128) return BSCNAV_EINVAL;
  return layernorm_dispatch<2>(x, g, b, eps, (long long)batch * (tokens - skip), D, out, tokens, skip, s);
}

// ---------------------------------------------------------------------------------------------------------
// Multi-head self-attention, head dim 64 or 80 (ViT-H/14), fp16 operands, fp32 softmax / accumulation.
// One CTA = one (frame, head): K and V are staged once in shared memory with cp.async (16-byte chunks XOR-swizzled
// by row so ldmatrix is conflict-free) and shared by all query rows; each of the 6 warps owns 16-row query groups
// and runs a flash-style online softmax over 64-key blocks with mma.sync.m16n8k16, no block-level barriers in the loop.
// 4% of the ViT's FLOPs; the GEMMs carry the tcgen05 work.

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  // not volatile: a pure register operation, so the scheduler may interleave it with the (ordered) ldmatrix stream
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fast_exp2(float x) {      // single MUFU.EX2; inputs are <= 0, flush-to-zero is the wanted limit
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of (row, 16-byte chunk) in a [rows][64] fp16 tile with the chunk index XOR-swizzled by row
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }
// K / V row layout in shared memory by head dim: 64 -> 128-byte rows, 16-byte chunks XOR-swizzled by row;
// 80 (ViT-H/14) -> 160-byte rows padded to 176 (44 words: eight consecutive rows start 12 banks apart, so the 8 x 16-byte
// reads of one ldmatrix phase cover all 32 banks once), no swizzle.
template <int HD> __device__ __forceinline__ uint32_t kv_off(int row, int chunk) {
  if (HD == 64) return sw_off(row, chunk);
  return (uint32_t)(row * (HD * 2 + 16) + (chunk << 4));
}
template <int HD> constexpr int kv_row_bytes() { return HD == 64 ? 128 : HD * 2 + 16; }

// 16-byte async copy global -> shared; src_bytes == 0 zero-fills (rows past the sequence end)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

static constexpr int kAttnWarps = 6;        // 257 (261) tokens = 17 row groups of 16 -> 3 rounds of 6 warps, 94% of lanes busy
static constexpr int kAttnThreads = kAttnWarps * 32;

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(kAttnThreads, 2)
k_attention(const __half* __restrict__ qkv, __half* __restrict__ out, int T, int heads, int t_pad) {
  extern __shared__ __align__(128) uint8_t smem_attn[];
  constexpr int kHd = HD, KS = HD / 16, NT = HD / 8, CH = HD / 8;     // k-steps of QK^T, 8-wide output tiles, 16-byte chunks per row
  const int D = heads * kHd;
  const int head = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __half* base = qkv + (long long)b * T * 3 * D + head * kHd;
  const uint32_t sKa = (uint32_t)__cvta_generic_to_shared(smem_attn);
  const uint32_t sVa = sKa + t_pad * kv_row_bytes<HD>();
  // K and V of this (frame, head): staged once, shared by every query row of the head
  for (int i = tid; i < t_pad * CH; i += kAttnThreads) {
    const int r = i / CH, c = i - r * CH;
    const __half* src = base + (long long)(r < T ? r : 0) * 3 * D + c * 8;
    const int nb = r < T ? 16 : 0;
    cp_async16(sKa + kv_off<HD>(r, c), src + D, nb);
    cp_async16(sVa + kv_off<HD>(r, c), src + 2 * D, nb);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const int n_kb = (T + 63) / 64;
  const int n_groups = (T + 15) / 16;
  const float kLog2e = 1.4426950408889634f;
  // gridDim.z CTAs share a (frame, head): each takes every gridDim.z-th round of row groups (small batches: one frame per call
  // has 16 (frame, head) pairs for 148 SMs; K and V are staged by each of them, 66 KB from L2)
  for (int g = (int)blockIdx.z * kAttnWarps + warp; g < n_groups; g += kAttnWarps * (int)gridDim.z) {
    // Q fragments (m16k16 A operand, 4 k-steps) straight from global memory: rows r0 and r0+8 of this group
    const int r0 = g * 16 + (lane >> 2), r1 = r0 + 8;
    uint32_t qf[KS][4];
    {
      const __half* q0 = base + (long long)r0 * 3 * D + 2 * (lane & 3);
      const __half* q1 = base + (long long)r1 * 3 * D + 2 * (lane & 3);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        qf[ks][0] = r0 < T ? *reinterpret_cast<const uint32_t*>(q0 + ks * 16) : 0u;
        qf[ks][1] = r1 < T ? *reinterpret_cast<const uint32_t*>(q1 + ks * 16) : 0u;
        qf[ks][2] = r0 < T ? *reinterpret_cast<const uint32_t*>(q0 + ks * 16 + 8) : 0u;
        qf[ks][3] = r1 < T ? *reinterpret_cast<const uint32_t*>(q1 + ks * 16 + 8) : 0u;
      }
    }
    float o[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // running max / sum of rows r0 and r1

    // causal (text tower): keys <= query row only, so key blocks past this row group are skipped; a visited block starts at
    // or before the group's first row (kb*64 <= g*16), hence still holds a valid key for every row of the group
    const int n_kb_g = CAUSAL ? min(n_kb, (g * 16 + 15) / 64 + 1) : n_kb;
    for (int kb = 0; kb < n_kb_g; ++kb) {
      // 16-key groups of this block that hold at least one real key (T = 257: the last block has one of four)
      const int ng = (kb == n_kb - 1) ? (T - kb * 64 + 15) >> 4 : 4;
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
      // S = Q K^T for 64 keys: B fragments straight from K rows (key-major, head dim contiguous).
      // k-step outermost: the eight 8-key accumulator tiles form eight independent mma chains.
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t kf[4][4];
#pragma unroll
        for (int np = 0; np < 4; ++np) {     // pairs of 8-key tiles
          if (np < ng) {
            const int key = kb * 64 + np * 16 + (lane & 7) + ((lane >> 4) << 3);
            ldsm_x4(sKa + kv_off<HD>(key, ks * 2 + ((lane >> 3) & 1)), kf[np][0], kf[np][1], kf[np][2], kf[np][3]);
          }
        }
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          if (np < ng) {
            mma_16816(s[2 * np], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], kf[np][0], kf[np][1]);
            mma_16816(s[2 * np + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], kf[np][2], kf[np][3]);
          }
        }
      }
      // mask keys >= T (last block only), block row maxima
      if (kb == n_kb - 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int key = kb * 64 + i * 8 + 2 * (lane & 3);
          if (key >= T) { s[i][0] = -INFINITY; s[i][2] = -INFINITY; }
          if (key + 1 >= T) { s[i][1] = -INFINITY; s[i][3] = -INFINITY; }
        }
      }
      if (CAUSAL && kb * 64 + 63 > g * 16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int key = kb * 64 + i * 8 + 2 * (lane & 3);
          if (key > r0) s[i][0] = -INFINITY;
          if (key + 1 > r0) s[i][1] = -INFINITY;
          if (key > r1) s[i][2] = -INFINITY;
          if (key + 1 > r1) s[i][3] = -INFINITY;
        }
      }
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        bm0 = fmaxf(bm0, fmaxf(s[i][0], s[i][1]));
        bm1 = fmaxf(bm1, fmaxf(s[i][2], s[i][3]));
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);      // finite: every block holds at least one valid key
      const float sc0 = fast_exp2((m0 - nm0) * kLog2e), sc1 = fast_exp2((m1 - nm1) * kLog2e);
      m0 = nm0; m1 = nm1;
      const float mo0 = m0 * kLog2e, mo1 = m1 * kLog2e;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i][0] = fast_exp2(fmaf(s[i][0], kLog2e, -mo0)); s[i][1] = fast_exp2(fmaf(s[i][1], kLog2e, -mo0));
        s[i][2] = fast_exp2(fmaf(s[i][2], kLog2e, -mo1)); s[i][3] = fast_exp2(fmaf(s[i][3], kLog2e, -mo1));
        ps0 += s[i][0] + s[i][1];
        ps1 += s[i][2] + s[i][3];
      }
      l0 = l0 * sc0 + ps0;
      l1 = l1 * sc1 + ps1;
#pragma unroll
      for (int i = 0; i < NT; ++i) { o[i][0] *= sc0; o[i][1] *= sc0; o[i][2] *= sc1; o[i][3] *= sc1; }
      // O += P V : P from the S accumulators (fp16), V fragments via transposing ldmatrix
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {       // 16 keys per step
        if (ks < ng) {
          const uint32_t a0 = pack_half2(s[2 * ks][0], s[2 * ks][1]), a1 = pack_half2(s[2 * ks][2], s[2 * ks][3]);
          const uint32_t a2 = pack_half2(s[2 * ks + 1][0], s[2 * ks + 1][1]), a3 = pack_half2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
          uint32_t vf[KS][4];
          const int key = kb * 64 + ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
#pragma unroll
          for (int np = 0; np < KS; ++np)      // pairs of 8-wide head-dim tiles
            ldsm_x4_t(sVa + kv_off<HD>(key, np * 2 + (lane >> 4)), vf[np][0], vf[np][1], vf[np][2], vf[np][3]);
#pragma unroll
          for (int np = 0; np < KS; ++np) {
            mma_16816(o[2 * np], a0, a1, a2, a3, vf[np][0], vf[np][1]);
            mma_16816(o[2 * np + 1], a0, a1, a2, a3, vf[np][2], vf[np][3]);
          }
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    __half* ob = out + (long long)b * T * D + head * kHd + 2 * (lane & 3);
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      if (r0 < T) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * D + i * 8) = pack_half2(o[i][0] * i0, o[i][1] * i0);
      if (r1 < T) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * D + i * 8) = pack_half2(o[i][2] * i1, o[i][3] * i1);
    }
  }
}

static int sm_count_cached() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int HD, bool CAUSAL>
static int attention_launch(const __half* qkv, __half* out, int batch, int T, int heads, cudaStream_t stream) {
  const int t_pad = (T + 15) / 16 * 16;      // only whole 16-key groups are ever read (see `ng` in the kernel)
  const int smem = 2 * t_pad * kv_row_bytes<HD>();
  if (smem > 200 * 1024) return BSCNAV_EINVAL;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    BSC_CUDA(cudaFuncSetAttribute((k_attention<HD, CAUSAL>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  // row-group rounds of one (frame, head) over up to three CTAs while the grid stays within one wave (two CTAs per SM)
  const int rounds = ((T + 15) / 16 + kAttnWarps - 1) / kAttnWarps;
  int z = (2 * sm_count_cached()) / (batch * heads);
  z = z < 1 ? 1 : (z > rounds ? rounds : z);
  if (z > 4) z = 4;
  BSC_LAUNCH((k_attention<HD, CAUSAL>), dim3((unsigned)heads, (unsigned)batch, (unsigned)z), kAttnThreads, smem, stream, qkv, out, T, heads, t_pad);
  return BSCNAV_OK;
}

int attention_f16(const __half* qkv, __half* out, int batch, int T, int heads, int head_dim, bool causal, cudaStream_t stream) {
  if (batch <= 0 || T <= 0 || heads <= 0) return BSCNAV_EINVAL;
  if (head_dim == 64) return causal ? attention_launch<64, true>(qkv, out, batch, T, heads, stream)
                                    : attention_launch<64, false>(qkv, out, batch, T, heads, stream);
  if (head_dim == 80 && !causal) return attention_launch<80, false>(qkv, out, batch, T, heads, stream);
  return BSCNAV_EINVAL;
}

// ---------------------------------------------------------------------------------------------------------
// Gaussian centre-weighted pooling (memory_2.py:596-608): q = mean_b sum_i w_i t[b,i,:], w_i ~ exp(-d_i / (2 (g/2)^2))
__global__ void k_gaussian_pool(const float* __restrict__ tokens, int batch, int grid, int D, float* __restrict__ out) {
  extern __shared__ float w[];
  const int n = grid * grid;
  const float center = (grid - 1) / 2.f, sigma = (grid / 2.f) * (grid / 2.f);
  __shared__ float wsum;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float xs = (float)(i % grid), ys = (float)(i / grid);
    const float d = (xs - center) * (xs - center) + (ys - center) * (ys - center);
    w[i] = expf(-d / (2.f * sigma));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += w[i];
    wsum = s;
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float acc_b = 0.f;
  for (int b = 0; b < batch; ++b) {
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc += tokens[((long long)b * n + i) * D + c] * (w[i] / wsum);
    acc_b += acc;
  }
  out[c] = acc_b / (float)batch;
}

// ---------------------------------------------------------------------------------------------------------
// Image-embedding head (open_clip VisionTransformer: pooled = ln_post(x[:, 0]); pooled @ proj): one block per frame.
// LayerNorm in two passes (mean, centred variance) like k_layernorm; the D x E product in fp32, proj read coalesced.
// ids != null (text tower, open_clip encode_text): the pooled row is the end-of-text position = argmax of the token ids
// (first position of the largest id), instead of row 0.
__global__ void __launch_bounds__(256)
k_cls_embed(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int tokens,
            int D, const float* __restrict__ proj, int E, float* __restrict__ out, const int32_t* __restrict__ ids) {
  extern __shared__ float cls_row[];
  __shared__ float red[8];
  __shared__ int pooled_row;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    int row = 0;
    if (ids) {
      int best = ids[(long long)b * tokens];
      for (int t = 1; t < tokens; ++t) {
        const int v = ids[(long long)b * tokens + t];
        if (v > best) { best = v; row = t; }
      }
    }
    pooled_row = row;
  }
  __syncthreads();
  const float* xr = x + ((long long)b * tokens + pooled_row) * D;
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    return t;
  };
  float s = 0.f;
  for (int i = tid; i < D; i += 256) { cls_row[i] = xr[i]; s += cls_row[i]; }
  const float mean = block_sum(s) / (float)D;
  float q = 0.f;
  for (int i = tid; i < D; i += 256) { const float d = cls_row[i] - mean; q += d * d; }
  const float rstd = rsqrtf(block_sum(q) / (float)D + eps);
  for (int i = tid; i < D; i += 256) cls_row[i] = (cls_row[i] - mean) * rstd * gamma[i] + beta[i];
  __syncthreads();
  for (int e = tid; e < E; e += 256) {
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc = fmaf(cls_row[d], proj[(long long)d * E + e], acc);
    out[(long long)b * E + e] = acc;
  }
}

int cls_embed(const float* x, const float* g, const float* b, float eps, int batch, int tokens, int D, const float* proj, int E,
              float* out, const int32_t* ids, cudaStream_t stream) {
  if (batch <= 0 || D <= 0 || E <= 0 || D * 4 > 48 * 1024) return BSCNAV_EINVAL;
  BSC_LAUNCH(k_cls_embed, (unsigned)batch, 256, D * sizeof(float), stream, x, g, b, eps, tokens, D, proj, E, out, ids);
  return BSCNAV_OK;
}

// Text tower input (open_clip encode_text): x[b, t, :] = token_embedding[ids[b, t]] + positional_embedding[t]
__global__ void k_text_embed(const int32_t* __restrict__ ids, const float* __restrict__ tok_emb, int vocab, const float* __restrict__ pos,
                             int batch, int ctx, int D, float* __restrict__ x) {
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // one float4 per thread
  const int d4 = D >> 2;
  if (i4 >= (long long)batch * ctx * d4) return;
  const int c = (int)(i4 % d4);
  const long long bt = i4 / d4;
  const int t = (int)(bt % ctx);
  int id = ids[bt];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float4 a = reinterpret_cast<const float4*>(tok_emb + (long long)id * D)[c];
  const float4 p = reinterpret_cast<const float4*>(pos + (long long)t * D)[c];
  reinterpret_cast<float4*>(x)[i4] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
}
int text_embed(const int32_t* ids, const float* tok_emb, int vocab, const float* pos, int batch, int ctx, int D, float* x,
               cudaStream_t stream) {
  if (batch <= 0 || ctx <= 0 || vocab <= 0 || D % 4) return BSCNAV_EINVAL;
  const long long total = (long long)batch * ctx * (D / 4);
  BSC_LAUNCH(k_text_embed, (unsigned)((total + 255) / 256), 256, 0, stream, ids, tok_emb, vocab, pos, batch, ctx, D, x);
  return BSCNAV_OK;
}

// View scoring (check_around, BSCAgent.py:713-718): cosine of each view embedding with the goal, softmax over the views, argmax.
__global__ void __launch_bounds__(256)
k_score_views(const float* __restrict__ embeds, const float* __restrict__ goal, int n, int D, float* __restrict__ prob,
              int32_t* __restrict__ best) {
  __shared__ float sim[1024];
  __shared__ float g_norm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) {
    float q = 0.f;
    for (int d = lane; d < D; d += 32) q += goal[d] * goal[d];
    q = warp_sum(q);
    if (lane == 0) g_norm = fmaxf(sqrtf(q), 1e-12f);      // F.normalize: x / max(|x|, eps), eps = 1e-12
  }
  __syncthreads();
  for (int i = warp; i < n; i += 8) {
    const float* v = embeds + (long long)i * D;
    float dot = 0.f, q = 0.f;
    for (int d = lane; d < D; d += 32) { dot = fmaf(v[d], goal[d], dot); q = fmaf(v[d], v[d], q); }
    dot = warp_sum(dot); q = warp_sum(q);
    if (lane == 0) sim[i] = dot / (fmaxf(sqrtf(q), 1e-12f) * g_norm);
  }
  __syncthreads();
  if (warp == 0) {
    float m = -INFINITY;
    int arg = 0x7fffffff;
    for (int i = lane; i < n; i += 32)
      if (sim[i] > m) { m = sim[i]; arg = i; }
    for (int o = 16; o; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > m || (om == m && oa < arg)) { m = om; arg = oa; }
    }
    float z = 0.f;
    for (int i = lane; i < n; i += 32) z += expf(sim[i] - m);
    z = warp_sum(z);
    for (int i = lane; i < n; i += 32) prob[i] = expf(sim[i] - m) / z;
    if (lane == 0) { best[0] = arg; best[1] = 0; }
  }
}

int64_t preprocess_table_bytes(int S) { return (int64_t)2 * S * kPreTabStride * sizeof(float); }

int preprocess_frames(const uint8_t* rgb, int batch, int H, int W, int ch, int S, int P, int k_pad, __half* a0, float* out_pre,
                      float* tables, cudaStream_t stream) {
  if (batch <= 0 || H <= 0 || W <= 0 || (ch != 3 && ch != 4) || S % P) return BSCNAV_EINVAL;
  if ((float)W / S > 16.f || (float)H / S > 16.f) return BSCNAV_EINVAL;   // tap buffer bound
  const float sx = (float)W / S, sy = (float)H / S;
  const bool aligned4 = ch != 4 || (reinterpret_cast<uintptr_t>(rgb) & 3) == 0;
  // window bound: 32 (8) outputs x scale + two supports + rounding must fit the staged window
  if (tables && sx <= 4.f && sy <= 4.f && aligned4 && kPreTX * sx + 2 * fmaxf(sx, 1.f) + 3 <= kPreWinW && kPreTY * sy + 2 * fmaxf(sy, 1.f) + 3 <= kPreWinH) {
    BSC_LAUNCH(k_aa_tables, (2 * S + 255) / 256, 256, 0, stream, H, W, S, tables);
    BSC_LAUNCH(k_preprocess_tiled, dim3((S + kPreTX - 1) / kPreTX, (S + kPreTY - 1) / kPreTY, batch), kPreTX * kPreTY, 0, stream, rgb,
               batch, H, W, ch, S, P, k_pad, tables, a0, out_pre);
    return BSCNAV_OK;
  }
  const long long total = (long long)batch * S * S;
  BSC_LAUNCH(k_preprocess_frames, (unsigned)((total + 255) / 256), 256, 0, stream, rgb, batch, H, W, ch, S, P, k_pad, a0, out_pre);
  return BSCNAV_OK;
}
int im2col_images(const float* img, int batch, int S, int P, int k_pad, __half* a0, cudaStream_t stream) {
  const long long total = (long long)batch * 3 * S * S;
  BSC_LAUNCH(k_im2col_images, (unsigned)((total + 255) / 256), 256, 0, stream, img, batch, S, P, k_pad, a0);
  return BSCNAV_OK;
}
int fill_prefix_tokens(float* x, const float* cls_pos, const float* reg, int batch, int tokens, int n_reg, int D, cudaStream_t stream) {
  const long long total = (long long)batch * (1 + n_reg) * D;
  BSC_LAUNCH(k_fill_prefix_tokens, (unsigned)((total + 255) / 256), 256, 0, stream, x, cls_pos, reg, batch, tokens, n_reg, D);
  return BSCNAV_OK;
}

}  // namespace bscnav

using namespace bscnav;

extern "C" int bscnav_score_views(const float* embeds, const float* goal, int32_t n, int32_t dim, float* out_prob, int32_t* out_best,
                                  void* stream) {
  if (!embeds || !goal || !out_prob || !out_best || n <= 0 || n > 1024 || dim <= 0) return BSCNAV_EINVAL;
  BSC_LAUNCH(k_score_views, 1, 256, 0, (cudaStream_t)stream, embeds, goal, n, dim, out_prob, out_best);
  return BSCNAV_OK;
}

extern "C" int bscnav_gaussian_pool(const float* tokens, int32_t batch, int32_t grid, int32_t dim, float* out, void* stream) {
  if (!tokens || !out || batch <= 0 || grid <= 0 || grid > 64 || dim <= 0) return BSCNAV_EINVAL;
  BSC_LAUNCH(k_gaussian_pool, (dim + 255) / 256, 256, grid * grid * sizeof(float), (cudaStream_t)stream, tokens, batch, grid, dim, out);
  return BSCNAV_OK;
}
