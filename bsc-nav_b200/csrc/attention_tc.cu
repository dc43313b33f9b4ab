// Multi-head self-attention on the 5th-generation tensor cores (sm_100a), head dim 64, up to 272 keys (ViT-L/14: 257/261).
//
// One CTA per (frame, head).  Q (all query tiles), K and V of the head are staged once by TMA (128-byte swizzle).
// Per 128-query tile:
//   S = Q K^T      tcgen05.mma kind::f16, M=128, N=256 (+16), K=64, fp32 accumulator in TMEM columns [0, 272)
//   softmax        8 warps read their rows of S with tcgen05.ld (thread = row, two warps split the columns), fp32 max / exp2 /
//                  sum, P written as fp16 into shared memory in the K-major 128B-swizzle layout the next MMA consumes
//   O = P V        tcgen05.mma, A = P (shared, K-major), B = V (shared, MN-major: V is [key][d] as it sits in the qkv tensor),
//                  fp32 accumulator in TMEM columns [288, 352)
//   epilogue       tcgen05.ld, scale by 1/rowsum, fp16, 64-byte stores
// S of tile m+1 is issued right behind O of tile m, so the tensor pipe works while the CUDA cores run the epilogue.
// The mma.sync kernel in vit_kernels.cu stays as the fallback for longer sequences and as the cross-check in the tests.
#include "gemm_tcgen05.cuh"
#include "tc_ptx.cuh"
#include "vit_kernels.cuh"

namespace bscnav {

static constexpr int kTcThreads = 256;
static constexpr int kTileQ = 128;
static constexpr int kMaxKeys = 272;                 // 256 + 16: one UMMA N=256 and one N=16
static constexpr int kKvBytes = kMaxKeys * 128;      // K (or V) tile: rows of 64 fp16
static constexpr int kPBytes = 5 * 16384;            // P: five 64-key atoms of 128 rows x 128 B
static constexpr int kColO = 288;                    // TMEM column of the O accumulator

// MN-major operand (V as [key][d], d contiguous), 128-byte swizzle: 8-key groups 1024 B apart (SBO); one 64-wide MN atom
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kTcThreads, 1)
k_attention_tc(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map16, __half* __restrict__ out,
               int T, int heads, int n_qt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int D = heads * 64;
  const int head = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + n_qt * 16384;
  const uint32_t sV = sK + kKvBytes;
  const uint32_t sP = sV + kKvBytes;
  uint8_t* tail = smem + n_qt * 16384 + 2 * kKvBytes + kPBytes;
  const uint32_t load_bar = smem_u32(tail), s_bar = load_bar + 8, o_bar = load_bar + 16, v_bar = load_bar + 24;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 32);
  float* red_max = reinterpret_cast<float*>(tail + 64);      // [2][128]
  float* red_sum = red_max + 256;                             // [2][128]

  const int Tk = (T + 15) & ~15;                              // keys rounded up to the UMMA K step
  const int n1 = Tk < 256 ? Tk : 256, n2 = Tk - n1;           // S = [N=n1 | N=n2] columns

  if (tid == 0) {
    mbar_init(load_bar, 1); mbar_init(s_bar, 1); mbar_init(o_bar, 1); mbar_init(v_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map128) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map16) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int row0 = b * T;                                      // first row of this frame in the qkv tensor
  if (tid == 0) {
    // first barrier: what S of the first tile needs (Q tile 0 and K); second: V and the remaining Q tiles
    mbar_arrive_expect_tx(load_bar, 16384u + 32768u + (n2 ? 2048u : 0u));
    tma_load_2d(sQ, &map128, load_bar, head * 64, row0);
    tma_load_2d(sK, &map128, load_bar, D + head * 64, row0);
    tma_load_2d(sK + 16384, &map128, load_bar, D + head * 64, row0 + 128);
    if (n2) tma_load_2d(sK + 32768, &map16, load_bar, D + head * 64, row0 + 256);
    mbar_arrive_expect_tx(v_bar, (uint32_t)(n_qt - 1) * 16384u + 32768u + (n2 ? 2048u : 0u));
    tma_load_2d(sV, &map128, v_bar, 2 * D + head * 64, row0);
    tma_load_2d(sV + 16384, &map128, v_bar, 2 * D + head * 64, row0 + 128);
    if (n2) tma_load_2d(sV + 32768, &map16, v_bar, 2 * D + head * 64, row0 + 256);
    for (int m = 1; m < n_qt; ++m) tma_load_2d(sQ + m * 16384, &map128, v_bar, head * 64, row0 + m * 128);
    mbar_wait(load_bar, 0);
  }

  auto issue_s = [&](int m) {        // S = Q_m K^T
    const uint64_t a = umma_desc_sw128(sQ + m * 16384);
    const uint64_t k1 = umma_desc_sw128(sK), k2 = umma_desc_sw128(sK + 32768);
    const uint32_t id1 = umma_idesc_f16(n1), id2 = umma_idesc_f16(16);
#pragma unroll
    for (int k = 0; k < 4; ++k) tcgen05_mma_f16(tmem, a + 2 * k, k1 + 2 * k, id1, k ? 1u : 0u);
    if (n2) {
#pragma unroll
      for (int k = 0; k < 4; ++k) tcgen05_mma_f16(tmem + 256, a + 2 * k, k2 + 2 * k, id2, k ? 1u : 0u);
    }
    tcgen05_commit(s_bar);
  };
  auto issue_o = [&]() {             // O = P V
    const uint32_t id = umma_idesc_f16(64, 128, 1);
    for (int k = 0; k < Tk / 16; ++k) {
      const uint64_t a = umma_desc_sw128(sP + (k >> 2) * 16384) + 2 * (k & 3);
      const uint64_t v = umma_desc_mn_sw128(sV + k * 2048);
      tcgen05_mma_f16(tmem + kColO, a, v, id, k ? 1u : 0u);
    }
    tcgen05_commit(o_bar);
  };

  const int quarter = warp & 3, half = warp >> 2;
  const int r_tile = quarter * 32 + lane;                      // row inside the 128-query tile == TMEM lane
  const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
  const int c_lo = half ? 128 : 0, c_hi = half ? Tk : (Tk < 128 ? Tk : 128);   // this warp's key columns
  const float kLog2e = 1.4426950408889634f;
  const uint32_t p_row = sP + r_tile * 128;
  const int sw = r_tile & 7;

  if (tid == 0) { tcgen05_fence_after(); issue_s(0); }
  for (int m = 0; m < n_qt; ++m) {
    const bool warp_live = m * kTileQ + quarter * 32 < T;      // warps whose 32 rows are all past the end only keep the barriers
    mbar_wait(s_bar, m & 1);
    tcgen05_fence_after();
    // ---- pass 1: row maximum over this warp's columns (four independent chains; masking only in the chunk that crosses T)
    float mx = -INFINITY;
    if (warp_live) {
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      for (int c = c_lo; c < c_hi; c += 32) {
        uint32_t r[32];
        const bool full = c + 32 <= c_hi;
        if (full) tmem_ld_32x32(lane_addr + c, r); else tmem_ld_32x16(lane_addr + c, r);
        tmem_ld_wait();
        if (c + 32 <= T) {
#pragma unroll
          for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(r[j]));
        } else {
          const int nv = full ? 32 : 16;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nv && c + j < T) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(r[j]));
        }
      }
      mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    }
    red_max[half * 128 + r_tile] = mx;
    __syncthreads();
    mx = fmaxf(red_max[r_tile], red_max[128 + r_tile]);
    // ---- pass 2: p = exp(s - max), row sum, P (fp16) into shared memory in the UMMA K-major swizzled layout
    float sum = 0.f;
    if (warp_live) {
      const float mo = mx * kLog2e;
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c = c_lo; c < c_hi; c += 32) {
        uint32_t r[32];
        const bool full = c + 32 <= c_hi;
        if (full) tmem_ld_32x32(lane_addr + c, r); else tmem_ld_32x16(lane_addr + c, r);
        tmem_ld_wait();
        const int nv = full ? 32 : 16;
        float p[32];
        if (c + 32 <= T) {
#pragma unroll
          for (int j = 0; j < 32; ++j) p[j] = ex2_fast(fmaf(__uint_as_float(r[j]), kLog2e, -mo));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) p[j] = (j < nv && c + j < T) ? ex2_fast(fmaf(__uint_as_float(r[j]), kLog2e, -mo)) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s4[j & 3] += (p[4 * j] + p[4 * j + 1]) + (p[4 * j + 2] + p[4 * j + 3]);
        const uint32_t atom = p_row + (c >> 6) * 16384;
        const int cb = (c & 63) >> 3;                           // first 16-byte chunk of this 32-column group in the atom row
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j * 8 < nv)
            sts128(atom + (((cb + j) ^ sw) << 4), pack_h2(p[8 * j], p[8 * j + 1]), pack_h2(p[8 * j + 2], p[8 * j + 3]),
                   pack_h2(p[8 * j + 4], p[8 * j + 5]), pack_h2(p[8 * j + 6], p[8 * j + 7]));
        }
      }
      sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    }
    red_sum[half * 128 + r_tile] = sum;
    if (m == 0) {
      // V (and the later Q tiles) landed behind the first softmax.  V rows in [T, Tk) take part in O = P V with P = 0:
      // they must be finite, and the tensor's next rows may hold anything -> zero them.
      mbar_wait(v_bar, 0);
      for (int i = tid; i < (Tk - T) * 8; i += kTcThreads) sts128(sV + (T + (i >> 3)) * 128 + ((i & 7) << 4), 0u, 0u, 0u, 0u);
    }
    fence_proxy_async();                                        // P (generic-proxy stores) -> visible to the tensor core's async proxy
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      issue_o();
      if (m + 1 < n_qt) issue_s(m + 1);                          // S of the next tile runs behind O of this one
    }
    // ---- epilogue: O / rowsum -> fp16 -> global
    mbar_wait(o_bar, m & 1);
    tcgen05_fence_after();
    const int q_row = m * kTileQ + r_tile;
    if (warp_live) {
      uint32_t r[32];
      tmem_ld_32x32(lane_addr + kColO + half * 32, r);
      tmem_ld_wait();
      if (q_row < T) {
        const float inv = 1.f / (red_sum[r_tile] + red_sum[128 + r_tile]);
        uint4* dst = reinterpret_cast<uint4*>(out + (long long)(row0 + q_row) * D + head * 64 + half * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          u.x = pack_h2(__uint_as_float(r[8 * j]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
          u.y = pack_h2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
          u.z = pack_h2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
          u.w = pack_h2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
          dst[j] = u;
        }
      }
    }
    tcgen05_fence_before();
    __syncthreads();                                            // O and the reduction scratch are free for the next tile
  }
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

bool attention_tc_supported(int T) { return T >= 1 && ((T + 15) & ~15) <= kMaxKeys; }

int attention_tc_launch(const CUtensorMap* map128, const CUtensorMap* map16, __half* out, int batch, int T, int heads,
                        cudaStream_t stream) {
  if (!attention_tc_supported(T) || batch <= 0 || heads <= 0) return BSCNAV_EINVAL;
  const int n_qt = (T + kTileQ - 1) / kTileQ;
  const int smem = n_qt * 16384 + 2 * kKvBytes + kPBytes + 64 + 4 * 128 * 4 + 1024;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    BSC_CUDA(cudaFuncSetAttribute(k_attention_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  BSC_LAUNCH(k_attention_tc, dim3((unsigned)heads, (unsigned)batch), kTcThreads, smem, stream, *map128, *map16, out, T, heads, n_qt);
  return BSCNAV_OK;
}

}  // namespace bscnav

using namespace bscnav;

// Stand-alone attention entry (tests): qkv fp16 [batch*T, 3*heads*64] with q pre-scaled -> out fp16 [batch*T, heads*64].
// impl 0 = mma.sync kernel, 1 = tcgen05 kernel.
extern "C" int bscnav_attention_f16(const void* qkv, void* out, int32_t batch, int32_t T, int32_t heads, int32_t impl, void* stream) {
  if (!qkv || !out || batch <= 0 || T <= 0 || heads <= 0) return BSCNAV_EINVAL;
  if (impl == 0) return attention_f16(static_cast<const __half*>(qkv), static_cast<__half*>(out), batch, T, heads, 64, false, (cudaStream_t)stream);
  CUtensorMap m128, m16;
  if (impl == 3) {
    CUtensorMap mo;
    if (int e = make_qkv_map3d(&m128, qkv, batch, T, 3LL * heads * 64, 128)) return e;
    if (int e = make_qkv_map3d(&m16, qkv, batch, T, 3LL * heads * 64, 16)) return e;
    if (int e = make_out_map3d(&mo, out, batch, T, 1LL * heads * 64)) return e;
    return attention_ws2_launch(&m128, &m16, &mo, static_cast<const __half*>(qkv), static_cast<__half*>(out), batch, T, heads,
                                (cudaStream_t)stream);
  }
  if (impl == 2) {
    if (int e = make_qkv_map3d(&m128, qkv, batch, T, 3LL * heads * 64, 128)) return e;
    if (int e = make_qkv_map3d(&m16, qkv, batch, T, 3LL * heads * 64, 16)) return e;
    return attention_ws_launch(&m128, &m16, static_cast<__half*>(out), batch, T, heads, (cudaStream_t)stream);
  }
  if (int e = make_operand_map(&m128, qkv, (long long)batch * T, 3LL * heads * 64, 128)) return e;
  if (int e = make_operand_map(&m16, qkv, (long long)batch * T, 3LL * heads * 64, 16)) return e;
  return attention_tc_launch(&m128, &m16, static_cast<__half*>(out), batch, T, heads, (cudaStream_t)stream);
}

// The mma.sync kernel with an explicit head dim (64 | 80; 80 = ViT-H/14) and optional causal mask (head dim 64, text tower).
extern "C" int bscnav_attention_f16_hd(const void* qkv, void* out, int32_t batch, int32_t T, int32_t heads, int32_t head_dim,
                                       int32_t causal, void* stream) {
  if (!qkv || !out || batch <= 0 || T <= 0 || heads <= 0) return BSCNAV_EINVAL;
  if (causal == 2) {      // causal mask on the second-generation tcgen05 kernel (head dim 64, 17 <= T <= 272): what the text tower runs
    if (head_dim != 64 || !attention_ws2_supported(T)) return BSCNAV_EINVAL;
    CUtensorMap m128, m16, mo;
    if (int e = make_qkv_map3d(&m128, qkv, batch, T, 3LL * heads * 64, 128)) return e;
    if (int e = make_qkv_map3d(&m16, qkv, batch, T, 3LL * heads * 64, 16)) return e;
    if (int e = make_out_map3d(&mo, out, batch, T, 1LL * heads * 64)) return e;
    return attention_ws2_launch(&m128, &m16, &mo, static_cast<const __half*>(qkv), static_cast<__half*>(out), batch, T, heads,
                                (cudaStream_t)stream, true);
  }
  return attention_f16(static_cast<const __half*>(qkv), static_cast<__half*>(out), batch, T, heads, head_dim, causal != 0,
                       (cudaStream_t)stream);
}
