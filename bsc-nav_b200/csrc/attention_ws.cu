// Warp-specialised multi-head self-attention on the 5th-generation tensor cores (sm_100a), head dim 64, 32..272 keys
// (ViT-L/14: T = 257 / 261; ViT-B/32: T = 50).  Replaces the attention inside dinov2.forward_features
// (memory_2.py:591-592, 738): softmax(Q K^T) V per (frame, head), Q pre-scaled by 1/sqrt(64) in the packed weights.
//
// Persistent kernel, one CTA per SM, work item = (frame, head); query tiles of 128 rows.  The keys of a tile are split
// into two halves [0, ka) and [ka, Tk) that run as two independent pipelines, so that the exponentials of one half
// (MUFU-bound) overlap the tensor-core work of the other:
//
//   warp 12      TMA producer   Q tiles (ring of 2), K (2 stages), V (1 stage) through a 3-D tensor map
//                               {column, token-in-frame, frame}: rows past the frame's last token are zero-filled by the
//                               hardware, so K/V/Q tails need no masking pass in shared memory
//   warp 13      MMA issuer     S_x = Q K_x^T  (tcgen05.mma kind::f16, M=128, N=ka|kb, K=64, fp32 in TMEM)
//                               O_x = P_x V_x  (A = P_x from shared memory, B = V MN-major, N=64), x in {a, b}
//   warps 0-3    softmax a      thread = query row = TMEM lane: two passes over S_a with tcgen05.ld (row max, then
//   warps 4-7    softmax b      exp2 / row sum), P_x written as fp16 in the K-major 128B-swizzle layout the MMA reads
//   warps 8-11   epilogue       O = (O_a e^(m_a-m) + O_b e^(m_b-m)) / (l_a e^(m_a-m) + l_b e^(m_b-m)) -> fp16 -> global
//
// Each half keeps its own running maximum and its own accumulator (flash-attention's split-K form), so the two
// pipelines never exchange a maximum.  All hand-offs are full/empty mbarrier pairs (tcgen05.commit on the MMA side).
// Measured (profiles/r01_attention_ws_ncu.json): 70 us for 880 items, tensor pipe 15 %, XU 25 %: the limit is the issue
// rate of the one or two softmax warps each SM sub-partition holds, not MUFU or the tensor pipe.  Forcing the two exp
// passes to alternate (named-barrier ping-pong) was measured slower (78 us) and is not kept.
#include "gemm_tcgen05.cuh"
#include "tc_ptx.cuh"
#include "vit_kernels.cuh"

namespace bscnav {

static constexpr int kWsThreads = 448;
static constexpr int kWsTileQ = 128;
static constexpr int kWsMaxKeys = 272;
static constexpr int kWsQBytes = 16384;                  // 128 rows x 64 fp16
static constexpr int kWsKvBytes = kWsMaxKeys * 128;      // 34816
static constexpr int kWsPBytes = 5 * 16384;              // P: five 64-key atoms of 128 rows x 128 B
static constexpr int kOffQ = 0;
static constexpr int kOffK = 2 * kWsQBytes;              // 32768
static constexpr int kOffV = kOffK + 2 * kWsKvBytes;     // 102400
static constexpr int kOffP = kOffV + kWsKvBytes;         // 137216
static constexpr int kOffStat = kOffP + kWsPBytes;       // 219136: float2 [2 slots][2 halves][128 rows]
static constexpr int kOffBar = kOffStat + 2 * 2 * 128 * 8;
static constexpr int kWsSmem = kOffBar + 256 + 1024;     // + slack for 1024-byte alignment
static constexpr int kColOa = 288, kColOb = 352;         // TMEM columns of the two O accumulators (S occupies [0, 272))

enum WsBar : int {  // byte offsets from the barrier block
  kQFull = 0, kQEmpty = 16, kKFull = 32, kKEmpty = 48, kVFull = 64, kVEmpty = 72, kSFull = 80, kPFull = 96,
  kOFull = 112, kOEmpty = 120, kStatFull = 128, kStatEmpty = 144, kTmemSlot = 160
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// MN-major operand (V as [key][d], d contiguous), 128-byte swizzle: 8-key groups 1024 B apart (SBO); one 64-wide MN atom
__device__ __forceinline__ uint64_t ws_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ float ws_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t ws_pack(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// key split: first half a multiple of 32 (so 32-column softmax chunks never straddle a 64-key swizzle atom), both >= 16
__host__ __device__ inline int ws_split(int Tk) {
  if (Tk <= 32) return 16;
  int ka = ((Tk / 2) + 31) & ~31;
  return ka > 128 ? 128 : ka;
}

__global__ void __launch_bounds__(kWsThreads, 1)
k_attention_ws(const __grid_constant__ CUtensorMap map128, const __grid_constant__ CUtensorMap map16, __half* __restrict__ out,
               int T, int heads, int n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t sQ = sbase + kOffQ, sK = sbase + kOffK, sV = sbase + kOffV, sP = sbase + kOffP;
  float2* stats = reinterpret_cast<float2*>(smem + kOffStat);
  const uint32_t bar = sbase + kOffBar;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + kTmemSlot);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = heads * 64;
  const int n_qt = (T + kWsTileQ - 1) / kWsTileQ;
  const int Tk = (T + 15) & ~15;
  const int ka = ws_split(Tk), kb = Tk - ka;
  const int n_my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_tot = n_my * n_qt;
  const uint32_t kv_bytes = 16384u * (1u + (Tk > 128 ? 1u : 0u)) + (Tk > 256 ? 2048u : 0u);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar + kQFull + 8 * i, 1); mbar_init(bar + kQEmpty + 8 * i, 1);
      mbar_init(bar + kKFull + 8 * i, 1); mbar_init(bar + kKEmpty + 8 * i, 1);
      mbar_init(bar + kSFull + 8 * i, 1); mbar_init(bar + kPFull + 8 * i, 128);
      mbar_init(bar + kStatFull + 8 * i, 256); mbar_init(bar + kStatEmpty + 8 * i, 128);
    }
    mbar_init(bar + kVFull, 1); mbar_init(bar + kVEmpty, 1);
    mbar_init(bar + kOFull, 1); mbar_init(bar + kOEmpty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map128) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map16) : "memory");
  }
  if (warp == 13) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  pdl_wait();                      // the qkv tensor comes from the previous kernel in the stream
  pdl_launch_dependents();

  if (warp == 12) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      auto load_kv = [&](uint32_t dst, uint32_t full, int col, int b) {
        mbar_arrive_expect_tx(full, kv_bytes);
        tma_load_3d(dst, &map128, full, col, 0, b);
        if (Tk > 128) tma_load_3d(dst + 16384, &map128, full, col, 128, b);
        if (Tk > 256) tma_load_3d(dst + 32768, &map16, full, col, 256, b);
      };
      auto load_k = [&](int li, int item) {
        const int st = li & 1;
        mbar_wait(bar + kKEmpty + 8 * st, ((li >> 1) & 1) ^ 1);
        load_kv(sK + st * kWsKvBytes, bar + kKFull + 8 * st, D + (item % heads) * 64, item / heads);
      };
      auto load_q = [&](int n, int item, int m) {
        const int s = n & 1;
        mbar_wait(bar + kQEmpty + 8 * s, ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(bar + kQFull + 8 * s, kWsQBytes);
        tma_load_3d(sQ + s * kWsQBytes, &map128, bar + kQFull + 8 * s, (item % heads) * 64, m * kWsTileQ, item / heads);
      };
      int li = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
        if (li == 0) load_k(0, item);
        const int n0 = li * n_qt;
        load_q(n0, item, 0);
        if (n_qt > 1) load_q(n0 + 1, item, 1);
        mbar_wait(bar + kVEmpty, (li & 1) ^ 1);
        load_kv(sV, bar + kVFull, 2 * D + (item % heads) * 64, item / heads);
        for (int m = 2; m < n_qt; ++m) load_q(n0 + m, item, m);
        if (item + (int)gridDim.x < n_items) load_k(li + 1, item + gridDim.x);
      }
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0 && n_tot > 0) {
      const uint32_t id_sa = umma_idesc_f16(ka), id_sb = umma_idesc_f16(kb), id_o = umma_idesc_f16(64, 128, 1);
      auto issue_s = [&](int x, int n) {
        const int li = n / n_qt;
        const uint64_t a = umma_desc_sw128(sQ + (n & 1) * kWsQBytes);
        const uint64_t b = umma_desc_sw128(sK + (li & 1) * kWsKvBytes + (x ? ka * 128 : 0));
        const uint32_t d = tmem + (x ? ka : 0), id = x ? id_sb : id_sa;
#pragma unroll
        for (int k = 0; k < 4; ++k) tcgen05_mma_f16(d, a + 2 * k, b + 2 * k, id, k ? 1u : 0u);
        tcgen05_commit(bar + kSFull + 8 * x);
      };
      auto issue_o = [&](int x) {
        const int j0 = x ? ka / 16 : 0, j1 = x ? Tk / 16 : ka / 16;
        const uint32_t d = tmem + (x ? kColOb : kColOa);
        for (int j = j0; j < j1; ++j) {
          const uint64_t a = umma_desc_sw128(sP + (j >> 2) * 16384) + 2 * (j & 3);
          const uint64_t v = ws_desc_mn_sw128(sV + j * 2048);
          tcgen05_mma_f16(d, a, v, id_o, j > j0 ? 1u : 0u);
        }
      };
      mbar_wait(bar + kKFull, 0);
      mbar_wait(bar + kQFull, 0);
      tcgen05_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      tcgen05_commit(bar + kQEmpty);
      if (n_qt == 1) tcgen05_commit(bar + kKEmpty);
      for (int n = 0; n < n_tot; ++n) {
        const int li = n / n_qt, m = n - li * n_qt;
        const bool more = n + 1 < n_tot;
        const int li1 = (n + 1) / n_qt, m1 = (n + 1) - li1 * n_qt;
        if (m == 0) mbar_wait(bar + kVFull, li & 1);
        mbar_wait(bar + kPFull, n & 1);
        if (n > 0) mbar_wait(bar + kOEmpty, (n - 1) & 1);
        tcgen05_fence_after();
        issue_o(0);
        if (more) {
          if (m1 == 0) mbar_wait(bar + kKFull + 8 * (li1 & 1), (li1 >> 1) & 1);
          mbar_wait(bar + kQFull + 8 * ((n + 1) & 1), ((n + 1) >> 1) & 1);
          tcgen05_fence_after();
          issue_s(0, n + 1);
        }
        mbar_wait(bar + kPFull + 8, n & 1);
        tcgen05_fence_after();
        issue_o(1);
        tcgen05_commit(bar + kOFull);
        if (m == n_qt - 1) tcgen05_commit(bar + kVEmpty);
        if (more) {
          issue_s(1, n + 1);
          tcgen05_commit(bar + kQEmpty + 8 * ((n + 1) & 1));
          if (m1 == n_qt - 1) tcgen05_commit(bar + kKEmpty + 8 * (li1 & 1));
        }
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------------------------------ softmax pipelines a / b
    const int x = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const int c_lo = x ? ka : 0, c_hi = x ? Tk : ka;
    const float kLog2e = 1.4426950408889634f;
    const uint32_t p_row = sP + r * 128;
    const int sw = r & 7;
    for (int n = 0; n < n_tot; ++n) {
      const int m = n % n_qt, slot = n & 1;
      const bool live = m * kWsTileQ + quarter * 32 < T;
      mbar_wait(bar + kSFull + 8 * x, n & 1);
      tcgen05_fence_after();
      float mx = -INFINITY, sum = 0.f;
      if (live) {
        // pass 1: row maximum of this half (keys >= T are zero-filled rows of K: excluded)
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int c = c_lo; c < c_hi;) {
          uint32_t v[32];
          const bool wide = (c & 31) == 0 && c + 32 <= c_hi;
          if (wide) tmem_ld_32x32(lane_addr + c, v); else tmem_ld_32x16(lane_addr + c, v);
          tmem_ld_wait();
          const int w = wide ? 32 : 16;
          if (c + w <= T) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < 16 || wide) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if ((j < 16 || wide) && c + j < T) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j]));
          }
          c += w;
        }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        // pass 2: p = exp(s - max) as exp2, row sum in fp32, P as fp16 into the swizzled K-major operand tile
        const float mo = mx * kLog2e;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = c_lo; c < c_hi;) {
          uint32_t v[32];
          const bool wide = (c & 31) == 0 && c + 32 <= c_hi;
          if (wide) tmem_ld_32x32(lane_addr + c, v); else tmem_ld_32x16(lane_addr + c, v);
          tmem_ld_wait();
          const int w = wide ? 32 : 16;
          float p[32];
          if (c + w <= T) {
#pragma unroll
            for (int j = 0; j < 32; ++j) p[j] = (j < 16 || wide) ? ws_ex2(fmaf(__uint_as_float(v[j]), kLog2e, -mo)) : 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              p[j] = ((j < 16 || wide) && c + j < T) ? ws_ex2(fmaf(__uint_as_float(v[j]), kLog2e, -mo)) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) s4[j & 3] += (p[4 * j] + p[4 * j + 1]) + (p[4 * j + 2] + p[4 * j + 3]);
          const uint32_t atom = p_row + (c >> 6) * 16384;
          const int cb = (c & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < 2 || wide)
              sts128(atom + (((cb + j) ^ sw) << 4), ws_pack(p[8 * j], p[8 * j + 1]), ws_pack(p[8 * j + 2], p[8 * j + 3]),
                     ws_pack(p[8 * j + 4], p[8 * j + 5]), ws_pack(p[8 * j + 6], p[8 * j + 7]));
          }
          c += w;
        }
        sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      }
      mbar_wait(bar + kStatEmpty + 8 * slot, ((n >> 1) & 1) ^ 1);       // the epilogue has read this slot's previous stats
      stats[(slot * 2 + x) * 128 + r] = make_float2(mx, sum);
      fence_proxy_async();                                               // P (generic-proxy stores) -> async proxy (MMA)
      tcgen05_fence_before();
      mbar_arrive(bar + kPFull + 8 * x);
      mbar_arrive(bar + kStatFull + 8 * slot);
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------------------------------ epilogue
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const float kLog2e = 1.4426950408889634f;
    for (int n = 0; n < n_tot; ++n) {
      const int li = n / n_qt, m = n - li * n_qt, slot = n & 1;
      const int item = blockIdx.x + li * gridDim.x;
      const int b = item / heads, head = item - b * heads;
      const bool live = m * kWsTileQ + quarter * 32 < T;
      mbar_wait(bar + kStatFull + 8 * slot, (n >> 1) & 1);
      mbar_wait(bar + kOFull, n & 1);
      tcgen05_fence_after();
      if (live) {
        const float2 sa = stats[(slot * 2 + 0) * 128 + r], sb = stats[(slot * 2 + 1) * 128 + r];
        const float mm = fmaxf(sa.x, sb.x);
        float wa = ws_ex2((sa.x - mm) * kLog2e), wb = ws_ex2((sb.x - mm) * kLog2e);
        const float inv = 1.f / (sa.y * wa + sb.y * wb);
        wa *= inv; wb *= inv;
        const int q_row = m * kWsTileQ + r;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t oa[32], ob[32];
          tmem_ld_32x32(lane_addr + kColOa + 32 * h, oa);
          tmem_ld_32x32(lane_addr + kColOb + 32 * h, ob);
          tmem_ld_wait();
          if (q_row < T) {
            uint4* dst = reinterpret_cast<uint4*>(out + ((long long)b * T + q_row) * D + head * 64 + 32 * h);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = fmaf(__uint_as_float(oa[8 * j + e]), wa, __uint_as_float(ob[8 * j + e]) * wb);
              uint4 u;
              u.x = ws_pack(o[0], o[1]); u.y = ws_pack(o[2], o[3]); u.z = ws_pack(o[4], o[5]); u.w = ws_pack(o[6], o[7]);
              dst[j] = u;
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(bar + kOEmpty);
      mbar_arrive(bar + kStatEmpty + 8 * slot);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 13) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

bool attention_ws_supported(int T) { const int Tk = (T + 15) & ~15; return T >= 17 && Tk <= kWsMaxKeys; }

int attention_ws_launch(const CUtensorMap* map128, const CUtensorMap* map16, __half* out, int batch, int T, int heads,
                        cudaStream_t stream) {
  if (!attention_ws_supported(T) || batch <= 0 || heads <= 0) return BSCNAV_EINVAL;
  static bool attr_set = false;
  if (!attr_set) {
    BSC_CUDA(cudaFuncSetAttribute(k_attention_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmem));
    attr_set = true;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int n_items = batch * heads;
  const int grid = n_items < sms ? n_items : sms;
  BSC_LAUNCH_PDL(k_attention_ws, grid, kWsThreads, kWsSmem, stream, *map128, *map16, out, T, heads, n_items);
  return BSCNAV_OK;
}

}  // namespace bscnav
