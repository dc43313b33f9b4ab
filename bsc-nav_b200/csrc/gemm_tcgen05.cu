// tcgen05 / TMEM / TMA GEMM for sm_100a -- see gemm_tcgen05.cuh for the design.
#include <cstdlib>
#include "gemm_tcgen05.cuh"
#include <cuda_bf16.h>
#include "tc_ptx.cuh"

namespace bscnav {

static constexpr int BM = 128;           // UMMA M (cta_group::1)
static constexpr int BK = 64;            // one 128-byte swizzle atom of fp16 along K
static constexpr int UK = 16;            // UMMA K for 16-bit operands
static constexpr int kEpiWarps = 8;
static constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
static constexpr int kTmaWarp = kEpiWarps;          // the scheduler favours the highest warp id: the single-thread
static constexpr int kMmaWarp = kEpiWarps + 1;      // MMA issuer must never wait behind ALU-bound epilogue warps

template <int BN, int CL = 1> struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / CL) * BK * 2;                        // CL == 2: this CTA's half of the weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256 && CL == 1) ? 4 : 6;
  static constexpr int kBarBytes = 256;
  static constexpr int kStagingBytes = kEpiWarps * 4096;                    // per epilogue warp: 32 rows x 128 B (TMA store box)
  static constexpr int kSmem = kStages * kStageBytes + kStagingBytes + kBarBytes + 1024;   // + slack for 1024-byte alignment
  static constexpr int kTmemCols = 2 * BN;                                  // two accumulator stages
};

// activation selected at compile time: evaluating both and selecting cost 35 instructions per element
template <int ACT> __device__ __forceinline__ float act_fn(float v) {
  if (ACT == 0) {
    // GELU (erf form, DINOv2): erf by Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7 (the result is rounded to fp16, ulp 2^-11
    // relative): two MUFU ops and a five-term Horner chain instead of erff's ~30-instruction branchy body -- the fc1
    // epilogue is what bounds that GEMM (1020 vs 1320 TFLOP/s with QuickGELU)
    const float x = fabsf(v) * 0.70710678118654752440f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, x, 1.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -1.4426950408889634f));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float erf_abs = fmaf(-p * t, e, 1.0f);                 // erf(|v| / sqrt 2)
    return 0.5f * v + 0.5f * fabsf(v) * erf_abs;                 // 0.5 v (1 + sign(v) erf(|v|/sqrt 2))
  }
  // QuickGELU v * sigmoid(1.702 v) = v / (1 + 2^(-1.702 log2(e) v)): two MUFU ops, |rel err| ~1e-6 << fp16 ulp
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -2.4554669595930157f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return v * r;
}

// CL = 2: the CTA pair of a cluster computes two vertically adjacent output tiles (same weight panel) as ONE 256 x BN MMA
// (tcgen05.mma.cta_group::2).  Each CTA keeps its own 128 rows of A and only HALF of the weight tile in shared memory, the
// leader (rank 0) issues the MMAs for both, each CTA's TMEM receives its own 128 x BN accumulator and runs its own epilogue.
// Why: at cta_group::1 a 128 x 256 x 64 k-block moves 48 KB into shared memory (TMA) and 48 KB out of it (MMA operands) per
// 512 tensor-pipe cycles, ~190 B/clk against ~128 B/clk of shared-memory bandwidth -- the tensor pipe measured 70 % active.
// The pair form moves 32 KB in and 32 KB out per SM.  (Multicasting the weight tile at cta_group::1 was measured: it only
// relieves L2, and was not faster.)
template <int BN, int MODE, int ACT, int CL>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_f16(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
           const __grid_constant__ CUtensorMap map_out, int M, int N, int K, const GemmEpilogue epi, int split) {
  using C = Cfg<BN, CL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bars = smem + C::kStages * C::kStageBytes + C::kStagingBytes;
  const uint32_t full_bar = smem_u32(bars);                       // kStages x 8 B
  const uint32_t empty_bar = full_bar + 8 * C::kStages;           // kStages x 8 B
  const uint32_t tfull_bar = empty_bar + 8 * C::kStages;          // 2 x 8 B
  const uint32_t tempty_bar = tfull_bar + 16;                     // 2 x 8 B
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 * (2 * C::kStages + 4));
  volatile int* kflag = reinterpret_cast<volatile int*>(bars + 8 * (2 * C::kStages + 4) + 8);
  const uint32_t smem_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kbk = K / BK;                                   // k-blocks of one pass; split: three passes (hi hi, lo hi, hi lo)
  const int m_tiles = (M + BM - 1) / BM, n_tiles = N / BN, k_blocks = split ? 3 * kbk : kbk;
  const int m_groups = (m_tiles + CL - 1) / CL;             // CL == 2: pairs of row blocks (the last pair may be half empty)
  const int total_tiles = m_groups * n_tiles;
  // K-sliced launches (small M, CL == 1 only): a work item is (tile, slice); the slices of a tile meet in the epilogue (below)
  const int S = (CL == 1 && BN == 128 && (MODE == kOutResidual || MODE == kOutFloat) && epi.kslices > 1) ? epi.kslices : 1;
  const int total_items = total_tiles * S;
  const int rank = CL == 2 ? (int)cluster_ctarank() : 0;
  const int tile0 = blockIdx.x / CL, tile_step = gridDim.x / CL;
  // Tile order.  m-fastest: the ~148 resident tiles share one weight panel (good when A fits L2, K <= N).
  // n-fastest: the n_tiles tiles of one row block run side by side and share the A panel, so a long-K activation
  // matrix (fc2: 116 MB) streams from HBM once instead of once per weight panel (measured 413 MB -> A-size reads).
  const bool n_fastest = K > N;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    // CL == 2: the leader's accumulator-empty barrier collects the epilogue threads of both CTAs
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, CL * 32 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == kMmaWarp) {
    if (CL == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();        // the peer's barriers are initialised before anything is multicast to them
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  // everything above touched only this CTA's shared memory / TMEM: from here on the previous kernel's results are needed
  pdl_wait();
  pdl_launch_dependents();

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = tile0; item < total_items; item += tile_step) {
        const int ks = S > 1 ? item / total_tiles : 0, tile = item - ks * total_tiles;
        const int k_lo = S > 1 ? ks * k_blocks / S : 0, k_hi = S > 1 ? (ks + 1) * k_blocks / S : k_blocks;
        const int m_blk = (n_fastest ? tile / n_tiles : tile % m_groups) * CL + rank, n_blk = n_fastest ? tile % n_tiles : tile / m_groups;
        for (int kb = k_lo; kb < k_hi; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);       // CL == 2: released by the leader's commit in both CTAs
          const uint32_t a_dst = smem_base + stage * C::kStageBytes;
          const uint32_t b_dst = a_dst + C::kABytes;
          // split operands: pass 0 a_hi w_hi, pass 1 a_lo w_hi, pass 2 a_hi w_lo (lo halves start at column K)
          const int pass = split ? kb / kbk : 0, kj = split ? kb - pass * kbk : kb;
          const int a_col = kj * BK + (pass == 1 ? K : 0), w_col = kj * BK + (pass == 2 ? K : 0);
          if (CL == 2) {
            // bytes of both CTAs complete on the leader's barrier; the leader expects the pair's total
            const uint32_t full_leader = mapa_shared(full_bar + 8 * stage, 0);
            if (rank == 0) mbar_arrive_expect_tx(full_bar + 8 * stage, 2 * C::kStageBytes);
            tma_load_2d_2sm(a_dst, &map_a, full_leader, a_col, m_blk * BM);                // rows past M arrive as zeros
            tma_load_2d_2sm(b_dst, &map_w, full_leader, w_col, n_blk * BN + rank * (BN / 2));   // map_w: BN/2-row box
          } else {
            mbar_arrive_expect_tx(full_bar + 8 * stage, C::kStageBytes);
            tma_load_2d(a_dst, &map_a, full_bar + 8 * stage, a_col, m_blk * BM);
            tma_load_2d(b_dst, &map_w, full_bar + 8 * stage, w_col, n_blk * BN);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (one thread; CL == 2: of the leader CTA)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_f16(BN, CL == 2 ? 256 : 128) | (split ? kUmmaBf16Operands : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = tile0; item < total_items; item += tile_step) {
        const int ks = S > 1 ? item / total_tiles : 0;
        const int k_lo = S > 1 ? ks * k_blocks / S : 0, k_hi = S > 1 ? (ks + 1) * k_blocks / S : k_blocks;
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = k_lo; kb < k_hi; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_base + stage * C::kStageBytes;
          const uint64_t a_desc = umma_desc_sw128(a_addr);
          const uint64_t b_desc = umma_desc_sw128(a_addr + C::kABytes);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            // advance 32 bytes (16 fp16) inside the swizzle atom: +2 in the (addr >> 4) field
            if (CL == 2) tcgen05_mma_f16_2sm(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, ((kb - k_lo) | k) != 0 ? 1u : 0u);
            else tcgen05_mma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, ((kb - k_lo) | k) != 0 ? 1u : 0u);
          }
          if (CL == 2) tcgen05_commit_2sm(empty_bar + 8 * stage, 3);   // both CTAs' stages are free once these MMAs have read them
          else tcgen05_commit(empty_bar + 8 * stage);   // smem stage reusable once these MMAs have read it
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (CL == 2) tcgen05_commit_2sm(tfull_bar + 8 * acc, 3);   // accumulators complete in both CTAs -> both epilogues
        else tcgen05_commit(tfull_bar + 8 * acc);        // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> registers -> global
    // 8 warps: TMEM lane quarter = warp % 4 (hardware restriction), column half = warp / 4.
    // Output tiles leave through shared memory: each warp stages 32 rows x 128 B in the TMA 128-byte-swizzle layout
    // (conflict-free 16-byte st.shared), then one lane issues a bulk tensor store -- or, for the fp32 residual stream,
    // a bulk tensor reduce-add performed at L2 (one contributor per element per launch: deterministic).
    // MODE is a template parameter and the column loop is not unrolled: the epilogue stays small enough for the
    // instruction cache (an unrolled multi-mode body spent a quarter of its samples in stall_no_inst).
    constexpr bool kHalfOut = MODE == kOutHalf || MODE == kOutHalfAct || MODE == kOutSims;
    constexpr bool kViaTma = kHalfOut || MODE == kOutResidual;
    constexpr int kCols = BN / 2;                        // columns handled by one warp
    constexpr int kIter = kHalfOut ? 64 : 32;            // columns per iteration == one 128-byte TMA box row
    constexpr int kLd = kIter / 32;                      // TMEM loads in flight per iteration
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const uint32_t stg = smem_base + C::kStages * C::kStageBytes + warp * 4096;
    const uint32_t row_base = stg + lane * 128;
    const int sw = lane & 7;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t range_chk = 0;        // fp16 outputs: bit 15 of a half-word is set when its exponent field is all ones (inf / nan)
    for (int item = tile0; item < total_items; item += tile_step) {
      const int ks = S > 1 ? item / total_tiles : 0, tile = item - ks * total_tiles;
      const int m_blk = (n_fastest ? tile / n_tiles : tile % m_groups) * CL + rank, n_blk = n_fastest ? tile % n_tiles : tile / m_groups;
      const int m0 = m_blk * BM + quarter * 32;
      const int m = m0 + lane;
      const bool in_m = m < M;
      long long out_row = m;
      const float* table_row = nullptr;
      if (MODE == kOutPatchEmbed && in_m) {
        const int f = m / epi.np, p = m - f * epi.np;
        out_row = (long long)f * epi.tokens + epi.row_off + p;
        table_row = epi.table + (long long)p * N;
      }
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tcgen05_fence_after();
      // K slices: every slice parks its raw fp32 accumulators in the launch's scratch (slot = (slice, tile); a thread's 32 columns
      // as eight float4 interleaved over the lanes: coalesced, private layout), then counts itself on the tile's counter.  The
      // slice that arrives last sums all slots IN SLICE ORDER -- the result does not depend on the arrival order -- and runs the
      // ordinary epilogue on the sums; it also re-arms the counter for the next launch.
      constexpr int kChunks = kCols / 32;
      constexpr bool kSliceable = CL == 1 && BN == 128 && kLd == 1;     // the small-M path; fp32 outputs (the residual GEMMs are the long-K ones)
      constexpr long long kSlotF4 = (long long)BM * BN / 4;
      bool from_ws = false;
      const float4* ws_rd = nullptr;
      if (kSliceable && S > 1) {
        float4* slot = reinterpret_cast<float4*>(epi.kws) + ((long long)ks * total_tiles + tile) * kSlotF4 + (long long)warp * kChunks * 256 + lane;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          uint32_t r0[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * kCols + 32 * c), r0);
          tmem_ld_wait();
          if (c == kChunks - 1) { tcgen05_fence_before(); mbar_arrive(tempty_bar + 8 * acc); }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            __stcg(slot + (c * 8 + j) * 32, make_float4(__uint_as_float(r0[4 * j]), __uint_as_float(r0[4 * j + 1]),
                                                         __uint_as_float(r0[4 * j + 2]), __uint_as_float(r0[4 * j + 3])));
        }
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        if (threadIdx.x == 0) {
          __threadfence();                                 // release: cumulative over the other epilogue threads' stores (ordered by the barrier)
          const int last = atomicAdd(epi.kcnt + tile, 1) == S - 1;
          __threadfence();                                 // acquire side of the same counter
          if (last) epi.kcnt[tile] = 0;
          *kflag = last;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        if (!*kflag) {
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          continue;
        }
        __threadfence();
        from_ws = true;
        ws_rd = reinterpret_cast<const float4*>(epi.kws) + (long long)tile * kSlotF4 + (long long)warp * kChunks * 256 + lane;
      }
#pragma unroll 1
      for (int it = 0; it < kCols / kIter; ++it) {
        const int col = half * kCols + it * kIter;
        const int n0 = n_blk * BN + col;
        uint32_t r[kLd][32];
        if (kSliceable && from_ws) {
          // kLd == 1.  Two slots' loads in flight at a time (the reads are L2 round trips: one slot at a time cost ~0.7 us each)
          const float4* p = ws_rd + it * 256;
          const long long stride = (long long)total_tiles * kSlotF4;
#pragma unroll
          for (int j = 0; j < 32; ++j) r[0][j] = 0u;     // 0.0f: the sum starts from slice 0 exactly (0 + a == a)
#pragma unroll 1
          for (int s0 = 0; s0 < S; s0 += 2) {
            float4 t[2][8];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              if (s0 + g < S) {
#pragma unroll
                for (int j = 0; j < 8; ++j) t[g][j] = __ldcg(p + (s0 + g) * stride + j * 32);
              }
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              if (s0 + g < S) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  r[0][4 * j] = __float_as_uint(__uint_as_float(r[0][4 * j]) + t[g][j].x);
                  r[0][4 * j + 1] = __float_as_uint(__uint_as_float(r[0][4 * j + 1]) + t[g][j].y);
                  r[0][4 * j + 2] = __float_as_uint(__uint_as_float(r[0][4 * j + 2]) + t[g][j].z);
                  r[0][4 * j + 3] = __float_as_uint(__uint_as_float(r[0][4 * j + 3]) + t[g][j].w);
                }
              }
            }
          }
        } else {
#pragma unroll
          for (int u = 0; u < kLd; ++u)
            tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + col + 32 * u), r[u]);
        }
        float4 bq[kLd][8];
        if (epi.bias) {                                   // in flight together with the TMEM loads
          const float4* b4 = reinterpret_cast<const float4*>(epi.bias + n0);
#pragma unroll
          for (int j = 0; j < 8 * kLd; ++j) bq[j / 8][j % 8] = __ldg(b4 + j);
        } else {
#pragma unroll
          for (int j = 0; j < 8 * kLd; ++j) bq[j / 8][j % 8] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait();
        if (it == kCols / kIter - 1 && !from_ws) {        // accumulator fully read: hand the TMEM stage back to the MMA warp
          tcgen05_fence_before();
          if (CL == 2) mbar_arrive_cluster(mapa_shared(tempty_bar + 8 * acc, 0));   // the MMA issuer lives in the leader CTA
          else mbar_arrive(tempty_bar + 8 * acc);
        }
        if (kViaTma) {                                    // staging buffer free again? (previous bulk op has read it)
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
        }
#pragma unroll
        for (int u = 0; u < kLd; ++u) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[4 * j] = __uint_as_float(r[u][4 * j]) + bq[u][j].x; v[4 * j + 1] = __uint_as_float(r[u][4 * j + 1]) + bq[u][j].y;
            v[4 * j + 2] = __uint_as_float(r[u][4 * j + 2]) + bq[u][j].z; v[4 * j + 3] = __uint_as_float(r[u][4 * j + 3]) + bq[u][j].w;
          }
          if (kHalfOut) {
            if (MODE == kOutHalfAct) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = act_fn<ACT>(v[j]);
            }
            if (MODE == kOutSims) {
              // batched queries: columns are store cells; cells that take no part become -inf, and the maximum of every
              // 32-cell block is kept in fp32 (the selection works on these 32x fewer numbers, csrc/query_batch.cu)
              const uint4* mk = reinterpret_cast<const uint4*>(epi.mask + n0 + 32 * u);
              const uint4 k0 = __ldg(mk), k1 = __ldg(mk + 1);
              const uint32_t kw[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
              float bm = -INFINITY;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (!((kw[j >> 2] >> (8 * (j & 3))) & 0xFFu)) v[j] = -INFINITY;
                bm = fmaxf(bm, v[j]);
              }
              if (in_m) epi.aux[(long long)m * epi.ld_aux + ((n0 + 32 * u) >> 5)] = bm;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __half2 h0 = __floats2half2_rn(v[8 * j + 0], v[8 * j + 1]), h1 = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
              __half2 h2 = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]), h3 = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
              const uint32_t u0 = *reinterpret_cast<uint32_t*>(&h0), u1 = *reinterpret_cast<uint32_t*>(&h1);
              const uint32_t u2 = *reinterpret_cast<uint32_t*>(&h2), u3 = *reinterpret_cast<uint32_t*>(&h3);
              if (in_m)      // (h & 0x7fff) + 0x0400 carries into bit 15 exactly when the exponent field is 0x1f
                range_chk |= (((u0 & 0x7FFF7FFFu) + 0x04000400u) | ((u1 & 0x7FFF7FFFu) + 0x04000400u)) |
                             (((u2 & 0x7FFF7FFFu) + 0x04000400u) | ((u3 & 0x7FFF7FFFu) + 0x04000400u));
              sts128(row_base + (((u * 4 + j) ^ sw) << 4), u0, u1, u2, u3);
            }
          } else if (MODE == kOutResidual) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(row_base + ((j ^ sw) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                     __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
          } else if (MODE == kOutSplitAct) {
            if (in_m) {
              // wide-range path: v = hi + lo with hi = bf16(v), lo = bf16(v - hi); row layout [hi (N columns) | lo (N columns)]
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float a = act_fn<ACT>(v[2 * j]), b = act_fn<ACT>(v[2 * j + 1]);
                const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
                const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah)), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
                hi[j] = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
                lo[j] = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
              }
              uint4* dh = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(epi.out) + out_row * epi.ld_out + n0 + 32 * u);
              uint4* dl = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(epi.out) + out_row * epi.ld_out + N + n0 + 32 * u);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                dh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                dl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
              }
            }
          } else if (in_m) {
            float* dstf = static_cast<float*>(epi.out) + out_row * epi.ld_out + n0 + 32 * u;
            if (MODE == kOutPatchEmbed) {
              const float4* t4 = reinterpret_cast<const float4*>(table_row + n0 + 32 * u);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 t = __ldg(t4 + j);
                v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
              }
            }
            float4* dst = reinterpret_cast<float4*>(dstf);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        if (kViaTma) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (MODE == kOutResidual) tma_reduce_add_2d(&map_out, stg, n0, m0);
            else tma_store_2d(&map_out, stg, n0, m0);
            bulk_commit();
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (kHalfOut && epi.range_flag && (range_chk & 0x80008000u)) atomicOr(epi.range_flag, 1);
    if (kViaTma && lane == 0) bulk_wait_all();           // outstanding bulk ops read shared memory: finish before exit
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();        // no CTA leaves while its peer may still read its shared memory or arrive on its barriers
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    if (CL == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_operand_map(CUtensorMap* map, const void* ptr, long long rows, long long K, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled entry point not found"); return BSCNAV_ECUDA; }
  if (!ptr || rows <= 0 || K <= 0 || K % BK != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15)) return BSCNAV_EINVAL;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled failed: %d", (int)r); return BSCNAV_ECUDA; }
  return BSCNAV_OK;
}

// packed qkv tensor [batch, T, width] fp16 as a 3-D map {column, token, frame} with box {64, box_rows, 1}: rows past a frame's
// last token are out of bounds in the token dimension and arrive zero-filled (attention_ws.cu relies on that)
int make_qkv_map3d(CUtensorMap* map, const void* ptr, long long batch, long long T, long long width, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled entry point not found"); return BSCNAV_ECUDA; }
  if (!ptr || batch <= 0 || T <= 0 || width <= 0 || width % BK != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15)) return BSCNAV_EINVAL;
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)T, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)width * 2, (cuuint64_t)T * width * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled(3d) failed: %d", (int)r); return BSCNAV_ECUDA; }
  return BSCNAV_OK;
}

// attention output [batch, T, width] fp16 as a 3-D store map {column, token, frame} with box {64, 32, 1}: rows past a frame's last
// token are clipped by the hardware (attention_ws2.cu stores whole 32-row boxes)
int make_out_map3d(CUtensorMap* map, const void* ptr, long long batch, long long T, long long width) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled entry point not found"); return BSCNAV_ECUDA; }
  if (!ptr || batch <= 0 || T <= 0 || width <= 0 || width % BK != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15)) return BSCNAV_EINVAL;
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)T, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)width * 2, (cuuint64_t)T * width * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled(out 3d) failed: %d", (int)r); return BSCNAV_ECUDA; }
  return BSCNAV_OK;
}

// output tensor [rows, N] -> map with a 32-row x 128-byte box (64 fp16 or 32 fp32 columns), 128-byte swizzle
int make_store_map(CUtensorMap* map, const void* ptr, long long rows, long long N, bool is_float) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled entry point not found"); return BSCNAV_ECUDA; }
  const int esz = is_float ? 4 : 2;
  if (!ptr || rows <= 0 || N <= 0 || (N * esz) % 128 != 0 || (reinterpret_cast<uintptr_t>(ptr) & 127)) return BSCNAV_EINVAL;
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)N * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, is_float ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled(store) failed: %d", (int)r); return BSCNAV_ECUDA; }
  return BSCNAV_OK;
}

int gemm_block_n(int N) { return N % 256 == 0 ? 256 : 128; }

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Few output tiles and a long K (one frame per call: 24 tiles for fc2 of ViT-L, each CTA streaming 1 MB of weights through 64
// k-blocks alone): slice K over otherwise idle SMs.  fp32-output modes only, K >= 2048, at least eight k-blocks per slice (the
// meeting of the slices costs ~3 us: slicing the K = 1024 GEMMs was measured slower), never more items than SMs (one wave).
static int kslices_for(int tiles, int k_blocks, int bn, const GemmEpilogue& epi) {
  static int use_slices = -1;
  if (use_slices < 0) { const char* e = getenv("BSCNAV_KSLICE"); use_slices = e ? atoi(e) : 1; }
  if (!use_slices || !epi.kws || !epi.kcnt || tiles > epi.kcnt_len || k_blocks < 32) return 1;
  if ((epi.mode != kOutResidual && epi.mode != kOutFloat) || bn != 128) return 1;
  int s = sm_count() / tiles;
  if (s > k_blocks / 8) s = k_blocks / 8;
  if (s > 8) s = 8;
  const long long slots = epi.kws_floats / ((long long)BM * bn);
  if ((long long)s * tiles > slots) s = (int)(slots / tiles);
  return s >= 2 ? s : 1;
}

static int max_cluster_pairs(const void* fn, int smem) {
  // how many 2-CTA clusters of this kernel the device runs at once (GPCs with an odd SM count leave one SM unpaired)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 74); cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = sm_count() / 2; }
  return n < sm_count() / 2 ? n : sm_count() / 2;
}

template <int BN, int MODE, int ACT = 0>
static int launch_bn_mode(const CUtensorMap* map_a, const CUtensorMap* map_w, const CUtensorMap* map_w_half, const CUtensorMap* map_out,
                          int M, int N, int K, const GemmEpilogue& epi, cudaStream_t stream, int split) {
  static bool attr_set = false;
  static int pairs = 0;
  if (!attr_set) {
    BSC_CUDA(cudaFuncSetAttribute((k_gemm_f16<BN, MODE, ACT, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, 1>::kSmem));
    BSC_CUDA(cudaFuncSetAttribute((k_gemm_f16<BN, MODE, ACT, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, 2>::kSmem));
    pairs = max_cluster_pairs(reinterpret_cast<const void*>(k_gemm_f16<BN, MODE, ACT, 2>), Cfg<BN, 2>::kSmem);
    attr_set = true;
  }
  const int m_tiles = (M + BM - 1) / BM, n_tiles = N / BN;
  if (map_w_half && m_tiles >= 2 && pairs > 0) {
    // CTA-pair variant (cta_group::2): pairs of row blocks; persistent over `pairs` clusters
    const int groups = ((m_tiles + 1) / 2) * n_tiles;
    const int clusters = groups < pairs ? groups : pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = Cfg<BN, 2>::kSmem; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 2;
    BSC_CUDA(cudaLaunchKernelEx(&cfg, k_gemm_f16<BN, MODE, ACT, 2>, *map_a, *map_w_half, *map_out, M, N, K, epi, split));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return BSCNAV_OK;
  }
  const int tiles = m_tiles * n_tiles;
  GemmEpilogue ep = epi;
  ep.kslices = kslices_for(tiles, (split ? 3 : 1) * (K / BK), BN, epi);
  const int items = tiles * ep.kslices;
  const int grid = items < sm_count() ? items : sm_count();
  constexpr int smem1 = Cfg<BN, 1>::kSmem;
  BSC_LAUNCH_PDL((k_gemm_f16<BN, MODE, ACT, 1>), grid, kGemmThreads, smem1, stream, *map_a, *map_w, *map_out, M, N, K, ep, split);
  return BSCNAV_OK;
}

template <int BN>
static int launch_bn(const CUtensorMap* map_a, const CUtensorMap* map_w, const CUtensorMap* map_w_half, const CUtensorMap* map_out,
                     int M, int N, int K, const GemmEpilogue& epi, cudaStream_t stream, int split) {
  switch (epi.mode) {
    case kOutHalf: return launch_bn_mode<BN, kOutHalf>(map_a, map_w, map_w_half, map_out, M, N, K, epi, stream, split);
    case kOutHalfAct:
      if (epi.act == 0) return launch_bn_mode<BN, kOutHalfAct, 0>(map_a, map_w, map_w_half, map_out, M, N, K, epi, stream, split);
      return launch_bn_mode<BN, kOutHalfAct, 1>(map_a, map_w, map_w_half, map_out, M, N, K, epi, stream, split);
    case kOutSplitAct:
      if (epi.act == 0) return launch_bn_mode<BN, kOutSplitAct, 0>(map_a, map_w, map_w_half, map_out, M, N, K, epi, stream, split);
      return launch_bn_mode<BN, kOutSplitAct, 1>(map_a, map_w, map_w_half, map_out, M, N, K, epi, stream, split);
    case kOutResidual: return launch_bn_mode<BN, kOutResidual>(map_a, map_w, map_w_half, map_out, M, N, K, epi, stream, split);
    case kOutPatchEmbed: return launch_bn_mode<BN, kOutPatchEmbed>(map_a, map_w, map_w_half, map_out, M, N, K, epi, stream, split);
    case kOutFloat: return launch_bn_mode<BN, kOutFloat>(map_a, map_w, map_w_half, map_out, M, N, K, epi, stream, split);
    case kOutSims: return launch_bn_mode<BN, kOutSims>(map_a, map_w, map_w_half, map_out, M, N, K, epi, stream, split);
    default: return BSCNAV_EINVAL;
  }
}

int gemm_f16_launch(const CUtensorMap* map_a, const CUtensorMap* map_w, const CUtensorMap* map_out, int M, int N, int K,
                    const GemmEpilogue& epi, cudaStream_t stream, const CUtensorMap* map_w_half, int split) {
  if (M <= 0 || N <= 0 || K <= 0 || N % 128 != 0 || K % BK != 0 || !epi.out) return BSCNAV_EINVAL;
  const bool via_tma = epi.mode == kOutHalf || epi.mode == kOutHalfAct || epi.mode == kOutResidual || epi.mode == kOutSims;
  if (via_tma && !map_out) return BSCNAV_EINVAL;
  if (!map_out) map_out = map_a;       // unused by the direct-store epilogues
  static int use_cluster = -1;
  if (use_cluster < 0) { const char* e = getenv("BSCNAV_GEMM_CLUSTER"); use_cluster = e ? atoi(e) : 1; }   // CTA pairs by default
  // few output tiles (small M: one goal image, one text prompt, a handful of queries): 128-wide tiles through the 128-row-box
  // weight map double the number of CTAs that stream the layer's weights -- these calls are bound by that, not by the MMAs
  if (gemm_block_n(N) == 256 && map_w_half && ((M + BM - 1) / BM) * (N / 256) * 2 <= sm_count())
    return launch_bn<128>(map_a, map_w_half, nullptr, map_out, M, N, K, epi, stream, split);
  if (!use_cluster) map_w_half = nullptr;
  if (gemm_block_n(N) == 256) return launch_bn<256>(map_a, map_w, map_w_half, map_out, M, N, K, epi, stream, split);
  return launch_bn<128>(map_a, map_w, nullptr, map_out, M, N, K, epi, stream, split);
}

}  // namespace bscnav

using namespace bscnav;

extern "C" int bscnav_gemm_f16(const void* a, const void* w, const float* bias, void* out, int32_t m, int32_t n, int32_t k,
                               int32_t out_mode, int32_t act, void* stream) {
  if (!a || !w || !out || m <= 0 || n <= 0 || k <= 0 || n % 128 != 0 || k % 64 != 0) return BSCNAV_EINVAL;
  if (out_mode != kOutHalf && out_mode != kOutHalfAct && out_mode != kOutResidual && out_mode != kOutFloat) return BSCNAV_EINVAL;
  CUtensorMap ma, mw, mo;
  if (int e = make_operand_map(&ma, a, m, k, BM)) return e;
  if (int e = make_operand_map(&mw, w, n, k, gemm_block_n(n))) return e;
  CUtensorMap mwh;
  const bool half_ok = gemm_block_n(n) == 256;
  if (half_ok) if (int e = make_operand_map(&mwh, w, n, k, 128)) return e;
  const bool via_tma = out_mode != kOutFloat;
  if (via_tma)
    if (int e = make_store_map(&mo, out, m, n, out_mode == kOutResidual)) return e;
  GemmEpilogue epi{};
  epi.out = out; epi.bias = bias; epi.table = nullptr; epi.ld_out = n; epi.mode = out_mode; epi.act = act; epi.range_flag = nullptr;
  return gemm_f16_launch(&ma, &mw, via_tma ? &mo : nullptr, m, n, k, epi, (cudaStream_t)stream, half_ok ? &mwh : nullptr);
}

// The same GEMM with K slices allowed (tests of the small-M path): scratch = [1 KB of zeroed tile counters | partial tiles];
// out_slices (host, may be null) receives the slice count the launcher chose for this shape.
extern "C" int bscnav_gemm_f16_sliced(const void* a, const void* w, const float* bias, void* out, int32_t m, int32_t n, int32_t k,
                                      int32_t out_mode, int32_t act, void* scratch, int64_t scratch_bytes, int32_t* out_slices,
                                      void* stream) {
  if (!a || !w || !out || m <= 0 || n <= 0 || k <= 0 || n % 128 != 0 || k % 64 != 0) return BSCNAV_EINVAL;
  if (out_mode != kOutHalf && out_mode != kOutHalfAct && out_mode != kOutResidual && out_mode != kOutFloat) return BSCNAV_EINVAL;
  if (!scratch || scratch_bytes < 1024 + 2LL * BM * 128 * 4 || (reinterpret_cast<uintptr_t>(scratch) & 15)) return BSCNAV_EINVAL;
  CUtensorMap ma, mw, mo;
  if (int e = make_operand_map(&ma, a, m, k, BM)) return e;
  if (int e = make_operand_map(&mw, w, n, k, 128)) return e;       // 128-wide tiles, one CTA per tile: the path that slices
  const bool via_tma = out_mode != kOutFloat;
  if (via_tma)
    if (int e = make_store_map(&mo, out, m, n, out_mode == kOutResidual)) return e;
  GemmEpilogue epi{};
  epi.out = out; epi.bias = bias; epi.ld_out = n; epi.mode = out_mode; epi.act = act;
  epi.kcnt = static_cast<int*>(scratch); epi.kcnt_len = 256;
  epi.kws = reinterpret_cast<float*>(static_cast<char*>(scratch) + 1024); epi.kws_floats = (scratch_bytes - 1024) / 4;
  if (out_slices) *out_slices = kslices_for(((m + BM - 1) / BM) * (n / 128), k / BK, 128, epi);
  return launch_bn<128>(&ma, &mw, nullptr, via_tma ? &mo : &ma, m, n, k, epi, (cudaStream_t)stream, 0);
}

// Stand-alone wide-range GEMM (tests): a [M, 2K] and w [N, 2K] bf16 (hi | lo) rows (vit.py:split_bf16), fp32 out = a w^T + bias.
extern "C" int bscnav_gemm_bf16x3(const void* a, const void* w, const float* bias, float* out, int32_t m, int32_t n, int32_t k,
                                  void* stream) {
  if (!a || !w || !out || m <= 0 || n <= 0 || k <= 0 || n % 128 != 0 || k % 64 != 0) return BSCNAV_EINVAL;
  CUtensorMap ma, mw, mwh;
  if (int e = make_operand_map(&ma, a, m, 2LL * k, BM)) return e;
  if (int e = make_operand_map(&mw, w, n, 2LL * k, gemm_block_n(n))) return e;
  const bool half_ok = gemm_block_n(n) == 256;
  if (half_ok) if (int e = make_operand_map(&mwh, w, n, 2LL * k, 128)) return e;
  GemmEpilogue epi{};
  epi.out = out; epi.bias = bias; epi.ld_out = n; epi.mode = kOutFloat;
  return gemm_f16_launch(&ma, &mw, nullptr, m, n, k, epi, (cudaStream_t)stream, half_ok ? &mwh : nullptr, 1);
}
