// Batched goal queries (BASELINE configs[2] and [3]: thousands of text-goal queries, batch-256 image queries):
// cosine similarity of Q queries against every stored token, per-cell max, top-K  (memory_2.py:642-671), for many
// queries per pass over the store.
//
// One query is HBM-bound (k_query_slots streams the fp32 rows once).  Eight or more queries per pass are not: on the
// CUDA cores the scan turns shared-memory bound (measured 0.65 ms for 8 queries, 5.2 ms for 64).  So the batch path puts
// the contraction on the tensor cores and keeps the reference's exact fp32 arithmetic only where it decides the result:
//
//   1. shadow      unit-normalised fp16 copy of the token rows (built once per store state by k_build_shadow)
//   2. GEMM        approx[q, row] = q^[q, :] . f^[row, :]   (k_gemm_f16 on tcgen05, fp32 accumulate) -- within
//                  eps = 1.1e-3 of the exact cosine: both operands are rounded once to fp16 (relative 2^-11 each),
//                  |sum(f^ q^) - cos| <= (2^-10 + 2^-22) sum|f^_i q^_i| <= 9.8e-4 by Cauchy-Schwarz, plus fp32
//                  accumulation and subnormal terms < 1e-4
//   (C == 1, the 512 x 512 BEV grid: steps 3-4 run on the GEMM epilogue's per-32-cell block maxima and fp16 similarities instead,
//    see k_blk_hist / k_blk_select below: the similarity matrix is written once in fp16 and read only where a block opens)
//   3. histogram   per query, 2048 bins of width 2^-10 over [-1, 1] of the per-cell max of approx (k_batch_hist)
//   4. select      tb = bin holding the K-th largest approx value; every cell whose bin is >= tb - 3 becomes a
//                  candidate (k_batch_select).  At least K cells have approx >= edge(tb), hence exact >= edge(tb) - eps;
//                  a rejected cell has approx < edge(tb) - 3 * 2^-10, hence exact < edge(tb) - 2.93e-3 + eps, which is
//                  below edge(tb) - eps because eps < 1.46e-3: no rejected cell can belong to the exact top-K
//   5. re-score    exact fp32 cosine of each candidate with the very FMA order of k_query_slots, composite key,
//                  bitonic top-K, decode (shared stages)  ->  results are bit-identical to the single-query path
//
// A query whose candidate count exceeds kCandCap (more than ~4000 cells within 4e-3 of its K-th similarity) is flagged
// with out_count = -1; the host re-issues it through bscnav_query_topk.
#include "gemm_tcgen05.cuh"
#include "memory_kernels.cuh"

namespace bscnav {

static constexpr int kBins = 2048;
static constexpr int kCandCap = kTopkChunk;
static constexpr int kMarginBins = 3;
static constexpr int kSelSlots = 8192;      // cells per block in the histogram / select passes (256 threads x 8 x 4)
static constexpr int kPassQueries = 1024;    // queries per GEMM pass

__device__ __forceinline__ int sim_bin(float s) {
  const int b = (int)((s + 1.0f) * 1024.0f);          // NaN converts to 0
  return b < 0 ? 0 : (b > kBins - 1 ? kBins - 1 : b);
}

// unit-normalised fp16 copy of rows [0, slots_in_use * C); one warp per row
__global__ void __launch_bounds__(kMemThreads)
k_build_shadow(const float* __restrict__ feat, const int32_t* __restrict__ counters, int C, int D, __half* __restrict__ shadow) {
  const int lane = threadIdx.x & 31;
  const long long n_rows = (long long)counters[0] * C;
  const long long warps = (long long)gridDim.x * (kMemThreads / 32);
  for (long long row = (long long)blockIdx.x * (kMemThreads / 32) + (threadIdx.x >> 5); row < n_rows; row += warps) {
    const float4* s = reinterpret_cast<const float4*>(feat + row * D);
    float acc = 0.f;
    for (int c = lane; c < (D >> 2); c += 32) {
      const float4 v = s[c];
      acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    acc = warp_sum(acc);
    const float inv = acc > 0.f ? 1.0f / sqrtf(acc) : 0.f;
    uint2* d = reinterpret_cast<uint2*>(shadow + row * D);
    for (int c = lane; c < (D >> 2); c += 32) {
      const float4 v = s[c];
      const __half2 a = __floats2half2_rn(v.x * inv, v.y * inv), b = __floats2half2_rn(v.z * inv, v.w * inv);
      d[c] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
  }
}

// queries: squared norm with the association k_query_slots uses (so the exact re-scoring matches it bit for bit) and
// the unit-normalised fp16 operand of the GEMM; one warp per query
__global__ void __launch_bounds__(kMemThreads)
k_prep_queries(const float* __restrict__ q, int n_query, int D, __half* __restrict__ qh, float* __restrict__ qn2,
               uint32_t* __restrict__ zero_words, long long n_zero, unsigned long long* __restrict__ empty_keys, long long n_keys) {
  // the pass's histogram / counters / flags start from zero and its candidate keys from KEY_EMPTY: filled here (grid-stride over
  // the whole grid) instead of by four memsets between the kernels of the pass
  for (long long i = (long long)blockIdx.x * kMemThreads + threadIdx.x; i < n_zero; i += (long long)gridDim.x * kMemThreads) zero_words[i] = 0u;
  for (long long i = (long long)blockIdx.x * kMemThreads + threadIdx.x; i < n_keys; i += (long long)gridDim.x * kMemThreads)
    empty_keys[i] = BSCNAV_KEY_EMPTY;
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * (kMemThreads / 32) + (threadIdx.x >> 5);
  if (qi >= n_query) return;
  const float4* s = reinterpret_cast<const float4*>(q + (long long)qi * D);
  float acc = 0.f;
  for (int c = lane; c < (D >> 2); c += 32) {
    const float4 v = s[c];
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) qn2[qi] = acc;
  const float inv = acc > 0.f ? 1.0f / sqrtf(acc) : 0.f;
  uint2* d = reinterpret_cast<uint2*>(qh + (long long)qi * D);
  for (int c = lane; c < (D >> 2); c += 32) {
    const float4 v = s[c];
    const __half2 a = __floats2half2_rn(v.x * inv, v.y * inv), b = __floats2half2_rn(v.z * inv, v.w * inv);
    d[c] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
}

// per token row: does it take part in this batch?  (non-empty, passes the region / floor filter of memory_2.py:624-640;
// voxel keys: the whole cell passes or fails; BEV keys: the token's own height decides, as in k_query_slots)
__global__ void __launch_bounds__(kMemThreads)
k_row_ok(const bscnav_geom_t g, const unsigned long long* __restrict__ row_key, const int32_t* __restrict__ slot_cell,
         const int32_t* __restrict__ tok_h, const int32_t* __restrict__ counters, const bscnav_filter_t flt,
         uint8_t* __restrict__ row_ok) {
  const long long r = (long long)blockIdx.x * kMemThreads + threadIdx.x;
  const int C = g.cache;
  if (r >= (long long)g.max_slots * C) return;
  if (r >= (long long)counters[0] * C) { row_ok[r] = 0; return; }        // rows past the slots in use never take part
  const int slot = (int)(r / C);
  bool ok = row_key[r] != BSCNAV_KEY_EMPTY;
  if (ok) {
    const int cell = slot_cell[slot];
    int row, col, h;
    if (g.levels == 1) { row = cell / g.gs; col = cell - row * g.gs; h = tok_h[r]; }
    else { h = cell % g.levels; const int rc = cell / g.levels; row = rc / g.gs; col = rc - row * g.gs; }
    ok = filter_pass(flt, row, col, h);
  }
  row_ok[r] = ok ? 1 : 0;
}

__device__ __forceinline__ float approx_cell(const float* __restrict__ sims_q, const uint8_t* __restrict__ row_ok, int slot, int C) {
  float best = -INFINITY;
  for (int j = 0; j < C; ++j) {
    const long long r = (long long)slot * C + j;
    if (row_ok[r]) best = fmaxf(best, sims_q[r]);
  }
  return best;
}

// Visit every participating cell of the block's chunk [slot0, slot0 + kSelSlots): f(slot, per-cell max of approx).
// C == 1 (the 512 x 512 BEV grid): each thread owns groups of four consecutive cells and issues all of its 128-bit
// similarity loads and 32-bit flag loads before using any of them (the passes are latency-bound otherwise: measured
// 219 us -> see profiles/); other C: plain loop.
template <typename F>
__device__ __forceinline__ void visit_chunk(const float* __restrict__ sims_q, const uint8_t* __restrict__ row_ok, int slot0, int n_slots,
                                            int C, F f) {
  if (C == 1) {
    constexpr int kIter = kSelSlots / (kMemThreads * 4);
    float4 v[kIter];
    uint32_t ok[kIter];
#pragma unroll
    for (int i = 0; i < kIter; ++i) {
      const int s = slot0 + (i * kMemThreads + threadIdx.x) * 4;
      if (s + 3 < n_slots) {             // slot0 and the row stride are multiples of 4: both loads are aligned
        v[i] = ld_stream_f4(reinterpret_cast<const float4*>(sims_q + s));
        ok[i] = *reinterpret_cast<const uint32_t*>(row_ok + s);
      } else {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        ok[i] = 0;
        for (int e = 0; e < 4; ++e)
          if (s + e < n_slots && row_ok[s + e]) { t[e] = sims_q[s + e]; ok[i] |= 1u << (8 * e); }
        v[i] = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < kIter; ++i) {
      const int s = slot0 + (i * kMemThreads + threadIdx.x) * 4;
      if (ok[i] & 0x000000FFu) f(s, v[i].x);
      if (ok[i] & 0x0000FF00u) f(s + 1, v[i].y);
      if (ok[i] & 0x00FF0000u) f(s + 2, v[i].z);
      if (ok[i] & 0xFF000000u) f(s + 3, v[i].w);
    }
  } else {
    for (int i = threadIdx.x; i < kSelSlots; i += kMemThreads) {
      const int slot = slot0 + i;
      if (slot < n_slots) {
        const float sv = approx_cell(sims_q, row_ok, slot, C);
        if (sv > -INFINITY) f(slot, sv);
      }
    }
  }
}

__global__ void __launch_bounds__(kMemThreads)
k_batch_hist(const float* __restrict__ sims, long long ld, const uint8_t* __restrict__ row_ok, const int32_t* __restrict__ counters,
             int C, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[kBins];
  const int qi = blockIdx.y;
  const int n_slots = counters[0];
  const int slot0 = blockIdx.x * kSelSlots;
  if (slot0 >= n_slots) return;
  for (int b = threadIdx.x; b < kBins; b += kMemThreads) h[b] = 0;
  __syncthreads();
  visit_chunk(sims + (long long)qi * ld, row_ok, slot0, n_slots, C, [&](int, float sv) {
    if (sv == sv) atomicAdd(&h[sim_bin(sv)], 1u);          // NaN similarities take no part (fmaxf drops them in the exact path)
  });
  __syncthreads();
  for (int b = threadIdx.x; b < kBins; b += kMemThreads)
    if (h[b]) atomicAdd(&hist[(long long)qi * kBins + b], h[b]);
}

// Block-wide (kMemThreads threads, all must call): the bin that holds the k-th largest entry counted in hist_q (0 when fewer
// than k entries exist), lowered by `margin` bins; *n_sel = number of entries in bins >= the returned bin.
__device__ __forceinline__ int threshold_bin(const uint32_t* __restrict__ hist_q, int k, int margin, uint32_t* n_sel) {
  static_assert(kBins == 8 * kMemThreads, "one thread owns eight bins");
  __shared__ uint32_t part[kMemThreads], above_of[kMemThreads];
  __shared__ int s_tb;
  __shared__ uint32_t s_ncand;
  const int tid = threadIdx.x, lane = tid & 31;
  uint32_t local[8], sum = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { local[i] = hist_q[tid * 8 + i]; sum += local[i]; }
  part[tid] = sum;
  if (tid == 0) { s_tb = 0; s_ncand = 0; }
  __syncthreads();
  if (tid < 32) {
    uint32_t p8[8], s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { p8[i] = part[8 * lane + i]; s += p8[i]; }
    uint32_t suf = s;                                   // inclusive suffix sum over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_down_sync(0xffffffffu, suf, o);
      if (lane + o < 32) suf += t;
    }
    uint32_t above = suf - s;                           // values in bins above this lane's 64 bins
#pragma unroll
    for (int i = 7; i >= 0; --i) { above_of[8 * lane + i] = above; above += p8[i]; }
  }
  __syncthreads();
  {
    uint32_t above = above_of[tid];
#pragma unroll
    for (int i = 7; i >= 0; --i) {
      const uint32_t c = above + local[i];
      if (above < (uint32_t)k && c >= (uint32_t)k) s_tb = tid * 8 + i;
      above = c;
    }
  }
  __syncthreads();
  const int t = s_tb > margin ? s_tb - margin : 0;
  if (t >= tid * 8 && t < tid * 8 + 8) {
    uint32_t n = above_of[tid];
    for (int i = t - tid * 8; i < 8; ++i) n += local[i];
    s_ncand = n;
  }
  __syncthreads();
  *n_sel = s_ncand;
  return t;
}

__global__ void __launch_bounds__(kMemThreads)
k_batch_select(const float* __restrict__ sims, long long ld, const uint8_t* __restrict__ row_ok, const int32_t* __restrict__ counters,
               int C, const uint32_t* __restrict__ hist, int k, int32_t* __restrict__ cand, int32_t* __restrict__ cand_count,
               int32_t* __restrict__ overflow) {
  const int qi = blockIdx.y, tid = threadIdx.x;
  const int n_slots = counters[0];
  const int slot0 = blockIdx.x * kSelSlots;
  if (slot0 >= n_slots) return;
  uint32_t n_sel;
  const int t = threshold_bin(hist + (long long)qi * kBins, k, kMarginBins, &n_sel);
  if (n_sel > (uint32_t)kCandCap) {
    if (blockIdx.x == 0 && tid == 0) overflow[qi] = (int32_t)n_sel;
    return;
  }
  visit_chunk(sims + (long long)qi * ld, row_ok, slot0, n_slots, C, [&](int slot, float sv) {
    if (sv == sv && sim_bin(sv) >= t) {
      const int pos = atomicAdd(&cand_count[qi], 1);
      if (pos < kCandCap) cand[(long long)qi * kCandCap + pos] = slot;
    }
  });
}

// ---- C == 1 (the 512 x 512 BEV grid): selection over the GEMM epilogue's block maxima.
// The epilogue (kOutSims) leaves sims16[q, cell] (fp16, -inf where the cell takes no part) and bmax[q, cell / 32] (fp32 maximum of
// each 32-cell block).  tb = bin of the K-th largest BLOCK maximum: K distinct blocks each hold a cell >= their maximum, so at
// least K cells have approx >= edge(tb) and the K-th exact similarity is >= edge(tb) - eps.  A cell is dropped when its block's
// maximum or its own fp16 value lies below bin tb - 3; then approx < edge(tb) - 3 * 2^-10 + 2^-11 (fp16 rounding of the stored
// value), exact < edge(tb) - 2.93e-3 + 4.9e-4 + eps < edge(tb) - eps for eps = 1.1e-3: it cannot belong to the exact top-K.
// The full similarity matrix is written once (fp16) and read only where a block opens; the histogram reads 1/32 of it.
static constexpr int kBlkChunk = 2048;       // block maxima per thread block in the two passes below

__global__ void __launch_bounds__(kMemThreads)
k_blk_hist(const float* __restrict__ bmax, long long ld, int n_blk, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[kBins];
  pdl_wait();
  pdl_launch_dependents();
  const int qi = blockIdx.y;
  const int b0 = blockIdx.x * kBlkChunk;
  for (int b = threadIdx.x; b < kBins; b += kMemThreads) h[b] = 0;
  __syncthreads();
  const float* bq = bmax + (long long)qi * ld;
  float v[kBlkChunk / kMemThreads];
#pragma unroll
  for (int i = 0; i < kBlkChunk / kMemThreads; ++i) {
    const int b = b0 + i * kMemThreads + threadIdx.x;
    v[i] = b < n_blk ? bq[b] : -INFINITY;
  }
#pragma unroll
  for (int i = 0; i < kBlkChunk / kMemThreads; ++i)
    if (v[i] > -INFINITY) atomicAdd(&h[sim_bin(v[i])], 1u);      // NaN fails the comparison: takes no part
  __syncthreads();
  for (int b = threadIdx.x; b < kBins; b += kMemThreads)
    if (h[b]) atomicAdd(&hist[(long long)qi * kBins + b], h[b]);
}

// one block per query: the threshold bin, once (every select block used to re-derive it from the 8 KB histogram)
__global__ void __launch_bounds__(kMemThreads)
k_blk_threshold(const uint32_t* __restrict__ hist, int k, int32_t* __restrict__ thr) {
  pdl_wait();
  pdl_launch_dependents();
  uint32_t n_sel;
  const int t = threshold_bin(hist + (long long)blockIdx.x * kBins, k, kMarginBins, &n_sel);
  if (threadIdx.x == 0) thr[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kMemThreads)
k_blk_select(const __half* __restrict__ sims16, long long ld, const float* __restrict__ bmax, long long ld_b, int n_blk,
             const int32_t* __restrict__ thr, int32_t* __restrict__ cand, int32_t* __restrict__ cand_count,
             int32_t* __restrict__ overflow) {
  pdl_wait();
  pdl_launch_dependents();
  const int qi = blockIdx.y;
  const int b0 = blockIdx.x * kBlkChunk;
  const int t = thr[qi];
  const float* bq = bmax + (long long)qi * ld_b;
  const __half* sq = sims16 + (long long)qi * ld;
  for (int i = threadIdx.x; i < kBlkChunk; i += kMemThreads) {
    const int b = b0 + i;
    if (b >= n_blk) break;
    const float bm = bq[b];
    if (!(bm > -INFINITY) || sim_bin(bm) < t) continue;
    const uint4* p = reinterpret_cast<const uint4*>(sq + (long long)b * 32);        // the block's 32 fp16 similarities
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 u = p[c];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float sv = __half2float(__ushort_as_half((unsigned short)(w[e >> 1] >> (16 * (e & 1)))));
        if (sv > -INFINITY && sim_bin(sv) >= t) {
          const int pos = atomicAdd(&cand_count[qi], 1);
          if (pos < kCandCap) cand[(long long)qi * kCandCap + pos] = b * 32 + c * 8 + e;
          else overflow[qi] = pos + 1;
        }
      }
    }
  }
}

// ---- the same two stages over exact composite keys (single-query path, bscnav_query_topk): the bins are a monotone function of
// the similarity, so every key in a bin at or above the k-th key's bin is kept and the sort of those few keys is exact
__device__ __forceinline__ int key_bin(unsigned long long key) { return sim_bin(from_orderable(~(uint32_t)(key >> 32))); }

__global__ void __launch_bounds__(kMemThreads)
k_keys_hist(const unsigned long long* __restrict__ keys, long long stride, const int32_t* __restrict__ counters,
            uint32_t* __restrict__ hist, int chunk) {
  __shared__ uint32_t h[kBins];
  pdl_wait();
  pdl_launch_dependents();
  const int qi = blockIdx.y;
  const int n_slots = counters[0];
  const int slot0 = blockIdx.x * chunk;
  if (slot0 >= n_slots) return;
  for (int b = threadIdx.x; b < kBins; b += kMemThreads) h[b] = 0;
  __syncthreads();
  const unsigned long long* kq = keys + (long long)qi * stride;
  for (int i = threadIdx.x; i < chunk; i += kMemThreads) {
    const int slot = slot0 + i;
    if (slot < n_slots) {
      const unsigned long long key = kq[slot];
      if (key != BSCNAV_KEY_EMPTY) atomicAdd(&h[key_bin(key)], 1u);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kBins; b += kMemThreads)
    if (h[b]) atomicAdd(&hist[(long long)qi * kBins + b], h[b]);
}

__global__ void __launch_bounds__(kMemThreads)
k_keys_select(const unsigned long long* __restrict__ keys, long long stride, const int32_t* __restrict__ counters,
              const uint32_t* __restrict__ hist, int k, unsigned long long* __restrict__ cand_keys, int32_t* __restrict__ cand_count,
              int32_t* __restrict__ overflow, int chunk) {
  pdl_wait();
  pdl_launch_dependents();
  const int qi = blockIdx.y, tid = threadIdx.x;
  const int n_slots = counters[0];
  const int slot0 = blockIdx.x * chunk;
  if (slot0 >= n_slots) return;
  uint32_t n_sel;
  const int t = threshold_bin(hist + (long long)qi * kBins, k, 0, &n_sel);
  if (n_sel > (uint32_t)kCandCap) {                      // thousands of ties in the k-th key's bin: the chunked sorter takes over
    if (blockIdx.x == 0 && tid == 0) overflow[qi] = 1;
    return;
  }
  const unsigned long long* kq = keys + (long long)qi * stride;
  for (int i = tid; i < chunk; i += kMemThreads) {
    const int slot = slot0 + i;
    if (slot < n_slots) {
      const unsigned long long key = kq[slot];
      if (key != BSCNAV_KEY_EMPTY && key_bin(key) >= t) {
        const int pos = atomicAdd(&cand_count[qi], 1);
        if (pos < kCandCap) cand_keys[(long long)qi * kCandCap + pos] = key;
      }
    }
  }
}

// exact fp32 cosine of every candidate cell, arithmetic identical to k_query_slots; one warp per candidate
__global__ void __launch_bounds__(kMemThreads)
k_batch_rescore(const bscnav_geom_t g, const float* __restrict__ feat, const float* __restrict__ norm2,
                const uint8_t* __restrict__ row_ok, const int32_t* __restrict__ slot_cell, const float* __restrict__ q,
                const float* __restrict__ qn2, const int32_t* __restrict__ cand, const int32_t* __restrict__ cand_count,
                const int32_t* __restrict__ overflow, unsigned long long* __restrict__ keys) {
  pdl_wait();
  pdl_launch_dependents();
  const int qi = blockIdx.y, lane = threadIdx.x & 31;
  const int D = g.dim, C = g.cache;
  const int n = overflow[qi] ? 0 : min(cand_count[qi], kCandCap);
  const float4* qv = reinterpret_cast<const float4*>(q + (long long)qi * D);
  const float w1 = qn2[qi];
  const int warps = gridDim.x * (kMemThreads / 32);
  for (int idx = blockIdx.x * (kMemThreads / 32) + (threadIdx.x >> 5); idx < n; idx += warps) {
    const int slot = cand[(long long)qi * kCandCap + idx];
    float best = -INFINITY;
    for (int j = 0; j < C; ++j) {
      const long long r = (long long)slot * C + j;
      if (!row_ok[r]) continue;
      const float4* fr = reinterpret_cast<const float4*>(feat + r * D);
      float acc = 0.f;
      for (int c = lane; c < (D >> 2); c += 32) {
        const float4 v = ld_stream_f4(fr + c), qq = qv[c];
        acc = fmaf(v.x, qq.x, acc); acc = fmaf(v.y, qq.y, acc); acc = fmaf(v.z, qq.z, acc); acc = fmaf(v.w, qq.w, acc);
      }
      acc = warp_sum(acc);
      const float s = acc / sqrtf(fmaxf(w1 * norm2[r], 1e-16f));      // F.cosine_similarity, memory_2.py:656
      best = fmaxf(best, s);
    }
    if (lane == 0 && best > -INFINITY)
      keys[(long long)qi * kCandCap + idx] = ((unsigned long long)(~orderable(best)) << 32) | (uint32_t)slot_cell[slot];
  }
}

// ascending sort of one query's candidate keys (at most kCandCap, usually a few hundred): bitonic network over the next
// power of two above the candidate count instead of the full 4096 the shared chunk sorter always sorts
__global__ void __launch_bounds__(1024)
k_sort_candidates(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ cand_count,
                  const int32_t* __restrict__ overflow, int k, unsigned long long* __restrict__ sorted, int32_t* __restrict__ zero_count) {
  __shared__ unsigned long long s[kCandCap];
  pdl_wait();
  pdl_launch_dependents();
  const int qi = blockIdx.x;
  if (zero_count && threadIdx.x == 0) zero_count[qi] = 0;        // the decode stage counts its results with atomics
  const int n = overflow[qi] ? 0 : min(cand_count[qi], kCandCap);
  int S = 32;
  while (S < n) S <<= 1;
  for (int i = threadIdx.x; i < S; i += 1024) s[i] = i < n ? keys[(long long)qi * kCandCap + i] : BSCNAV_KEY_EMPTY;
  __syncthreads();
  for (int size = 2; size <= S; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < S / 2; t += 1024) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long a = s[lo], b = s[hi];
        if ((a > b) == up) { s[lo] = b; s[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += 1024) sorted[(long long)qi * k + i] = i < S ? s[i] : BSCNAV_KEY_EMPTY;
}

__global__ void k_mark_overflow(const int32_t* __restrict__ overflow, int n_query, int32_t* __restrict__ out_count) {
  pdl_wait();
  pdl_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_query && overflow[i]) out_count[i] = -1;
}

struct BatchScratch {
  __half* qh; float* qn2; uint8_t* row_ok; float* sims; float* bmax; uint32_t* hist; int32_t* cand_count; int32_t* overflow; int32_t* thr;
  int32_t* cand; unsigned long long* keys; unsigned long long* sorted;
};
static long long batch_rows(const bscnav_geom_t* g, long long n_rows_hint) {
  const long long cap = align_up((long long)g->max_slots * g->cache, 256);
  long long n = n_rows_hint > 0 ? align_up(n_rows_hint, 256) : cap;
  return n < cap ? n : cap;
}
static int64_t batch_layout(const bscnav_geom_t* g, int n_query, int k, long long n_rows_hint, BatchScratch* s, char* base) {
  const long long P = n_query < kPassQueries ? n_query : kPassQueries, N = batch_rows(g, n_rows_hint);
  long long off = 0;
  auto take = [&](long long bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes, 1024); return p; };
  char* p;
  p = take(align_up(P, 128) * g->dim * 2); if (s) s->qh = reinterpret_cast<__half*>(p);
  p = take(P * 4); if (s) s->qn2 = reinterpret_cast<float*>(p);
  p = take((long long)g->max_slots * g->cache); if (s) s->row_ok = reinterpret_cast<uint8_t*>(p);
  p = take(P * N * (g->cache == 1 ? 2 : 4)); if (s) s->sims = reinterpret_cast<float*>(p);      // C == 1: fp16 similarities
  p = take(g->cache == 1 ? align_up(P, 128) * (N / 32) * 4 : 0); if (s) s->bmax = reinterpret_cast<float*>(p);
  p = take(P * kBins * 4); if (s) s->hist = reinterpret_cast<uint32_t*>(p);
  p = take(P * 4); if (s) s->cand_count = reinterpret_cast<int32_t*>(p);
  p = take(P * 4); if (s) s->overflow = reinterpret_cast<int32_t*>(p);
  p = take(P * 4); if (s) s->thr = reinterpret_cast<int32_t*>(p);
  p = take(P * kCandCap * 4); if (s) s->cand = reinterpret_cast<int32_t*>(p);
  p = take(P * kCandCap * 8); if (s) s->keys = reinterpret_cast<unsigned long long*>(p);
  p = take(P * (long long)k * 8); if (s) s->sorted = reinterpret_cast<unsigned long long*>(p);
  return off;
}

int64_t exact_select_zero_bytes(int n_query) {
  return align_up((long long)n_query * kBins * 4, 256) + 2 * align_up((long long)n_query * 4, 256);
}

int64_t exact_select_scratch_bytes(int n_query) {
  return align_up((long long)n_query * kBins * 4, 256) + 2 * align_up((long long)n_query * 4, 256) +
         align_up((long long)n_query * kCandCap * 8, 256);
}

// keys [n_query][stride] (composite keys, BSCNAV_KEY_EMPTY = not a candidate) -> sorted [n_query][k] ascending; queries whose
// k-th bin holds more than kCandCap keys are flagged in *overflow_out (device int32[n_query]) and left to the caller
int launch_exact_select(const unsigned long long* keys, long long stride, const int32_t* counters, int max_slots, int n_query, int k,
                        void* scratch, unsigned long long* sorted, int32_t** overflow_out, cudaStream_t stream, bool zeroed,
                        int32_t* zero_count) {
  char* base = static_cast<char*>(scratch);
  uint32_t* hist = reinterpret_cast<uint32_t*>(base);
  base += align_up((long long)n_query * kBins * 4, 256);
  int32_t* cand_count = reinterpret_cast<int32_t*>(base);
  base += align_up((long long)n_query * 4, 256);
  int32_t* overflow = reinterpret_cast<int32_t*>(base);
  base += align_up((long long)n_query * 4, 256);
  unsigned long long* cand_keys = reinterpret_cast<unsigned long long*>(base);
  if (!zeroed) {
    BSC_CUDA(cudaMemsetAsync(hist, 0, (size_t)n_query * kBins * 4, stream));
    BSC_CUDA(cudaMemsetAsync(cand_count, 0, (size_t)n_query * 4, stream));
    BSC_CUDA(cudaMemsetAsync(overflow, 0, (size_t)n_query * 4, stream));
  }
  // at most eight queries per call: small chunks, so that a few hundred blocks share the 2 MB of keys per query (with the batch
  // path's 8192-cell chunks the two stages were latency-bound at ~21 us each)
  const int chunk = 1024;
  const unsigned blocks = (unsigned)((max_slots + chunk - 1) / chunk);
  BSC_LAUNCH_PDL(k_keys_hist, dim3(blocks, (unsigned)n_query), kMemThreads, 0, stream, keys, stride, counters, hist, chunk);
  BSC_LAUNCH_PDL(k_keys_select, dim3(blocks, (unsigned)n_query), kMemThreads, 0, stream, keys, stride, counters, (const uint32_t*)hist, k,
                 cand_keys, cand_count, overflow, chunk);
  BSC_LAUNCH_PDL(k_sort_candidates, (unsigned)n_query, 1024, 0, stream, (const unsigned long long*)cand_keys, (const int32_t*)cand_count,
                 (const int32_t*)overflow, k, sorted, zero_count);
  *overflow_out = overflow;
  return BSCNAV_OK;
}

}  // namespace bscnav

using namespace bscnav;

extern "C" {

int64_t bscnav_shadow_bytes(const bscnav_geom_t* g) {
  if (mem_check_geom(g)) return BSCNAV_EINVAL;
  return align_up((long long)g->max_slots * g->cache, 256) * g->dim * 2;
}

int bscnav_store_build_shadow(const bscnav_geom_t* g, const bscnav_store_t* st, void* shadow, void* stream_) {
  if (int e = mem_check_geom(g)) return e;
  if (!st || !st->feat || !st->counters || !shadow) return BSCNAV_EINVAL;
  BSC_LAUNCH(k_build_shadow, mem_num_sms() * 8, kMemThreads, 0, (cudaStream_t)stream_, st->feat, st->counters, g->cache, g->dim,
             static_cast<__half*>(shadow));
  return BSCNAV_OK;
}

int64_t bscnav_query_batch_scratch_bytes(const bscnav_geom_t* g, int32_t n_query, int32_t k, int64_t n_rows_hint) {
  if (mem_check_geom(g) || n_query <= 0 || k <= 0 || k > 1024) return BSCNAV_EINVAL;
  return batch_layout(g, n_query, k, n_rows_hint, nullptr, nullptr);
}

int bscnav_query_topk_batch(const bscnav_geom_t* g, const bscnav_store_t* st, const void* shadow, const float* q, int32_t n_query,
                            int32_t k, const bscnav_filter_t* filter, int64_t n_rows_hint, void* scratch, int64_t scratch_bytes,
                            float* out_sim, int32_t* out_pos, int32_t* out_count, void* stream_) {
  if (int e = mem_check_geom(g)) return e;
  if (!st || !shadow || !q || !scratch || !out_sim || !out_pos || !out_count || n_query <= 0 || k <= 0 || k > 1024) return BSCNAV_EINVAL;
  BatchScratch s;
  if (batch_layout(g, n_query, k, n_rows_hint, &s, static_cast<char*>(scratch)) > scratch_bytes) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  bscnav_filter_t flt;
  if (filter) flt = *filter; else { flt.radius2 = -1; flt.cur[0] = flt.cur[1] = flt.cur[2] = 0; flt.floor_min = 1; flt.floor_max = 0; flt.reserved = 0; }
  const int D = g->dim, C = g->cache;
  const long long N = batch_rows(g, n_rows_hint);                  // GEMM columns: token rows, a multiple of 256
  const long long cap_rows = (long long)g->max_slots * C;
  const int slot_blocks = (int)((N / C + kSelSlots - 1) / kSelSlots);
  BSC_LAUNCH(k_row_ok, (unsigned)((cap_rows + kMemThreads - 1) / kMemThreads), kMemThreads, 0, stream, *g,
             reinterpret_cast<const unsigned long long*>(st->row_key), st->slot_cell, st->tok_h, st->counters, flt, s.row_ok);
  CUtensorMap map_w, map_wh;
  if (int e = make_operand_map(&map_w, shadow, N, D, gemm_block_n((int)N))) return e;
  if (int e = make_operand_map(&map_wh, shadow, N, D, 128)) return e;          // 128-row boxes: the CTA-pair GEMM (N % 256 == 0)
  for (int q0 = 0; q0 < n_query; q0 += kPassQueries) {
    const int P = n_query - q0 < kPassQueries ? n_query - q0 : kPassQueries;
    const float* qp = q + (long long)q0 * D;
    {
      // hist, cand_count and overflow are consecutive in the scratch layout: one span of words to clear
      const long long n_zero = (reinterpret_cast<char*>(s.overflow) + (size_t)P * 4 - reinterpret_cast<char*>(s.hist)) / 4;
      const int need = (P + kMemThreads / 32 - 1) / (kMemThreads / 32), fill = mem_num_sms() * 4;
      BSC_LAUNCH(k_prep_queries, need > fill ? need : fill, kMemThreads, 0, stream, qp, P, D, s.qh, s.qn2,
                 reinterpret_cast<uint32_t*>(s.hist), n_zero, s.keys, (long long)P * kCandCap);
    }
    CUtensorMap map_a;
    if (int e = make_operand_map(&map_a, s.qh, align_up(P, 128), D, 128)) return e;   // rows past P: never stored (M = P)
    GemmEpilogue epi{};
    epi.out = s.sims; epi.ld_out = N;
    if (C == 1) {
      // selection fused into the GEMM: masked fp16 similarities + fp32 block maxima from the epilogue, histogram over the maxima
      // (1/32 of the matrix), blocks opened only where a candidate can be
      const int n_blk = (int)(N / 32);
      CUtensorMap map_o;
      if (int e = make_store_map(&map_o, s.sims, P, N, false)) return e;                   // rows past P are clipped by the map
      epi.mode = kOutSims; epi.mask = s.row_ok; epi.aux = s.bmax; epi.ld_aux = n_blk;
      if (int e = gemm_f16_launch(&map_a, &map_w, &map_o, P, (int)N, D, epi, stream, &map_wh)) return e;
      const unsigned blk_blocks = (unsigned)((n_blk + kBlkChunk - 1) / kBlkChunk);
      BSC_LAUNCH_PDL(k_blk_hist, dim3(blk_blocks, (unsigned)P), kMemThreads, 0, stream, (const float*)s.bmax, (long long)n_blk, n_blk, s.hist);
      BSC_LAUNCH_PDL(k_blk_threshold, (unsigned)P, kMemThreads, 0, stream, (const uint32_t*)s.hist, k, s.thr);
      BSC_LAUNCH_PDL(k_blk_select, dim3(blk_blocks, (unsigned)P), kMemThreads, 0, stream, reinterpret_cast<const __half*>(s.sims), (long long)N,
                     (const float*)s.bmax, (long long)n_blk, n_blk, (const int32_t*)s.thr, s.cand, s.cand_count, s.overflow);
    } else {
      epi.mode = kOutFloat;
      if (int e = gemm_f16_launch(&map_a, &map_w, nullptr, P, (int)N, D, epi, stream)) return e;
      BSC_LAUNCH(k_batch_hist, dim3((unsigned)slot_blocks, (unsigned)P), kMemThreads, 0, stream, s.sims, N, s.row_ok, st->counters, C,
                 s.hist);
      BSC_LAUNCH(k_batch_select, dim3((unsigned)slot_blocks, (unsigned)P), kMemThreads, 0, stream, s.sims, N, s.row_ok, st->counters,
                 C, s.hist, k, s.cand, s.cand_count, s.overflow);
    }
    BSC_LAUNCH_PDL(k_batch_rescore, dim3(16, (unsigned)P), kMemThreads, 0, stream, *g, (const float*)st->feat, (const float*)st->norm2,
                   (const uint8_t*)s.row_ok, (const int32_t*)st->slot_cell, qp, (const float*)s.qn2, (const int32_t*)s.cand,
                   (const int32_t*)s.cand_count, (const int32_t*)s.overflow, s.keys);
    BSC_LAUNCH_PDL(k_sort_candidates, (unsigned)P, 1024, 0, stream, (const unsigned long long*)s.keys, (const int32_t*)s.cand_count,
                   (const int32_t*)s.overflow, k, s.sorted, (int32_t*)nullptr);
    if (int e = launch_topk_decode(g, st, s.sorted, k, k, qp, P, out_sim + (long long)q0 * k, out_pos + (long long)q0 * k * 3,
                                   out_count + q0, stream)) return e;
    BSC_LAUNCH_PDL(k_mark_overflow, (P + 255) / 256, 256, 0, stream, (const int32_t*)s.overflow, P, out_count + q0);
  }
  return BSCNAV_OK;
}

}  // extern "C"
