// Declarations shared by memory_kernels.cu and query_batch.cu (query key encoding, top-K stages, helpers).
#pragma once
#include "common.cuh"

namespace bscnav {

static constexpr int kMemThreads = 256;
static constexpr int kTopkChunk = 4096;     // keys one k_topk_chunks block sorts

// descending similarity <-> ascending unsigned: composite key (~orderable(sim) << 32) | cell
__device__ __forceinline__ uint32_t orderable(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

// candidate filter of voxel_localized (memory_2.py:624-640) for one token of a cell
__device__ __forceinline__ bool filter_pass(const bscnav_filter_t& flt, int row, int col, int h) {
  bool ok = true;
  if (flt.radius2 >= 0) {
    const long long dr = row - flt.cur[0], dc = col - flt.cur[1], dh = h - flt.cur[2];
    ok = dr * dr + dc * dc + dh * dh <= flt.radius2;
  }
  // explicit enable bit: an empty range from load_single_floor (min > max) must select nothing, not switch the filter off
  if ((flt.reserved & 1) || flt.floor_min <= flt.floor_max) ok = ok && h >= flt.floor_min && h <= flt.floor_max;
  return ok;
}

// host launchers of the shared top-K stages (kernels live in memory_kernels.cu)
//   topk_chunks: for each of n_query key rows, sort chunks of kTopkChunk keys and keep the k smallest of each chunk
//   topk_decode: keys -> (similarity, row, col, h) and the per-query result count
int launch_topk_chunks(const unsigned long long* in, long long in_stride, const int32_t* n_dev, int n_host, int n_chunks,
                       int n_query, unsigned long long* out, long long out_stride, int k, cudaStream_t stream);
int launch_topk_decode(const bscnav_geom_t* g, const bscnav_store_t* st, const unsigned long long* keys, long long key_stride, int k,
                       const float* q, int n_query, float* out_sim, int32_t* out_pos, int32_t* out_count, cudaStream_t stream);
// histogram-select + sort of exact composite keys (query_batch.cu), used by the single-query path
int64_t exact_select_scratch_bytes(int n_query);
int64_t exact_select_zero_bytes(int n_query);     // leading bytes of that scratch that must be zero on entry (histogram, counters, flags)
// zeroed: the caller cleared those bytes already (k_query_slots does it in its prologue); zero_count: int32[n_query] cleared by the
// sort stage for the decode stage's atomic count, or null
int launch_exact_select(const unsigned long long* keys, long long stride, const int32_t* counters, int max_slots, int n_query, int k,
                        void* scratch, unsigned long long* sorted, int32_t** overflow_out, cudaStream_t stream, bool zeroed = false,
                        int32_t* zero_count = nullptr);
int mem_num_sms();
int mem_check_geom(const bscnav_geom_t* g);

}  // namespace bscnav
