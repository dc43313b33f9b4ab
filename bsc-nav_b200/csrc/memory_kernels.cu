// Spatial-memory build + query kernels (sm_100a): depth unprojection + pose transform + voxel index (float64),
// deterministic slot assignment, top-C winner selection with 64-bit atomics, 128-bit feature-row commit,
// cosine / per-cell max / top-K query.  HBM-bound integer/byte work: coalesced 16-byte records, warp-level
// pre-aggregation in front of the atomics, grid sizes in multiples of the SM count.
//
// Reference semantics restated here (each kernel cites its lines): memory_2.py:744-756, 842-894, 326-358,
// 623-671; utils.py:153-214.
#include <cstdlib>
#include "memory_kernels.cuh"

namespace bscnav {

thread_local char g_last_error[256] = {0};
std::atomic<long long> g_launches{0};
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("BSCNAV_PDL"); on = e ? (atoi(e) != 0) : 1; }
  return on != 0;
}

static constexpr int kThreads = 256;
static constexpr int kItems = 4;                 // points per thread in the ordered-scan kernels
static constexpr int kTile = kThreads * kItems;  // 1024 points per block
static constexpr uint32_t kNoTouch = 0xFFFFFFFFu;
static constexpr int kSrcBits = 20;              // record.aux = src_row | (h << 20)
static constexpr uint32_t kSrcMask = (1u << kSrcBits) - 1;

// One record per visited point (16 B, written/read as one 128-bit access):
//   x = linear cell or -1 (rejected)   y = aux (token source row | h << 20)   z = ordinal   w = float bits of r^2
// key = (w << 32) | z : smaller is better (r^2 >= 0 so its float bits order as unsigned), ties -> lower ordinal.
__device__ __forceinline__ uint64_t rec_key(const uint4& r) { return ((uint64_t)r.w << 32) | r.z; }

// row-vector times column with the FMA chain numpy/OpenBLAS dgemm uses (k ascending, accumulator starts at a0*b0):
// reproduces `cam_mat_inv @ p_2d`, `pose @ pc_homo`, `cam_mat @ p` of utils.py:168,197,209 bit for bit.
__device__ __forceinline__ double dot3(const double* m, double a, double b, double c) {
  return __fma_rn(m[2], c, __fma_rn(m[1], b, __dmul_rn(m[0], a)));
}
__device__ __forceinline__ double dot4(const double* m, double a, double b, double c, double d) {
  return __fma_rn(m[3], d, __fma_rn(m[2], c, __fma_rn(m[1], b, __dmul_rn(m[0], a))));
}
// python int(x) on a float64: truncation toward zero, saturated so the range test below rejects it
__device__ __forceinline__ int trunc_i32(double v) {
  if (!(v > -2.0e9)) return -2000000000;
  if (!(v < 2.0e9)) return 2000000000;
  return (int)v;
}

// int(x / cs) of utils.py:201-205 without the IEEE division on the common path: t = x * (1 / cs) differs from the correctly
// rounded quotient by at most 3 ulp, so unless t lies within 8 ulp (|t| 2^-49) of an integer both truncate to the same value;
// next to an integer the exact division decides.  Verified against division + truncation on 10^8 inputs including planted
// cell boundaries (tools/micro/div_trunc_check.py, tests/test_kernel_arith_emulation.py).
__device__ __forceinline__ int trunc_div_cs(double x, double cs, double rcs) {
  const double t = __dmul_rn(x, rcs);
  if (fabs(t - rint(t)) > fabs(t) * 0x1p-49) return trunc_i32(t);
  return trunc_i32(__ddiv_rn(x, cs));
}
// int(q0 / q2 - 0.5) of project_point (utils.py:208-213) with one reciprocal r = 1 / q2 shared by both coordinates; same guard
__device__ __forceinline__ int trunc_quot_half(double q0, double q2, double r) {
  const double t = __dmul_rn(q0, r);
  const double u = __dsub_rn(t, 0.5);
  if (fabs(u - rint(u)) > fmax(fabs(t), 1.0) * 0x1p-48) return trunc_i32(u);
  return trunc_i32(__dsub_rn(__ddiv_rn(q0, q2), 0.5));
}

// ---------------------------------------------------------------------------------------------------------
// K1: unproject + pose + voxel index + patch projection + winner key.          (a3-a7 of SURVEY 8a)
// fresh[i / 32] bit i % 32 = point i heads a run of points on a cell that has no slot yet: only these points can be the first
// toucher of a new cell, so the slot-assignment passes read this 1-bit-per-point mask (2 MB per 54-frame batch) and open the
// 16-byte record of a point only where its bit is set, instead of streaming all 270 MB of records twice.
__global__ void __launch_bounds__(kThreads)
k_index_points(const bscnav_geom_t g, const double rcs, const float* __restrict__ depth, const double* __restrict__ poses,
               const uint32_t* __restrict__ ord_base, const int32_t* __restrict__ sample_idx, int n_points,
               long long total, uint4* __restrict__ rec, const int32_t* __restrict__ occupied,
               uint32_t* __restrict__ first_touch, uint32_t* __restrict__ fresh) {
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int cell = -1;
  uint32_t aux = 0, ord = 0, r2bits = 0;
  if (i < total) {
    const int b = (int)(i / n_points);
    const int j = (int)(i - (long long)b * n_points);
    const int pix = sample_idx ? sample_idx[i] : j;
    const int hw = g.height * g.width;
    ord = ord_base[b] + (uint32_t)j;
    const double z = (double)depth[(long long)b * hw + pix];
    const int v = pix / g.width, u = pix - v * g.width;
    const double x = (double)u + 0.5, y = (double)v + 0.5;
    // utils.py:166-171  pc = (cam_mat_inv @ [x,y,1]) * z
    const double p0 = __dmul_rn(dot3(g.kinv + 0, x, y, 1.0), z);
    const double p1 = __dmul_rn(dot3(g.kinv + 3, x, y, 1.0), z);
    const double p2 = __dmul_rn(dot3(g.kinv + 6, x, y, 1.0), z);
    if (p2 > g.min_depth && p2 < g.max_depth) {                   // utils.py:174-176 (strict)
      const double* M = poses + (long long)b * 16;
      const double gx = dot4(M + 0, p0, p1, p2, 1.0);             // utils.py:195-197
      const double gy = dot4(M + 4, p0, p1, p2, 1.0);
      const double gz = dot4(M + 8, p0, p1, p2, 1.0);
      const double half = (double)g.gs / 2.0;
      // utils.py:201-205  row = int(gs/2 - int(x/cs)) ...
      const int row = trunc_i32(__dsub_rn(half, (double)trunc_div_cs(gx, g.cs, rcs)));
      const int col = trunc_i32(__dsub_rn(half, (double)trunc_div_cs(gy, g.cs, rcs)));
      const int h = trunc_div_cs(gz, g.cs, rcs);
      const bool out = col >= g.gs || row >= g.gs || h >= g.maxh || col < 0 || row < 0 || h < g.minh;  // memory_2.py:755-756
      if (!out) {
        // utils.py:208-214 with the patch intrinsics (memory_2.py:857,871)
        const double q0 = dot3(g.kpatch + 0, p0, p1, p2), q1 = dot3(g.kpatch + 3, p0, p1, p2),
                     q2 = dot3(g.kpatch + 6, p0, p1, p2);
        const double rq = __drcp_rn(q2);
        const int ppx = trunc_quot_half(q0, q2, rq);
        const int ppy = trunc_quot_half(q1, q2, rq);
        if (!(ppx < 0 || ppy < 0 || ppx >= g.n_patch_h || ppy >= g.n_patch_w)) {   // memory_2.py:878 (token array is (W', H', D))
          const int hh = h - g.minh;                                               // memory_2.py:867
          cell = (g.levels == 1) ? row * g.gs + col : (row * g.gs + col) * g.levels + hh;
          const uint32_t src = (uint32_t)(b * (g.n_patch_h * g.n_patch_w) + ppy * g.n_patch_h + ppx);
          aux = src | ((uint32_t)hh << kSrcBits);
          // memory_2.py:873,885  float32(np.sum(np.square(p_local)))
          const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(p0, p0), __dmul_rn(p1, p1)), __dmul_rn(p2, p2));
          r2bits = __float_as_uint(__double2float_rn(r2));
        }
      }
    }
    rec[i] = make_uint4((uint32_t)cell, aux, ord, r2bits);
  }
  // first-touch ordinal of still-unassigned cells; within a warp only the head of a run of equal cells issues
  // the atomic (ordinals ascend with the lane, so the head holds the run's minimum)
  const int prev = __shfl_up_sync(0xffffffffu, cell, 1);
  const bool fresh_head = cell >= 0 && (lane == 0 || prev != cell) && occupied[cell] < 0;
  if (fresh_head) atomicMin(&first_touch[cell], ord);
  const unsigned fm = __ballot_sync(0xffffffffu, fresh_head);
  if (lane == 0 && i < total) fresh[i >> 5] = fm;
}

// block-wide exclusive scan of one int per thread (kThreads threads); returns exclusive prefix, total via *sum
__device__ __forceinline__ int block_excl_scan(int v, int* sum) {
  __shared__ int warp_tot[kThreads / 32];
  __shared__ int block_tot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    int t = lane < kThreads / 32 ? warp_tot[lane] : 0;
    int ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int s = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += s;
    }
    if (lane < kThreads / 32) warp_tot[lane] = ti - t;
    if (lane == 31) block_tot = ti;
  }
  __syncthreads();
  *sum = block_tot;
  const int r = warp_tot[w] + inc - v;
  __syncthreads();
  return r;
}

__device__ __forceinline__ bool is_first_toucher(const uint4& r, const int32_t* occupied, const uint32_t* first_touch) {
  const int cell = (int)r.x;
  return cell >= 0 && occupied[cell] < 0 && first_touch[cell] == r.z;
}

// K2a: number of newly touched cells per 1024-point tile (a point is the "first toucher" of an unassigned cell)
// the four mask bits of a thread's points (kItems consecutive points from `base`, a multiple of 4); null mask: every point
__device__ __forceinline__ unsigned fresh_bits(const uint32_t* __restrict__ fresh, long long base) {
  return fresh ? (fresh[base >> 5] >> (base & 31)) & 0xFu : 0xFu;
}

__global__ void __launch_bounds__(kThreads)
k_count_new(const uint4* __restrict__ rec, long long total, const int32_t* __restrict__ occupied,
            const uint32_t* __restrict__ first_touch, int* __restrict__ tile_count, const uint32_t* __restrict__ fresh) {
  static_assert(kItems == 4, "fresh_bits extracts four bits per thread");
  const long long base = (long long)blockIdx.x * kTile + threadIdx.x * kItems;
  const unsigned bits = fresh_bits(fresh, base);
  int c = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k)
    if (((bits >> k) & 1u) && base + k < total) c += is_first_toucher(rec[base + k], occupied, first_touch) ? 1 : 0;
  int tot;
  block_excl_scan(c, &tot);
  if (threadIdx.x == 0) tile_count[blockIdx.x] = tot;
}

// K2a / K2c with k_index_points' mask: one WARP per 1024-point tile, lane l owns mask word l (32 consecutive points).  In steady
// state a batch of 16.9 M points has ~10^3 first touches: nearly every word is zero, and a block per tile with its block-wide
// scans (the unmasked forms above / below, still used by the merge path) cost 48 + 21 us in launches and barriers alone.
__device__ __forceinline__ uint32_t first_toucher_bits(uint32_t word, long long word_base, long long total, const uint4* __restrict__ rec,
                                                       const int32_t* __restrict__ occupied, const uint32_t* __restrict__ first_touch) {
  uint32_t out = 0;
  while (word) {
    const int b = __ffs(word) - 1;
    word &= word - 1;
    const long long i = word_base + b;
    if (i < total && is_first_toucher(rec[i], occupied, first_touch)) out |= 1u << b;
  }
  return out;
}

__global__ void __launch_bounds__(kThreads)
k_count_new_masked(const uint4* __restrict__ rec, long long total, const int32_t* __restrict__ occupied,
                   const uint32_t* __restrict__ first_touch, int* __restrict__ tile_count, const uint32_t* __restrict__ fresh,
                   int n_tiles) {
  const int tile = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (tile >= n_tiles) return;
  const uint32_t word = fresh[(long long)tile * (kTile / 32) + lane];
  int c = 0;
  if (__any_sync(0xffffffffu, word != 0u)) {
    c = __popc(first_toucher_bits(word, (long long)tile * kTile + lane * 32, total, rec, occupied, first_touch));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if (lane == 0) tile_count[tile] = c;
}

__global__ void __launch_bounds__(kThreads)
k_assign_slots_masked(const uint4* __restrict__ rec, long long total, int32_t* __restrict__ occupied,
                      const uint32_t* __restrict__ first_touch, const int* __restrict__ tile_off, const int* __restrict__ hdr,
                      int32_t* __restrict__ slot_cell, const uint32_t* __restrict__ fresh, const int* __restrict__ tile_count,
                      int n_tiles) {
  const int tile = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (tile >= n_tiles || !hdr[1] || tile_count[tile] == 0) return;
  const long long word_base = (long long)tile * kTile + lane * 32;
  uint32_t ft = first_toucher_bits(fresh[(long long)tile * (kTile / 32) + lane], word_base, total, rec, occupied, first_touch);
  const int c = __popc(ft);
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  int rank = tile_off[tile] + hdr[0] + inc - c;
  __syncwarp();                         // every flag of the tile is read before any cell of it is marked (see k_assign_slots)
  while (ft) {
    const int b = __ffs(ft) - 1;
    ft &= ft - 1;
    const int cell = (int)rec[word_base + b].x;
    slot_cell[rank] = cell;
    occupied[cell] = rank;
    ++rank;
  }
}

// K2b: exclusive scan of the tile counts (single block), slot-range reservation, capacity check.
// One block of 1024 threads, every thread owns a contiguous run of tiles: two serial passes over the run around one block-wide
// scan of the 1024 partial sums (a loop of 256-wide block scans took 57 us for the 16 500 tiles of a 54-frame batch).
static constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads)
k_scan_tiles(const int* __restrict__ tile_count, int* __restrict__ tile_off, int n_tiles, int* __restrict__ hdr,
             int32_t* __restrict__ counters, int max_slots) {
  __shared__ int warp_tot[kScanThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int per = (n_tiles + kScanThreads - 1) / kScanThreads;
  const int t0 = tid * per, t1 = min(t0 + per, n_tiles);
  int sum = 0;
  for (int t = t0; t < t1; ++t) sum += tile_count[t];
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    const int v = warp_tot[lane];
    int vi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, vi, o);
      if (lane >= o) vi += u;
    }
    warp_tot[lane] = vi - v;                      // exclusive prefix of the warp totals
    if (lane == 31) {
      const int total = vi;
      const int base_slot = counters[0];
      const bool ok = (long long)base_slot + total <= (long long)max_slots;
      hdr[0] = base_slot;
      hdr[1] = ok ? 1 : 0;
      if (ok) counters[0] = base_slot + total; else counters[1] = 1;   // overflow: nothing from this batch is assigned
    }
  }
  __syncthreads();
  int run = warp_tot[w] + inc - sum;
  for (int t = t0; t < t1; ++t) { tile_off[t] = run; run += tile_count[t]; }
}

// K2c: slots in first-touch order == the reference's sequential numbering (memory_2.py:888-894)
__global__ void __launch_bounds__(kThreads)
k_assign_slots(const uint4* __restrict__ rec, long long total, int32_t* __restrict__ occupied,
               const uint32_t* __restrict__ first_touch, const int* __restrict__ tile_off,
               const int* __restrict__ hdr, int32_t* __restrict__ slot_cell, const uint32_t* __restrict__ fresh,
               const int* __restrict__ tile_count) {
  if (!hdr[1] || tile_count[blockIdx.x] == 0) return;
  const long long base = (long long)blockIdx.x * kTile + threadIdx.x * kItems;
  const unsigned bits = fresh_bits(fresh, base);
  bool f[kItems];
  int cells[kItems];
  int c = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    f[k] = false;
    cells[k] = -1;
    if (((bits >> k) & 1u) && base + k < total) {
      const uint4 r = rec[base + k];
      f[k] = is_first_toucher(r, occupied, first_touch);
      cells[k] = (int)r.x;
    }
    c += f[k] ? 1 : 0;
  }
  int tot;
  int rank = block_excl_scan(c, &tot) + tile_off[blockIdx.x] + hdr[0];
  // note: `occupied` is written below while other threads may still be reading it in is_first_toucher above:
  // safe because a flagged cell is written by exactly one point (its first toucher) and only cells whose
  // first_touch ordinal equals the reader's own ordinal can flip the reader's flag -- i.e. its own cell.
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kItems; ++k)
    if (f[k]) {
      slot_cell[rank] = cells[k];
      occupied[cells[k]] = rank;
      ++rank;
    }
}

// K3: top-C winner selection.  C == 1: warp-segmented min then one 64-bit atomicMin per run of equal cells.
// C > 1: lock-free sorted insertion -- offer the key to slot position 0..C-1 with atomicMin and carry the
// displaced key forward; position j ends up holding the (j+1)-th smallest key whatever the arrival order.
// Every key that is offered to the store is also written, compacted per block, into the block's own segment of a candidate
// array {slot, aux, key} (no global counter: one contended atomic per warp cost 100 us): the commit pass then looks at these few
// entries instead of re-reading the whole record array (16 B x 16.6 M points per 54-frame batch).
__global__ void __launch_bounds__(kThreads)
k_select(const uint4* __restrict__ rec, long long total, const int32_t* __restrict__ occupied,
         uint32_t* __restrict__ first_touch, unsigned long long* __restrict__ slot_keys, int C,
         uint4* __restrict__ cand, int* __restrict__ cand_n, const uint32_t* __restrict__ fresh) {
  __shared__ int warp_off[kThreads / 32 + 1];
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int cell = -1, slot = -1;
  uint32_t aux = 0;
  unsigned long long key = BSCNAV_KEY_EMPTY;
  // first_touch holds an ordinal only where k_index_points issued its atomicMin: at the points of its mask.  Only those restore
  // the scratch invariant for the next batch (was: every point, 16.9 M scattered 4-byte stores per 55 frames); without a mask
  // (merge path) every record does.
  const uint32_t fword = (fresh && i - lane < total) ? fresh[i >> 5] : 0xFFFFFFFFu;
  if (i < total) {
    const uint4 r = rec[i];
    cell = (int)r.x;
    if (cell >= 0) {
      if ((fword >> lane) & 1u) first_touch[cell] = kNoTouch;
      key = rec_key(r);
      aux = r.y;
    }
  }
  bool offered = false;
  if (C == 1) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long k2 = __shfl_down_sync(0xffffffffu, key, d);
      const int c2 = __shfl_down_sync(0xffffffffu, cell, d);
      const uint32_t a2 = __shfl_down_sync(0xffffffffu, aux, d);
      if (lane + d < 32 && c2 == cell && k2 < key) { key = k2; aux = a2; }
    }
    const int prev = __shfl_up_sync(0xffffffffu, cell, 1);
    if (cell >= 0 && (lane == 0 || prev != cell)) {      // the head of a run of equal cells holds the run's minimum: it alone
      slot = occupied[cell];                              // looks its slot up (< 0 only after a capacity overflow)
      if (slot >= 0 && key < slot_keys[slot]) {
        atomicMin(&slot_keys[slot], key);
        offered = true;
      }
    }
  } else if (cell >= 0 && (slot = occupied[cell]) >= 0) {
    unsigned long long* ks = slot_keys + (long long)slot * C;
    if (key < ks[C - 1]) {
      offered = true;
      unsigned long long k = key;
      for (int j = 0; j < C; ++j) {
        const unsigned long long old = atomicMin(&ks[j], k);
        if (old == k) break;                   // already stored (the same point offered again): idempotent
        if (old > k) {
          k = old;
          if (k == BSCNAV_KEY_EMPTY) break;
        }
      }
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, offered);
  const int w = threadIdx.x >> 5;
  if (lane == 0) warp_off[w + 1] = __popc(m);
  if (threadIdx.x == 0) warp_off[0] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 1; k <= kThreads / 32; ++k) warp_off[k] += warp_off[k - 1];
    cand_n[blockIdx.x] = warp_off[kThreads / 32];
  }
  __syncthreads();
  if (offered)
    cand[(long long)blockIdx.x * kThreads + warp_off[w] + __popc(m & ((1u << lane) - 1))] =
        make_uint4((uint32_t)slot, aux, (uint32_t)key, (uint32_t)(key >> 32));
}

// K4: winners copy their 4 KB token row into the store with 128-bit loads/stores, one warp per row.
// Decide for one offered key whether it is kept and which physical row it takes (-1: not a winner).
__device__ __forceinline__ long long claim_row(int slot, unsigned long long key, const unsigned long long* __restrict__ slot_keys,
                                               unsigned long long* __restrict__ row_key, int C) {
  const unsigned long long* ks = slot_keys + (long long)slot * C;
  if (C == 1) {
    if (ks[0] != key) return -1;
    row_key[slot] = key;
    return slot;
  }
  bool win = false;
  for (int j = 0; j < C && ks[j] <= key; ++j) win |= (ks[j] == key);
  if (!win) return -1;
  // claim a physical row whose current token is no longer among the kept keys (or was never used)
  unsigned long long* rk = row_key + (long long)slot * C;
  for (int rr = 0; rr < C; ++rr)         // the token may already own a row (same point offered again)
    if (rk[rr] == key) return -1;
  for (int rr = 0; rr < C; ++rr) {
    const unsigned long long cur = rk[rr];
    bool kept = false;
    for (int j = 0; j < C; ++j) kept |= (ks[j] == cur);
    if (cur == BSCNAV_KEY_EMPTY) kept = false;
    if (!kept && atomicCAS(&rk[rr], cur, key) == cur) return (long long)slot * C + rr;
  }
  return -1;
}
// the warp copies the rows its lanes won, one after the other; norm2 / tok_h / the committed-row counter follow
__device__ __forceinline__ void copy_won_rows(long long dst_row, uint32_t aux, int32_t* __restrict__ tok_h, float* __restrict__ norm2,
                                              float* __restrict__ feat, const float* __restrict__ tokens, long long tok_stride, int D,
                                              int32_t* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  unsigned winners = __ballot_sync(0xffffffffu, dst_row >= 0);
  const int nvec = D >> 2;
  if (lane == 0 && winners) atomicAdd(&counters[3], __popc(winners));     // rows committed since the last reset (wraps)
  while (winners) {
    const int src_lane = __ffs(winners) - 1;
    winners &= winners - 1;
    const long long row = __shfl_sync(0xffffffffu, dst_row, src_lane);
    const uint32_t a = __shfl_sync(0xffffffffu, aux, src_lane);
    const float4* s = reinterpret_cast<const float4*>(tokens + (long long)(a & kSrcMask) * tok_stride);
    float4* d = reinterpret_cast<float4*>(feat + row * D);
    float acc = 0.f;
    for (int c = lane; c < nvec; c += 32) {
      const float4 v = s[c];
      d[c] = v;
      acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) { norm2[row] = acc; tok_h[row] = (int)(a >> kSrcBits); }
  }
}

// one warp per k_select block: its segment of offered keys (usually a handful, at most 256)
__global__ void __launch_bounds__(kThreads)
k_commit(const uint4* __restrict__ cand, const int* __restrict__ cand_n, int n_segments,
         const unsigned long long* __restrict__ slot_keys, unsigned long long* __restrict__ row_key,
         int32_t* __restrict__ tok_h, float* __restrict__ norm2, float* __restrict__ feat,
         const float* __restrict__ tokens, long long tok_stride, int C, int D, int32_t* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (kThreads / 32);
  for (int seg = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); seg < n_segments; seg += warps) {
    const int n = cand_n[seg];
    for (int i0 = 0; i0 < n; i0 += 32) {
      long long dst_row = -1;
      uint32_t aux = 0;
      if (i0 + lane < n) {
        const uint4 c = cand[(long long)seg * kThreads + i0 + lane];
        aux = c.y;
        dst_row = claim_row((int)c.x, ((unsigned long long)c.w << 32) | c.z, slot_keys, row_key, C);
      }
      copy_won_rows(dst_row, aux, tok_h, norm2, feat, tokens, tok_stride, D, counters);
    }
  }
}

// squared norms of rows [0, n_rows) with the association k_commit uses (lane-strided fma, shuffle tree), so a store
// filled from disk scores bit-identically to the store that was saved
__global__ void __launch_bounds__(kThreads)
k_row_norms(const float* __restrict__ feat, float* __restrict__ norm2, long long n_rows, int D) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (kThreads / 32);
  for (long long row = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); row < n_rows; row += warps) {
    const float4* s = reinterpret_cast<const float4*>(feat + row * D);
    float acc = 0.f;
    for (int c = lane; c < (D >> 2); c += 32) {
      const float4 v = ld_stream_f4(s + c);
      acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) norm2[row] = acc;
  }
}

__global__ void k_fill_u32(uint32_t* p, long long n, uint32_t v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_fill_u64(unsigned long long* p, long long n, unsigned long long v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// records for a compacted delta (multi-GPU merge): cell, source row i | h << 20, key
__global__ void k_delta_records(const int32_t* __restrict__ d_cell, const unsigned long long* __restrict__ d_key,
                                const int32_t* __restrict__ d_h, int n, uint4* __restrict__ rec,
                                const int32_t* __restrict__ occupied, uint32_t* __restrict__ first_touch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cell = d_cell[i];
  const unsigned long long key = d_key[i];
  rec[i] = make_uint4((uint32_t)cell, (uint32_t)i | ((uint32_t)d_h[i] << kSrcBits), (uint32_t)key, (uint32_t)(key >> 32));
  if (cell >= 0 && occupied[cell] < 0) atomicMin(&first_touch[cell], (uint32_t)key);
}

// ordered compaction of the rows whose ordinal lies in [lo, hi): count / scan / write (same 3-kernel shape as K2)
__device__ __forceinline__ bool row_in_range(const unsigned long long* row_key, long long r, uint32_t lo, uint32_t hi) {
  const unsigned long long k = row_key[r];
  if (k == BSCNAV_KEY_EMPTY) return false;
  const uint32_t o = (uint32_t)k;
  return o >= lo && o < hi;
}
__global__ void __launch_bounds__(kThreads)
k_export_count(const unsigned long long* __restrict__ row_key, const int32_t* __restrict__ counters, int C, uint32_t lo,
               uint32_t hi, int* __restrict__ tile_count) {
  const long long n_rows = (long long)counters[0] * C;
  const long long base = (long long)blockIdx.x * kTile + threadIdx.x * kItems;
  int c = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k)
    if (base + k < n_rows) c += row_in_range(row_key, base + k, lo, hi) ? 1 : 0;
  int tot;
  block_excl_scan(c, &tot);
  if (threadIdx.x == 0) tile_count[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(kThreads)
k_export_scan(const int* __restrict__ tile_count, int* __restrict__ tile_off, int n_tiles, int* __restrict__ out_n, int max_out) {
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += kThreads) {
    const int idx = base + threadIdx.x;
    const int v = idx < n_tiles ? tile_count[idx] : 0;
    int tot;
    const int ex = block_excl_scan(v, &tot);
    if (idx < n_tiles) tile_off[idx] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) { out_n[0] = carry < max_out ? carry : max_out; out_n[1] = carry; }   // [1] > max_out: truncated
}
__global__ void __launch_bounds__(kThreads)
k_export_write(const unsigned long long* __restrict__ row_key, const int32_t* __restrict__ counters,
               const int32_t* __restrict__ slot_cell, const int32_t* __restrict__ tok_h, const float* __restrict__ feat,
               int C, int D, uint32_t lo, uint32_t hi, const int* __restrict__ tile_off, int32_t* __restrict__ d_cell,
               unsigned long long* __restrict__ d_key, int32_t* __restrict__ d_h, float* __restrict__ d_feat, int max_out) {
  const long long n_rows = (long long)counters[0] * C;
  const long long base = (long long)blockIdx.x * kTile + threadIdx.x * kItems;
  bool f[kItems];
  int c = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    f[k] = base + k < n_rows && row_in_range(row_key, base + k, lo, hi);
    c += f[k] ? 1 : 0;
  }
  int tot;
  int rank = block_excl_scan(c, &tot) + tile_off[blockIdx.x];
  long long my_dst = -1, my_src = -1;   // each thread publishes its rows one at a time for a warp-wide copy
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    my_dst = -1;
    if (f[k]) {
      if (rank < max_out) {
        my_dst = rank;
        my_src = base + k;
        d_cell[rank] = slot_cell[(base + k) / C];
        d_key[rank] = row_key[base + k];
        d_h[rank] = tok_h[base + k];
      }
      ++rank;
    }
    unsigned m = __ballot_sync(0xffffffffu, my_dst >= 0);
    while (m) {
      const int sl = __ffs(m) - 1;
      m &= m - 1;
      const long long dr = __shfl_sync(0xffffffffu, my_dst, sl), sr = __shfl_sync(0xffffffffu, my_src, sl);
      const float4* s = reinterpret_cast<const float4*>(feat + sr * D);
      float4* d = reinterpret_cast<float4*>(d_feat + dr * D);
      for (int q = lane; q < (D >> 2); q += 32) d[q] = s[q];
    }
  }
}

// Packed delta rows (one buffer for the all-gather merge): { int32 cell; int32 h; uint64 key; float feat[D] }, 16 + 4 D bytes
struct PackedHdr { int32_t cell; int32_t h; unsigned long long key; };

__global__ void __launch_bounds__(kThreads)
k_export_packed(const unsigned long long* __restrict__ row_key, const int32_t* __restrict__ counters,
                const int32_t* __restrict__ slot_cell, const int32_t* __restrict__ tok_h, const float* __restrict__ feat,
                int C, int D, uint32_t lo, uint32_t hi, const int* __restrict__ tile_off, char* __restrict__ packed, int max_out) {
  const long long n_rows = (long long)counters[0] * C;
  const long long base = (long long)blockIdx.x * kTile + threadIdx.x * kItems;
  const long long row_bytes = 16 + 4ll * D;
  bool f[kItems];
  int c = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    f[k] = base + k < n_rows && row_in_range(row_key, base + k, lo, hi);
    c += f[k] ? 1 : 0;
  }
  int tot;
  int rank = block_excl_scan(c, &tot) + tile_off[blockIdx.x];
  long long my_dst = -1, my_src = -1;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    my_dst = -1;
    if (f[k]) {
      if (rank < max_out) {
        my_dst = rank;
        my_src = base + k;
        PackedHdr h;
        h.cell = slot_cell[(base + k) / C]; h.h = tok_h[base + k]; h.key = row_key[base + k];
        *reinterpret_cast<uint4*>(packed + rank * row_bytes) = *reinterpret_cast<const uint4*>(&h);
      }
      ++rank;
    }
    unsigned m = __ballot_sync(0xffffffffu, my_dst >= 0);
    while (m) {
      const int sl = __ffs(m) - 1;
      m &= m - 1;
      const long long dr = __shfl_sync(0xffffffffu, my_dst, sl), sr = __shfl_sync(0xffffffffu, my_src, sl);
      const float4* s = reinterpret_cast<const float4*>(feat + sr * D);
      float4* d = reinterpret_cast<float4*>(packed + dr * row_bytes + 16);
      for (int q = lane; q < (D >> 2); q += 32) d[q] = s[q];
    }
  }
}

// records of a chunk [first, first + n) of the gathered packed buffer: n_segments segments of seg_capacity rows, segment s holds
// seg_counts[s * count_stride] valid rows; skip_segment (this rank's own rows, already in the store) is ignored
__global__ void k_packed_records(const char* __restrict__ packed, long long row_bytes, long long first, int n, int seg_capacity,
                                 const int32_t* __restrict__ seg_counts, int count_stride, int skip_segment, uint4* __restrict__ rec,
                                 const int32_t* __restrict__ occupied, uint32_t* __restrict__ first_touch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = first + i;
  const int seg = (int)(r / seg_capacity), j = (int)(r - (long long)seg * seg_capacity);
  int cell = -1;
  PackedHdr h;
  h.cell = -1; h.h = 0; h.key = 0;
  if (seg != skip_segment && j < seg_counts[seg * count_stride]) {
    *reinterpret_cast<uint4*>(&h) = *reinterpret_cast<const uint4*>(packed + r * row_bytes);
    cell = h.cell;
  }
  rec[i] = make_uint4((uint32_t)cell, (uint32_t)i | ((uint32_t)h.h << kSrcBits), (uint32_t)h.key, (uint32_t)(h.key >> 32));
  if (cell >= 0 && occupied[cell] < 0) atomicMin(&first_touch[cell], (uint32_t)h.key);
}

// ---------------------------------------------------------------------------------------------------------
// Query.  Q1: one warp per slot: cosine of up to kQ queries against the slot's C rows, per-slot max,
// candidate filter, composite sort key  (~orderable(sim) << 32) | cell  -> ascending key == descending sim,
// ties by ascending (row, col, h) because cell is the row-major linear index.
static constexpr int kQ = 8;
template <int NV>   // NV = D / 128 float4 per lane
__global__ void __launch_bounds__(kThreads)
k_query_slots(const bscnav_geom_t g, const float* __restrict__ feat, const float* __restrict__ norm2,
              const unsigned long long* __restrict__ row_key, const int32_t* __restrict__ slot_cell,
              const int32_t* __restrict__ tok_h, const int32_t* __restrict__ counters, const float* __restrict__ q,
              int q0, int nq, const bscnav_filter_t flt, unsigned long long* __restrict__ keys, long long key_stride,
              uint32_t* __restrict__ zero_words, int n_zero) {
  extern __shared__ float4 sq[];            // nq * D floats
  __shared__ float q_n2[kQ];
  // the histogram / counters of the select stages behind this kernel start from zero: cleared here instead of by three memsets
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n_zero; i += gridDim.x * kThreads) zero_words[i] = 0u;
  const int D = NV * 128;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < nq * (D >> 2); i += kThreads) sq[i] = reinterpret_cast<const float4*>(q + (long long)q0 * D)[i];
  __syncthreads();
  if (wib < nq) {      // w1 = sum(q*q), same association as the rows' norm2 (lane-strided fma, shuffle tree)
    float acc = 0.f;
    for (int c = lane; c < (D >> 2); c += 32) {
      const float4 v = sq[wib * (D >> 2) + c];
      acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) q_n2[wib] = acc;
  }
  __syncthreads();
  const int n_slots = counters[0];
  const int C = g.cache;
  const int warps = gridDim.x * (kThreads / 32);
  for (int slot = blockIdx.x * (kThreads / 32) + wib; slot < n_slots; slot += warps) {
    const int cell = slot_cell[slot];
    float best[kQ];
#pragma unroll
    for (int k = 0; k < kQ; ++k) best[k] = -INFINITY;
    // candidate filter (memory_2.py:624-640)
    int row, col, h;
    if (g.levels == 1) { row = cell / g.gs; col = cell - row * g.gs; h = -1; }
    else { h = cell % g.levels; const int rc = cell / g.levels; row = rc / g.gs; col = rc - row * g.gs; }
    bool cand = true;
    for (int j = 0; j < C; ++j) {
      const long long r = (long long)slot * C + j;
      // the row's 128-bit loads go out before its metadata is looked at (the row always exists in the buffer; an empty or
      // filtered row costs one wasted read): the scan is bound by loads in flight, not by bytes
      float4 v[NV];
      const float4* fr = reinterpret_cast<const float4*>(feat + r * D);
#pragma unroll
      for (int c = 0; c < NV; ++c) v[c] = ld_stream_f4(fr + lane + 32 * c);
      const unsigned long long rk = row_key[r];
      const float n2 = norm2[r];
      if (g.levels == 1) h = tok_h[r];            // BEV keys: the token's own height takes part in the filter
      if (rk == BSCNAV_KEY_EMPTY) continue;
      const bool ok = filter_pass(flt, row, col, h);
      if (!ok) { if (g.levels != 1) { cand = false; break; } else continue; }
#pragma unroll
      for (int k = 0; k < kQ; ++k) {
        if (k < nq) {
          float acc = 0.f;
#pragma unroll
          for (int c = 0; c < NV; ++c) {
            const float4 qq = sq[k * (D >> 2) + lane + 32 * c];
            acc = fmaf(v[c].x, qq.x, acc); acc = fmaf(v[c].y, qq.y, acc);
            acc = fmaf(v[c].z, qq.z, acc); acc = fmaf(v[c].w, qq.w, acc);
          }
          acc = warp_sum(acc);
          // F.cosine_similarity (memory_2.py:656): w12 / sqrt(clamp(w1*w2, eps^2)), eps = 1e-8
          const float s = acc / sqrtf(fmaxf(q_n2[k] * n2, 1e-16f));
          best[k] = fmaxf(best[k], s);
        }
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < kQ; ++k)
        if (k < nq) {
          unsigned long long key = BSCNAV_KEY_EMPTY;
          if (cand && best[k] > -INFINITY) key = ((unsigned long long)(~orderable(best[k])) << 32) | (uint32_t)cell;
          keys[(long long)k * key_stride + slot] = key;
        }
    }
  }
}

// Q2: per-chunk top-K by bitonic sort in shared memory; applied repeatedly until one chunk remains.
static constexpr int kChunk = kTopkChunk;
__global__ void __launch_bounds__(1024)
k_topk_chunks(const unsigned long long* __restrict__ in, long long in_stride, const int32_t* __restrict__ n_dev, int n_host,
              unsigned long long* __restrict__ out, long long out_stride, int k, const int32_t* __restrict__ run_flag) {
  __shared__ unsigned long long s[kChunk];
  pdl_wait();                           // chained behind the previous stage with programmatic dependent launch
  pdl_launch_dependents();
  const int n = n_dev ? *n_dev : n_host;
  const int qi = blockIdx.y;
  if (run_flag && !run_flag[qi]) return;        // fall-back sorter: only for queries the histogram select handed over
  const long long base = (long long)blockIdx.x * kChunk;
  for (int i = threadIdx.x; i < kChunk; i += 1024) {
    const long long idx = base + i;
    s[i] = idx < n ? in[(long long)qi * in_stride + idx] : BSCNAV_KEY_EMPTY;
  }
  __syncthreads();
  for (int size = 2; size <= kChunk; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < kChunk / 2; t += 1024) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long a = s[lo], b = s[hi];
        if ((a > b) == up) { s[lo] = b; s[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += 1024) out[(long long)qi * out_stride + (long long)blockIdx.x * k + i] = s[i];
}

// Q3: decode the K winners: similarity from the key, (row, col, h) from the cell; BEV keys report the height
// of the best-scoring token of the cell (recomputed here, K*C dot products).
__global__ void __launch_bounds__(kThreads)
k_topk_decode(const bscnav_geom_t g, const unsigned long long* __restrict__ keys, long long key_stride, int k,
              const int32_t* __restrict__ occupied, const unsigned long long* __restrict__ row_key,
              const int32_t* __restrict__ tok_h, const float* __restrict__ feat, const float* __restrict__ norm2,
              const float* __restrict__ q, int q0, float* __restrict__ out_sim, int32_t* __restrict__ out_pos,
              int32_t* __restrict__ out_count) {
  pdl_wait();
  pdl_launch_dependents();
  const int qi = blockIdx.x;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int D = g.dim, C = g.cache;
  const int i = blockIdx.y * (kThreads / 32) + wib;        // one warp per winner
  if (i >= k) return;
  const unsigned long long key = keys[(long long)qi * key_stride + i];
  float sim = -INFINITY;
  int row = -1, col = -1, h = -1;
  if (key != BSCNAV_KEY_EMPTY) {
    sim = from_orderable(~(uint32_t)(key >> 32));
    const int cell = (int)(uint32_t)key;
    if (g.levels == 1) {
      row = cell / g.gs; col = cell - row * g.gs;
      const int slot = occupied[cell];
      if (C == 1) {
        h = tok_h[slot];                          // one token per cell: its height is the answer, nothing to re-score
      } else {
      float best = -INFINITY;
      const float* qv = q + (long long)(q0 + qi) * D;
      float qn = 0.f;
      for (int c = lane; c < D; c += 32) qn = fmaf(qv[c], qv[c], qn);
      qn = warp_sum(qn);
      for (int j = 0; j < C; ++j) {
        const long long r = (long long)slot * C + j;
        if (row_key[r] == BSCNAV_KEY_EMPTY) continue;
        float acc = 0.f;
        for (int c = lane; c < D; c += 32) acc = fmaf(feat[r * D + c], qv[c], acc);
        acc = warp_sum(acc);
        const float s = acc / sqrtf(fmaxf(qn * norm2[r], 1e-16f));
        if (s > best) { best = s; h = tok_h[r]; }
      }
      }
    } else {
      h = cell % g.levels; const int rc = cell / g.levels; row = rc / g.gs; col = rc - row * g.gs;
    }
    if (lane == 0) atomicAdd(&out_count[qi], 1);
  }
  if (lane == 0) {
    out_sim[(long long)qi * k + i] = sim;
    int32_t* p = out_pos + ((long long)qi * k + i) * 3;
    p[0] = row; p[1] = col; p[2] = h;
  }
}

int mem_num_sms();
static int num_sms() { return mem_num_sms(); }

// queries flagged by the histogram select take the chunked sorter's result
__global__ void k_pick_fallback(const int32_t* __restrict__ overflow, const unsigned long long* __restrict__ fallback,
                                long long fb_stride, unsigned long long* __restrict__ sorted, int k) {
  pdl_wait();
  pdl_launch_dependents();
  const int qi = blockIdx.x;
  if (!overflow[qi]) return;
  for (int i = threadIdx.x; i < k; i += blockDim.x) sorted[(long long)qi * k + i] = fallback[(long long)qi * fb_stride + i];
}
int mem_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int check_geom(const bscnav_geom_t* g) { return mem_check_geom(g); }
int mem_check_geom(const bscnav_geom_t* g) {
  if (!g) return BSCNAV_EINVAL;
  if (g->height <= 0 || g->width <= 0 || g->gs <= 0 || g->levels <= 0 || g->cache <= 0 || g->cache > 64) return BSCNAV_EINVAL;
  if (g->dim <= 0 || g->dim % 128 != 0 || g->dim > 2048) return BSCNAV_EINVAL;
  if (g->levels != 1 && g->levels != g->maxh - g->minh) return BSCNAV_EINVAL;
  if (g->maxh - g->minh > (1 << (32 - kSrcBits))) return BSCNAV_EINVAL;
  if ((long long)g->gs * g->gs * g->levels >= (1ll << 31)) return BSCNAV_EINVAL;
  if (g->max_slots <= 0) return BSCNAV_EINVAL;
  return BSCNAV_OK;
}

struct FuseScratch {
  uint4* rec;
  int* tile_count;
  int* tile_off;
  int* hdr;            // [0] first new slot, [1] capacity ok
  uint4* cand;         // offered keys {slot, aux, key}: one 256-entry segment per k_select block
  int* cand_n;         // entries used in each segment
  uint32_t* fresh;     // 1 bit per point: heads a run on a cell without a slot (k_index_points); padded to whole tiles
};
static int64_t scratch_layout(long long total, FuseScratch* s, char* base) {
  const long long n_tiles = (total + kTile - 1) / kTile;
  int64_t off = 0;
  if (s) s->rec = reinterpret_cast<uint4*>(base + off);
  off += align_up(total * 16, 256);
  if (s) s->tile_count = reinterpret_cast<int*>(base + off);
  off += align_up(n_tiles * 4, 256);
  if (s) s->tile_off = reinterpret_cast<int*>(base + off);
  off += align_up(n_tiles * 4, 256);
  if (s) s->hdr = reinterpret_cast<int*>(base + off);
  off += 256;
  const long long n_blocks = (total + kThreads - 1) / kThreads;
  if (s) s->cand = reinterpret_cast<uint4*>(base + off);
  off += align_up(n_blocks * kThreads * 16, 256);
  if (s) s->cand_n = reinterpret_cast<int*>(base + off);
  off += align_up(n_blocks * 4, 256);
  if (s) s->fresh = reinterpret_cast<uint32_t*>(base + off);
  off += align_up(n_tiles * (kTile / 32) * 4, 256);
  return off;
}

// shared tail of fuse / merge: slots, winners, rows.  fresh: k_index_points' mask, or null (every record is examined)
static int run_store_update(const bscnav_geom_t* g, const bscnav_store_t* st, const FuseScratch& s, long long total,
                            const float* tokens, long long tok_stride, const uint32_t* fresh, cudaStream_t stream) {
  const int n_tiles = (int)((total + kTile - 1) / kTile);
  const int n_blocks = (int)((total + kThreads - 1) / kThreads);
  const int tile_blocks = (n_tiles + kThreads / 32 - 1) / (kThreads / 32);
  if (fresh)
    BSC_LAUNCH(k_count_new_masked, tile_blocks, kThreads, 0, stream, s.rec, total, st->occupied_ids, st->first_touch, s.tile_count,
               fresh, n_tiles);
  else
    BSC_LAUNCH(k_count_new, n_tiles, kThreads, 0, stream, s.rec, total, st->occupied_ids, st->first_touch, s.tile_count, fresh);
  BSC_LAUNCH(k_scan_tiles, 1, kScanThreads, 0, stream, s.tile_count, s.tile_off, n_tiles, s.hdr, st->counters, g->max_slots);
  if (fresh)
    BSC_LAUNCH(k_assign_slots_masked, tile_blocks, kThreads, 0, stream, s.rec, total, st->occupied_ids, st->first_touch, s.tile_off,
               s.hdr, st->slot_cell, fresh, s.tile_count, n_tiles);
  else
    BSC_LAUNCH(k_assign_slots, n_tiles, kThreads, 0, stream, s.rec, total, st->occupied_ids, st->first_touch, s.tile_off,
               s.hdr, st->slot_cell, fresh, s.tile_count);
  BSC_LAUNCH(k_select, n_blocks, kThreads, 0, stream, s.rec, total, st->occupied_ids, st->first_touch,
             reinterpret_cast<unsigned long long*>(st->slot_keys), g->cache, s.cand, s.cand_n, fresh);
  const int want = (n_blocks + kThreads / 32 - 1) / (kThreads / 32), cap = num_sms() * 16;
  BSC_LAUNCH(k_commit, want < cap ? want : cap, kThreads, 0, stream, s.cand, s.cand_n, n_blocks,
             reinterpret_cast<const unsigned long long*>(st->slot_keys), reinterpret_cast<unsigned long long*>(st->row_key),
             st->tok_h, st->norm2, st->feat, tokens, tok_stride, g->cache, g->dim, st->counters);
  return BSCNAV_OK;
}

int launch_topk_chunks(const unsigned long long* in, long long in_stride, const int32_t* n_dev, int n_host, int n_chunks,
                       int n_query, unsigned long long* out, long long out_stride, int k, cudaStream_t stream) {
  BSC_LAUNCH(k_topk_chunks, dim3((unsigned)n_chunks, (unsigned)n_query), 1024, 0, stream, in, in_stride, n_dev, n_host, out,
             out_stride, k, (const int32_t*)nullptr);
  return BSCNAV_OK;
}

int launch_topk_decode(const bscnav_geom_t* g, const bscnav_store_t* st, const unsigned long long* keys, long long key_stride, int k,
                       const float* q, int n_query, float* out_sim, int32_t* out_pos, int32_t* out_count, cudaStream_t stream) {
  BSC_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int32_t) * n_query, stream));
  BSC_LAUNCH(k_topk_decode, dim3((unsigned)n_query, (unsigned)((k + kThreads / 32 - 1) / (kThreads / 32))), kThreads, 0, stream, *g,
             keys, key_stride, k, st->occupied_ids, reinterpret_cast<const unsigned long long*>(st->row_key), st->tok_h, st->feat,
             st->norm2, q, 0, out_sim, out_pos, out_count);
  return BSCNAV_OK;
}

}  // namespace bscnav

using namespace bscnav;

extern "C" {

int bscnav_abi_version(void) { return 3; }
const char* bscnav_last_error(void) { return g_last_error; }
int64_t bscnav_launch_count(void) { return g_launches.load(); }

int bscnav_store_reset(const bscnav_geom_t* g, const bscnav_store_t* st, void* stream_) {
  if (int e = check_geom(g)) return e;
  if (!st || !st->occupied_ids || !st->first_touch || !st->slot_keys || !st->row_key || !st->counters) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long cells = (long long)g->gs * g->gs * g->levels;
  const long long rows = (long long)g->max_slots * g->cache;
  const int grid = num_sms() * 8;
  BSC_LAUNCH(k_fill_u32, grid, 256, 0, stream, reinterpret_cast<uint32_t*>(st->occupied_ids), cells, 0xFFFFFFFFu);
  BSC_LAUNCH(k_fill_u32, grid, 256, 0, stream, st->first_touch, cells, kNoTouch);
  BSC_LAUNCH(k_fill_u64, grid, 256, 0, stream, reinterpret_cast<unsigned long long*>(st->slot_keys), rows, BSCNAV_KEY_EMPTY);
  BSC_LAUNCH(k_fill_u64, grid, 256, 0, stream, reinterpret_cast<unsigned long long*>(st->row_key), rows, BSCNAV_KEY_EMPTY);
  BSC_CUDA(cudaMemsetAsync(st->counters, 0, 8 * sizeof(int32_t), stream));
  return BSCNAV_OK;
}

// counters[0..3] -> four int32 in page-locked host memory, written by a kernel through the mapped address.  A 16-byte
// cudaMemcpyAsync on the scatter's stream took a copy engine: behind a 133 MB host->device copy of the next batches' frames it held
// the scatter stream back by milliseconds (0.55 ms per 54-frame step end to end).
__global__ void k_publish_counts(const int32_t* __restrict__ counters, volatile int32_t* host_dst) {
  if (threadIdx.x < 4) host_dst[threadIdx.x] = counters[threadIdx.x];
  __threadfence_system();
}
int bscnav_store_publish_counts(const bscnav_store_t* st, int32_t* host_dst, void* stream_) {
  if (!st || !st->counters || !host_dst) return BSCNAV_EINVAL;
  void* dptr = nullptr;
  if (cudaHostGetDevicePointer(&dptr, host_dst, 0) != cudaSuccess || !dptr) {
    cudaGetLastError();
    snprintf(g_last_error, sizeof(g_last_error), "publish_counts: destination is not page-locked, device-mapped host memory");
    return BSCNAV_EINVAL;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  BSC_LAUNCH(k_publish_counts, 1, 32, 0, stream, (const int32_t*)st->counters, static_cast<volatile int32_t*>(dptr));
  return BSCNAV_OK;
}

int bscnav_store_row_norms(const bscnav_geom_t* g, const bscnav_store_t* st, int64_t n_rows, void* stream_) {
  if (int e = check_geom(g)) return e;
  if (!st || !st->feat || !st->norm2 || n_rows < 0 || n_rows > (int64_t)g->max_slots * g->cache) return BSCNAV_EINVAL;
  if (n_rows == 0) return BSCNAV_OK;
  const long long blocks = (n_rows + kThreads / 32 - 1) / (kThreads / 32);
  const int grid = (int)(blocks < (long long)num_sms() * 8 ? blocks : (long long)num_sms() * 8);
  BSC_LAUNCH(k_row_norms, grid, kThreads, 0, (cudaStream_t)stream_, st->feat, st->norm2, (long long)n_rows, g->dim);
  return BSCNAV_OK;
}

int64_t bscnav_fuse_scratch_bytes(int32_t n_frames, int32_t n_points) {
  if (n_frames <= 0 || n_points <= 0) return BSCNAV_EINVAL;
  return scratch_layout((long long)n_frames * n_points, nullptr, nullptr);
}

int bscnav_fuse_frames(const bscnav_geom_t* g, const bscnav_store_t* st, const float* depth, const double* poses,
                       const uint32_t* ord_base, const int32_t* sample_idx, int32_t n_frames, int32_t n_points,
                       const float* tokens, void* scratch, int64_t scratch_bytes, void* records_out, void* stream_) {
  if (int e = check_geom(g)) return e;
  if (!st || !depth || !poses || !ord_base || !tokens || !scratch) return BSCNAV_EINVAL;
  if (n_frames <= 0 || n_points <= 0) return BSCNAV_EINVAL;
  if (!sample_idx && n_points != g->height * g->width) return BSCNAV_EINVAL;
  const long long total = (long long)n_frames * n_points;
  if (total >= (1ll << 31)) return BSCNAV_EINVAL;
  if ((long long)n_frames * g->n_patch_h * g->n_patch_w > (1ll << kSrcBits)) return BSCNAV_EINVAL;
  FuseScratch s;
  if (scratch_layout(total, &s, static_cast<char*>(scratch)) > scratch_bytes) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int n_blocks = (int)((total + kThreads - 1) / kThreads);
  // mask words of the last, partial tile whose warps do not exist stay zero
  const long long n_tiles = (total + kTile - 1) / kTile;
  BSC_CUDA(cudaMemsetAsync(s.fresh + (n_tiles - 1) * (kTile / 32), 0, (kTile / 32) * 4, stream));
  BSC_LAUNCH(k_index_points, n_blocks, kThreads, 0, stream, *g, 1.0 / g->cs, depth, poses, ord_base, sample_idx, n_points, total,
             s.rec, st->occupied_ids, st->first_touch, s.fresh);
  if (records_out) BSC_CUDA(cudaMemcpyAsync(records_out, s.rec, total * 16, cudaMemcpyDeviceToDevice, stream));
  return run_store_update(g, st, s, total, tokens, g->dim, s.fresh, stream);
}

int bscnav_merge_delta(const bscnav_geom_t* g, const bscnav_store_t* st, const int32_t* d_cell, const uint64_t* d_key,
                       const int32_t* d_h, const float* d_feat, int32_t n, void* scratch, int64_t scratch_bytes,
                       void* stream_) {
  if (int e = check_geom(g)) return e;
  if (!st || !d_cell || !d_key || !d_h || !d_feat || !scratch || n < 0) return BSCNAV_EINVAL;
  if (n == 0) return BSCNAV_OK;
  if (n > (1 << kSrcBits)) return BSCNAV_EINVAL;
  FuseScratch s;
  if (scratch_layout(n, &s, static_cast<char*>(scratch)) > scratch_bytes) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  BSC_LAUNCH(k_delta_records, (n + kThreads - 1) / kThreads, kThreads, 0, stream, d_cell,
             reinterpret_cast<const unsigned long long*>(d_key), d_h, n, s.rec, st->occupied_ids, st->first_touch);
  return run_store_update(g, st, s, n, d_feat, g->dim, nullptr, stream);
}

int64_t bscnav_export_scratch_bytes(const bscnav_geom_t* g) {
  if (check_geom(g)) return BSCNAV_EINVAL;
  const long long max_rows = (long long)g->max_slots * g->cache;
  return 2 * align_up(((max_rows + kTile - 1) / kTile) * 4, 256);
}

int bscnav_export_delta(const bscnav_geom_t* g, const bscnav_store_t* st, uint32_t ord_lo, uint32_t ord_hi,
                        int32_t* d_cell, uint64_t* d_key, int32_t* d_h, float* d_feat, int32_t max_out,
                        int32_t* out_n, void* scratch, int64_t scratch_bytes, void* stream_) {
  if (int e = check_geom(g)) return e;
  if (!st || !d_cell || !d_key || !d_h || !d_feat || !out_n || !scratch || max_out <= 0) return BSCNAV_EINVAL;
  if (bscnav_export_scratch_bytes(g) > scratch_bytes) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long max_rows = (long long)g->max_slots * g->cache;
  const int n_tiles = (int)((max_rows + kTile - 1) / kTile);
  int* tile_count = static_cast<int*>(scratch);
  int* tile_off = reinterpret_cast<int*>(static_cast<char*>(scratch) + align_up((long long)n_tiles * 4, 256));
  const unsigned long long* rk = reinterpret_cast<const unsigned long long*>(st->row_key);
  BSC_LAUNCH(k_export_count, n_tiles, kThreads, 0, stream, rk, st->counters, g->cache, ord_lo, ord_hi, tile_count);
  BSC_LAUNCH(k_export_scan, 1, kThreads, 0, stream, tile_count, tile_off, n_tiles, out_n, max_out);
  BSC_LAUNCH(k_export_write, n_tiles, kThreads, 0, stream, rk, st->counters, st->slot_cell, st->tok_h, st->feat, g->cache,
             g->dim, ord_lo, ord_hi, tile_off, d_cell, reinterpret_cast<unsigned long long*>(d_key), d_h, d_feat, max_out);
  return BSCNAV_OK;
}

int64_t bscnav_packed_row_bytes(const bscnav_geom_t* g) {
  if (check_geom(g)) return BSCNAV_EINVAL;
  return 16 + 4ll * g->dim;
}

int bscnav_export_packed(const bscnav_geom_t* g, const bscnav_store_t* st, uint32_t ord_lo, uint32_t ord_hi, void* packed,
                         int32_t max_out, int32_t* out_n, void* scratch, int64_t scratch_bytes, void* stream_) {
  if (int e = check_geom(g)) return e;
  if (!st || !packed || !out_n || !scratch || max_out <= 0 || (reinterpret_cast<uintptr_t>(packed) & 15)) return BSCNAV_EINVAL;
  if (bscnav_export_scratch_bytes(g) > scratch_bytes) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long max_rows = (long long)g->max_slots * g->cache;
  const int n_tiles = (int)((max_rows + kTile - 1) / kTile);
  int* tile_count = static_cast<int*>(scratch);
  int* tile_off = reinterpret_cast<int*>(static_cast<char*>(scratch) + align_up((long long)n_tiles * 4, 256));
  const unsigned long long* rk = reinterpret_cast<const unsigned long long*>(st->row_key);
  BSC_LAUNCH(k_export_count, n_tiles, kThreads, 0, stream, rk, st->counters, g->cache, ord_lo, ord_hi, tile_count);
  BSC_LAUNCH(k_export_scan, 1, kThreads, 0, stream, tile_count, tile_off, n_tiles, out_n, max_out);
  BSC_LAUNCH(k_export_packed, n_tiles, kThreads, 0, stream, rk, st->counters, st->slot_cell, st->tok_h, st->feat, g->cache,
             g->dim, ord_lo, ord_hi, tile_off, static_cast<char*>(packed), max_out);
  return BSCNAV_OK;
}

int64_t bscnav_merge_packed_scratch_bytes(int32_t n_segments, int32_t seg_capacity) {
  if (n_segments <= 0 || seg_capacity <= 0) return BSCNAV_EINVAL;
  const long long total = (long long)n_segments * seg_capacity;
  return scratch_layout(total < (1ll << kSrcBits) ? total : (1ll << kSrcBits), nullptr, nullptr);
}

int bscnav_merge_packed(const bscnav_geom_t* g, const bscnav_store_t* st, const void* packed, int32_t n_segments,
                        int32_t seg_capacity, const int32_t* seg_counts, int32_t count_stride, int32_t skip_segment,
                        void* scratch, int64_t scratch_bytes, void* stream_) {
  if (int e = check_geom(g)) return e;
  if (!st || !packed || !seg_counts || !scratch || n_segments <= 0 || seg_capacity <= 0 || count_stride <= 0) return BSCNAV_EINVAL;
  if ((reinterpret_cast<uintptr_t>(packed) & 15)) return BSCNAV_EINVAL;
  if (bscnav_merge_packed_scratch_bytes(n_segments, seg_capacity) > scratch_bytes) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long total = (long long)n_segments * seg_capacity, row_bytes = 16 + 4ll * g->dim;
  const long long chunk = 1ll << kSrcBits;                    // the record's source-row field
  for (long long first = 0; first < total; first += chunk) {
    const int n = (int)(total - first < chunk ? total - first : chunk);
    FuseScratch s;
    scratch_layout(n, &s, static_cast<char*>(scratch));
    BSC_LAUNCH(k_packed_records, (n + kThreads - 1) / kThreads, kThreads, 0, stream, static_cast<const char*>(packed), row_bytes,
               first, n, seg_capacity, seg_counts, count_stride, skip_segment, s.rec, st->occupied_ids, st->first_touch);
    const float* rows = reinterpret_cast<const float*>(static_cast<const char*>(packed) + first * row_bytes + 16);
    if (int e = run_store_update(g, st, s, n, rows, row_bytes / 4, nullptr, stream)) return e;
  }
  return BSCNAV_OK;
}

int64_t bscnav_query_scratch_bytes(const bscnav_geom_t* g, int32_t n_query, int32_t k) {
  if (check_geom(g) || n_query <= 0 || k <= 0 || k > 1024) return BSCNAV_EINVAL;
  // level-0 keys: kQ * max_slots; then ping-pong candidate buffers
  const long long lvl0 = (long long)kQ * g->max_slots;
  const long long chunks = (g->max_slots + kChunk - 1) / kChunk;
  const long long cand = (long long)kQ * chunks * k;
  return align_up(lvl0 * 8, 256) + 2 * align_up(cand * 8, 256) + 256 + exact_select_scratch_bytes(kQ) + align_up((long long)kQ * k * 8, 256);
}

int bscnav_query_topk(const bscnav_geom_t* g, const bscnav_store_t* st, const float* q, int32_t n_query, int32_t k,
                      const bscnav_filter_t* filter, void* scratch, int64_t scratch_bytes, float* out_sim,
                      int32_t* out_pos, int32_t* out_count, void* stream_) {
  if (int e = check_geom(g)) return e;
  if (!st || !q || !scratch || !out_sim || !out_pos || !out_count || n_query <= 0 || k <= 0 || k > 1024) return BSCNAV_EINVAL;
  if (bscnav_query_scratch_bytes(g, n_query, k) > scratch_bytes) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  bscnav_filter_t flt;
  if (filter) flt = *filter; else { flt.radius2 = -1; flt.cur[0] = flt.cur[1] = flt.cur[2] = 0; flt.floor_min = 1; flt.floor_max = 0; flt.reserved = 0; }
  const long long lvl0 = (long long)kQ * g->max_slots;
  const long long chunks0 = (g->max_slots + kChunk - 1) / kChunk;
  const long long cand = (long long)kQ * chunks0 * k;
  char* base = static_cast<char*>(scratch);
  unsigned long long* keys0 = reinterpret_cast<unsigned long long*>(base);
  unsigned long long* bufA = reinterpret_cast<unsigned long long*>(base + align_up(lvl0 * 8, 256));
  unsigned long long* bufB = reinterpret_cast<unsigned long long*>(base + align_up(lvl0 * 8, 256) + align_up(cand * 8, 256));
  const int D = g->dim;
  const int grid = num_sms() * 3;      // 3 resident blocks of 256 threads per SM at ~76 registers: one full wave, grid-stride
  for (int q0 = 0; q0 < n_query; q0 += kQ) {
    const int nq = (n_query - q0) < kQ ? (n_query - q0) : kQ;
    const size_t smem = (size_t)nq * D * sizeof(float);
    char* sel_base = base + align_up(lvl0 * 8, 256) + 2 * align_up(cand * 8, 256) + 256;
#define BSC_QUERY_CASE(NV)                                                                                            \
  case NV: {                                                                                                          \
    if (smem > 48 * 1024)                                                                                             \
      BSC_CUDA(cudaFuncSetAttribute(k_query_slots<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    BSC_LAUNCH(k_query_slots<NV>, grid, kThreads, smem, stream, *g, st->feat, st->norm2,                             \
               reinterpret_cast<const unsigned long long*>(st->row_key), st->slot_cell, st->tok_h, st->counters, q,  \
               q0, nq, flt, keys0, (long long)g->max_slots, reinterpret_cast<uint32_t*>(sel_base),                   \
               (int)(exact_select_zero_bytes(kQ) / 4));                                                              \
  } break;
    switch (D / 128) {
      BSC_QUERY_CASE(1) BSC_QUERY_CASE(2) BSC_QUERY_CASE(3) BSC_QUERY_CASE(4) BSC_QUERY_CASE(5) BSC_QUERY_CASE(6)
      BSC_QUERY_CASE(7) BSC_QUERY_CASE(8) BSC_QUERY_CASE(10) BSC_QUERY_CASE(12) BSC_QUERY_CASE(16)
      default: return BSCNAV_EINVAL;
    }
#undef BSC_QUERY_CASE
    // top-K: histogram select over the exact keys + one small sort (three launches of a few microseconds); a query with more
    // than 4096 keys in its k-th bin (massive ties) is handed to the chunked bitonic sorter below, which otherwise exits at once
    char* sel = base + align_up(lvl0 * 8, 256) + 2 * align_up(cand * 8, 256) + 256;
    unsigned long long* sorted = reinterpret_cast<unsigned long long*>(sel + exact_select_scratch_bytes(kQ));
    int32_t* overflow = nullptr;
    if (int e = launch_exact_select(keys0, (long long)g->max_slots, st->counters, g->max_slots, nq, k, sel, sorted, &overflow, stream,
                                    /*zeroed=*/true, out_count + q0)) return e;
    // level 0: n is the device-side slot count, but the grid covers the capacity
    long long n_chunks = chunks0;
    BSC_LAUNCH_PDL(k_topk_chunks, dim3((unsigned)n_chunks, (unsigned)nq), 1024, 0, stream, keys0, (long long)g->max_slots,
                   (const int32_t*)st->counters, 0, bufA, (long long)(n_chunks * k), k, (const int32_t*)overflow);
    unsigned long long* in = bufA;
    unsigned long long* out = bufB;
    long long n_in = n_chunks * k, in_stride = n_chunks * k;
    while (n_in > k) {
      n_chunks = (n_in + kChunk - 1) / kChunk;
      BSC_LAUNCH_PDL(k_topk_chunks, dim3((unsigned)n_chunks, (unsigned)nq), 1024, 0, stream, (const unsigned long long*)in, in_stride,
                     (const int32_t*)nullptr, (int)n_in, out, (long long)(n_chunks * k), k, (const int32_t*)overflow);
      in_stride = n_chunks * k;
      n_in = n_chunks * k;
      unsigned long long* t = in; in = out; out = t;
      if (n_chunks == 1) break;
    }
    BSC_LAUNCH_PDL(k_pick_fallback, (unsigned)nq, 128, 0, stream, (const int32_t*)overflow, (const unsigned long long*)in, in_stride, sorted, k);
    in = sorted; in_stride = k;
    // out_count was cleared by k_sort_candidates (launch_exact_select): no memset between the kernels of the chain
    BSC_LAUNCH_PDL(k_topk_decode, dim3((unsigned)nq, (unsigned)((k + kThreads / 32 - 1) / (kThreads / 32))), kThreads, 0, stream, *g,
                   (const unsigned long long*)in, in_stride, k, (const int32_t*)st->occupied_ids,
                   reinterpret_cast<const unsigned long long*>(st->row_key), (const int32_t*)st->tok_h, (const float*)st->feat,
                   (const float*)st->norm2, q, q0, out_sim + (long long)q0 * k, out_pos + (long long)q0 * k * 3, out_count + q0);
  }
  return BSCNAV_OK;
}

}  // extern "C"
