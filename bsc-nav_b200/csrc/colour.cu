// Colour / occupancy side channels of the memory (memory_2.py:869-870, 873-875, 888-903): per voxel the alpha-weighted colour
// mean `grid_rgb` and its accumulated `weight`, and the BEV map `cv_map` / `max_height` (colour of the highest point of each
// (row, col), last writer wins among equals).  Optional pass, run after bscnav_fuse_frames on the same batch (it needs the
// batch's slots); voxel keys only.
//
// The reference's colour mean is cast to uint8 after EVERY update (memory_2.py:896-898), so a voxel's colour is a strictly
// sequential, non-associative chain over its points in visiting order.  To reproduce it bit for bit the batch's points are
// brought into (slot, visiting order) with a stable radix sort by slot (cub::DeviceRadixSort -- library code, like the NCCL
// collectives: this is a side channel, not the hot path) and one thread walks each voxel's run with the reference's exact
// numpy promotions:  u8 * f32 -> f32;  + u8 * f64 -> f64;  / (f32 + f64) -> f64;  -> uint8 by truncation;  weight = f32(f64(w) + alpha).
// The BEV map is a true scatter-max: 64-bit atomicMax over (h, ordinal) followed by a commit pass.
// Cost: one more pass over the depth/rgb frames plus the sort; with the reference's sampling (depth_sample_rate = 1000) it is
// negligible, for dense frames the longest per-voxel chain bounds it.
#include <cub/device/device_radix_sort.cuh>

#include "memory_kernels.cuh"

namespace bscnav {

static constexpr int kColThreads = 256;

__device__ __forceinline__ double c_dot3(const double* m, double a, double b, double c) {
  return __fma_rn(m[2], c, __fma_rn(m[1], b, __dmul_rn(m[0], a)));
}
__device__ __forceinline__ double c_dot4(const double* m, double a, double b, double c, double d) {
  return __fma_rn(m[3], d, __fma_rn(m[2], c, __fma_rn(m[1], b, __dmul_rn(m[0], a))));
}
__device__ __forceinline__ int c_trunc(double v) {
  if (!(v > -2.0e9)) return -2000000000;
  if (!(v < 2.0e9)) return 2000000000;
  return (int)v;
}

// per visited point: slot (sort key), packed rgb of the pixel the reference looks up, alpha, BEV key; same float64 arithmetic
// as k_index_points (utils.py:153-214) plus the image-intrinsics projection (memory_2.py:869-870)
__global__ void __launch_bounds__(kColThreads)
k_colour_points(const bscnav_geom_t g, const float* __restrict__ depth, const uint8_t* __restrict__ rgb, int ch,
                const double* __restrict__ poses, const uint32_t* __restrict__ ord_base, const int32_t* __restrict__ sample_idx,
                int n_points, long long total, const int32_t* __restrict__ occupied, uint32_t* __restrict__ sort_key,
                uint32_t* __restrict__ sort_val, uint32_t* __restrict__ pt_rgb, double* __restrict__ pt_alpha,
                unsigned long long* __restrict__ pt_cvkey, int32_t* __restrict__ pt_rc, unsigned long long* __restrict__ cv_key) {
  const long long i = (long long)blockIdx.x * kColThreads + threadIdx.x;
  if (i >= total) return;
  uint32_t key = (uint32_t)g.max_slots;          // sentinel: sorts behind every slot
  int rc = -1;
  unsigned long long ck = 0;
  const int b = (int)(i / n_points);
  const int j = (int)(i - (long long)b * n_points);
  const int pix = sample_idx ? sample_idx[i] : j;
  const int hw = g.height * g.width;
  const uint32_t ord = ord_base[b] + (uint32_t)j;
  const double z = (double)depth[(long long)b * hw + pix];
  const int v = pix / g.width, u = pix - v * g.width;
  const double x = (double)u + 0.5, y = (double)v + 0.5;
  const double p0 = __dmul_rn(c_dot3(g.kinv + 0, x, y, 1.0), z);
  const double p1 = __dmul_rn(c_dot3(g.kinv + 3, x, y, 1.0), z);
  const double p2 = __dmul_rn(c_dot3(g.kinv + 6, x, y, 1.0), z);
  if (p2 > g.min_depth && p2 < g.max_depth) {
    const double* M = poses + (long long)b * 16;
    const double gx = c_dot4(M + 0, p0, p1, p2, 1.0), gy = c_dot4(M + 4, p0, p1, p2, 1.0), gz = c_dot4(M + 8, p0, p1, p2, 1.0);
    const double half = (double)g.gs / 2.0;
    const int row = c_trunc(__dsub_rn(half, (double)c_trunc(__ddiv_rn(gx, g.cs))));
    const int col = c_trunc(__dsub_rn(half, (double)c_trunc(__ddiv_rn(gy, g.cs))));
    const int h = c_trunc(__ddiv_rn(gz, g.cs));
    const bool out = col >= g.gs || row >= g.gs || h >= g.maxh || col < 0 || row < 0 || h < g.minh;
    if (!out) {
      const double q0 = c_dot3(g.kpatch + 0, p0, p1, p2), q1 = c_dot3(g.kpatch + 3, p0, p1, p2), q2 = c_dot3(g.kpatch + 6, p0, p1, p2);
      const int ppx = c_trunc(__dsub_rn(__ddiv_rn(q0, q2), 0.5));
      const int ppy = c_trunc(__dsub_rn(__ddiv_rn(q1, q2), 0.5));
      if (!(ppx < 0 || ppy < 0 || ppx >= g.n_patch_h || ppy >= g.n_patch_w)) {
        // rgb[py, px] with the image intrinsics (memory_2.py:869-870; not bounds-checked there: a negative index wraps as in
        // numpy, an index past the end -- an IndexError in the reference -- leaves the point without colour here)
        const double i0 = c_dot3(g.kimg + 0, p0, p1, p2), i1 = c_dot3(g.kimg + 3, p0, p1, p2), i2 = c_dot3(g.kimg + 6, p0, p1, p2);
        int ipx = c_trunc(__dsub_rn(__ddiv_rn(i0, i2), 0.5));
        int ipy = c_trunc(__dsub_rn(__ddiv_rn(i1, i2), 0.5));
        if (ipx < 0) ipx += g.width;
        if (ipy < 0) ipy += g.height;
        if (ipx >= 0 && ipx < g.width && ipy >= 0 && ipy < g.height) {
          const int hh = h - g.minh;
          const int cell = (row * g.gs + col) * g.levels + hh;
          const int slot = occupied[cell];
          if (slot >= 0) {
            key = (uint32_t)slot;
            const uint8_t* px = rgb + ((long long)b * hw + (long long)ipy * g.width + ipx) * ch;
            pt_rgb[i] = (uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16);
            const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(p0, p0), __dmul_rn(p1, p1)), __dmul_rn(p2, p2));
            pt_alpha[i] = exp(__ddiv_rn(-r2, 1.2));                       // memory_2.py:873-875, 2 * sigma_sq = 1.2
            rc = row * g.gs + col;
            ck = (((unsigned long long)(uint32_t)hh << 32) | ord) + 1ull;  // height >= max_height: later ordinal wins a tie
            atomicMax(&cv_key[rc], ck);
          }
        }
      }
    }
  }
  sort_key[i] = key;
  sort_val[i] = (uint32_t)i;
  pt_cvkey[i] = ck;
  pt_rc[i] = rc;
}

// one thread per run of equal slots in the sorted order: the sequential colour chain of that voxel (memory_2.py:888-899)
__global__ void __launch_bounds__(kColThreads)
k_colour_chain(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, long long total, uint32_t sentinel,
               const uint32_t* __restrict__ pt_rgb, const double* __restrict__ pt_alpha, const int32_t* __restrict__ counters,
               uint8_t* __restrict__ grid_rgb, float* __restrict__ weight) {
  const long long p0 = (long long)blockIdx.x * kColThreads + threadIdx.x;
  if (p0 >= total) return;
  const uint32_t slot = keys[p0];
  if (slot == sentinel || (p0 > 0 && keys[p0 - 1] == slot)) return;       // not the head of a run
  const bool fresh = (int)slot >= counters[2];                            // voxel created by this batch: its first point initialises
  int r = grid_rgb[3 * slot], gch = grid_rgb[3 * slot + 1], bch = grid_rgb[3 * slot + 2];
  float w = weight[slot];
  for (long long p = p0; p < total && keys[p] == slot; ++p) {
    const uint32_t i = vals[p];
    const uint32_t c = pt_rgb[i];
    const double a = pt_alpha[i];
    const int cr = c & 255, cg = (c >> 8) & 255, cb = (c >> 16) & 255;
    if (fresh && p == p0) {
      r = cr; gch = cg; bch = cb;
      w = (float)__dadd_rn((double)w, a);                                  // weight[max_id] += alpha on a zero entry
    } else {
      const double den = __dadd_rn((double)w, a);
      // (grid_rgb * weight + rgb_v * alpha) / (weight + alpha): u8 * f32 -> f32, u8 * f64 -> f64, sum and quotient in f64
      r = (int)__ddiv_rn(__dadd_rn((double)__fmul_rn((float)r, w), __dmul_rn((double)cr, a)), den);
      gch = (int)__ddiv_rn(__dadd_rn((double)__fmul_rn((float)gch, w), __dmul_rn((double)cg, a)), den);
      bch = (int)__ddiv_rn(__dadd_rn((double)__fmul_rn((float)bch, w), __dmul_rn((double)cb, a)), den);
      w = (float)den;
    }
  }
  grid_rgb[3 * slot] = (uint8_t)r; grid_rgb[3 * slot + 1] = (uint8_t)gch; grid_rgb[3 * slot + 2] = (uint8_t)bch;
  weight[slot] = w;
}

// BEV map commit: the point that holds a (row, col)'s maximum key writes its colour
__global__ void __launch_bounds__(kColThreads)
k_cv_commit(const unsigned long long* __restrict__ pt_cvkey, const int32_t* __restrict__ pt_rc, const uint32_t* __restrict__ pt_rgb,
            long long total, const unsigned long long* __restrict__ cv_key, uint8_t* __restrict__ cv_map) {
  const long long i = (long long)blockIdx.x * kColThreads + threadIdx.x;
  if (i >= total) return;
  const int rc = pt_rc[i];
  if (rc < 0 || cv_key[rc] != pt_cvkey[i]) return;
  const uint32_t c = pt_rgb[i];
  cv_map[3 * rc] = c & 255; cv_map[3 * rc + 1] = (c >> 8) & 255; cv_map[3 * rc + 2] = (c >> 16) & 255;
}

__global__ void k_colour_mark(int32_t* counters) { counters[2] = counters[0]; }

struct ColourScratch {
  uint32_t *key_in, *key_out, *val_in, *val_out, *pt_rgb;
  double* pt_alpha;
  unsigned long long* pt_cvkey;
  int32_t* pt_rc;
  void* cub_temp;
  size_t cub_bytes;
};
static int64_t colour_layout(long long total, int end_bit, ColourScratch* s, char* base) {
  long long off = 0;
  auto take = [&](long long bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes, 256); return p; };
  char* p;
  p = take(total * 4); if (s) s->key_in = reinterpret_cast<uint32_t*>(p);
  p = take(total * 4); if (s) s->key_out = reinterpret_cast<uint32_t*>(p);
  p = take(total * 4); if (s) s->val_in = reinterpret_cast<uint32_t*>(p);
  p = take(total * 4); if (s) s->val_out = reinterpret_cast<uint32_t*>(p);
  p = take(total * 4); if (s) s->pt_rgb = reinterpret_cast<uint32_t*>(p);
  p = take(total * 8); if (s) s->pt_alpha = reinterpret_cast<double*>(p);
  p = take(total * 8); if (s) s->pt_cvkey = reinterpret_cast<unsigned long long*>(p);
  p = take(total * 4); if (s) s->pt_rc = reinterpret_cast<int32_t*>(p);
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)total, 0, end_bit);
  p = take((long long)cub_bytes); if (s) { s->cub_temp = p; s->cub_bytes = cub_bytes; }
  return off;
}
static int slot_bits(int max_slots) {
  int b = 1;
  while ((1ll << b) <= (long long)max_slots) ++b;       // the sentinel max_slots itself must be representable
  return b;
}

}  // namespace bscnav

using namespace bscnav;

extern "C" {

int bscnav_colour_reset(const bscnav_geom_t* g, const bscnav_colour_t* c, void* stream_) {
  if (int e = mem_check_geom(g)) return e;
  if (!c || !c->grid_rgb || !c->weight || !c->cv_key || !c->cv_map) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  BSC_CUDA(cudaMemsetAsync(c->grid_rgb, 0, (size_t)g->max_slots * 3, stream));
  BSC_CUDA(cudaMemsetAsync(c->weight, 0, (size_t)g->max_slots * 4, stream));
  BSC_CUDA(cudaMemsetAsync(c->cv_key, 0, (size_t)g->gs * g->gs * 8, stream));
  BSC_CUDA(cudaMemsetAsync(c->cv_map, 0, (size_t)g->gs * g->gs * 3, stream));
  return BSCNAV_OK;
}

int64_t bscnav_colour_scratch_bytes(const bscnav_geom_t* g, int32_t n_frames, int32_t n_points) {
  if (mem_check_geom(g) || n_frames <= 0 || n_points <= 0) return BSCNAV_EINVAL;
  return colour_layout((long long)n_frames * n_points, slot_bits(g->max_slots), nullptr, nullptr);
}

int bscnav_colour_frames(const bscnav_geom_t* g, const bscnav_store_t* st, const bscnav_colour_t* c, const float* depth,
                         const uint8_t* rgb, int32_t channels, const double* poses, const uint32_t* ord_base,
                         const int32_t* sample_idx, int32_t n_frames, int32_t n_points, void* scratch, int64_t scratch_bytes,
                         void* stream_) {
  if (int e = mem_check_geom(g)) return e;
  if (!st || !c || !depth || !rgb || !poses || !ord_base || !scratch || n_frames <= 0 || n_points <= 0) return BSCNAV_EINVAL;
  if (channels < 3 || channels > 4) return BSCNAV_EINVAL;
  if (g->levels == 1) return BSCNAV_EUNSUPPORTED;           // the colour mean is defined per voxel (row, col, h)
  if (!sample_idx && n_points != g->height * g->width) return BSCNAV_EINVAL;
  const long long total = (long long)n_frames * n_points;
  if (total >= (1ll << 31)) return BSCNAV_EINVAL;
  const int end_bit = slot_bits(g->max_slots);
  ColourScratch s;
  if (colour_layout(total, end_bit, &s, static_cast<char*>(scratch)) > scratch_bytes) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  const unsigned blocks = (unsigned)((total + kColThreads - 1) / kColThreads);
  BSC_LAUNCH(k_colour_points, blocks, kColThreads, 0, stream, *g, depth, rgb, (int)channels, poses, ord_base, sample_idx, n_points,
             total, st->occupied_ids, s.key_in, s.val_in, s.pt_rgb, s.pt_alpha, s.pt_cvkey, s.pt_rc,
             reinterpret_cast<unsigned long long*>(c->cv_key));
  BSC_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_temp, s.cub_bytes, s.key_in, s.key_out, s.val_in, s.val_out, (int)total, 0, end_bit,
                                           stream));
  BSC_LAUNCH(k_colour_chain, blocks, kColThreads, 0, stream, s.key_out, s.val_out, total, (uint32_t)g->max_slots, s.pt_rgb, s.pt_alpha,
             st->counters, c->grid_rgb, c->weight);
  BSC_LAUNCH(k_cv_commit, blocks, kColThreads, 0, stream, s.pt_cvkey, s.pt_rc, s.pt_rgb, total,
             reinterpret_cast<const unsigned long long*>(c->cv_key), c->cv_map);
  BSC_LAUNCH(k_colour_mark, 1, 1, 0, stream, st->counters);
  return BSCNAV_OK;
}

}  // extern "C"
