/* bscnav_b200 -- C ABI of the B200-native spatial-memory build + query path.
 *
 * The reference (Heathcliff-saku/BSC-Nav) has no FFI layer: its boundary is the Python class
 * VoxelTokenMemory (memory_2.py:38).  These entry points are what a binding for that class's hot path
 * would bind; each cites the reference lines it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions: every pointer is a DEVICE pointer unless named host_*; `stream` is a cudaStream_t passed as
 * void*; calls enqueue work on that stream and return without synchronising (unless stated); the return
 * value is 0 on success or a negative BSCNAV_E* code; nothing throws; re-entrant per stream.
 * The library never allocates device memory: the caller owns every buffer (sizes given by *_bytes helpers).
 * Process model: one process per GPU (the torch.distributed layout); per-kernel attributes (shared-memory opt-in,
 * cluster occupancy) are cached per process for the device that is current at the first launch.
 */
#ifndef BSCNAV_B200_H
#define BSCNAV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSCNAV_OK            0
#define BSCNAV_EINVAL       -1   /* bad argument (null pointer, size out of range, unsupported shape) */
#define BSCNAV_ECAPACITY    -2   /* slot capacity / ordinal range exceeded                            */
#define BSCNAV_ECUDA        -3   /* a CUDA runtime/driver call failed (see bscnav_last_error)         */
#define BSCNAV_EUNSUPPORTED -4   /* device is not sm_100                                              */

#define BSCNAV_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull

int bscnav_abi_version(void);
const char* bscnav_last_error(void);           /* host string describing the last BSCNAV_ECUDA on this thread */

/* ------------------------------------------------------------------------------------------------------
 * Geometry of one memory (ctor state, memory_2.py:78-123; utils.py:144-150,181-186).
 * All matrices row-major float64, computed on the host exactly as the reference does. */
typedef struct {
  int32_t height, width;          /* frame size (args.height/width)                                   */
  int32_t gs;                     /* grid_size                                                         */
  int32_t minh, maxh;             /* int(floor_height/cs), int(map_height/cs)   memory_2.py:122-123    */
  int32_t levels;                 /* maxh-minh (voxel keys) or 1 (BEV keys: key reduced over h)        */
  int32_t n_patch_h, n_patch_w;   /* token grid of the feature extractor (16x16 for /14 at 224)       */
  int32_t cache;                  /* C: tokens kept per cell (cache_size, memory_2.py:111)             */
  int32_t dim;                    /* D: token dimension (token_dim, memory_2.py:107), multiple of 128  */
  int32_t max_slots;              /* slot capacity (reference: gs*gs, memory_2.py:715-718)             */
  int32_t reserved;
  double  cs;                     /* cell_size                                                         */
  double  min_depth, max_depth;   /* strict bounds (utils.py:174-176)                                  */
  double  kinv[9];                /* inv(calib_mat)                    utils.py:164                    */
  double  kimg[9];                /* calib_mat (image intrinsics)      memory_2.py:102                 */
  double  kpatch[9];              /* get_sim_cam_mat(n_patch)          memory_2.py:857                 */
} bscnav_geom_t;

/* Device buffers of one memory ("store").  Sizes in elements, with cells = gs*gs*levels, S = max_slots, C = cache:
 *   occupied_ids int32[cells]   cell -> slot, -1 = empty            (occupied_ids, memory_2.py:716)
 *   first_touch  uint32[cells]  scratch, must be 0xFFFFFFFF between calls
 *   slot_cell    int32[S]       slot -> linear cell                 (grid_rgb_pos, memory_2.py:715)
 *   slot_keys    uint64[S*C]    per slot the C best keys ascending; key = float32 bits of r^2 << 32 | ordinal
 *   row_key      uint64[S*C]    key of the token held by physical row slot*C+j (unordered within a slot)
 *   tok_h        int32[S*C]     h-minh of the point that produced the row (the position reported in BEV mode)
 *   norm2        float[S*C]     squared L2 norm of the row
 *   feat         float[S*C*D]   token rows                          (feat.h5df 'features', memory_2.py:335-354)
 *   counters     int32[8]       [0] = number of slots in use (max_id, memory_2.py:894), [1] = capacity-overflow flag,
 *                               [2] = slots whose colour is initialised (bscnav_colour_frames),
 *                               [3] = token rows copied into the store since the last reset (bench.py's rows_committed_per_step)
 */
typedef struct {
  int32_t*  occupied_ids;
  uint32_t* first_touch;
  int32_t*  slot_cell;
  uint64_t* slot_keys;
  uint64_t* row_key;
  int32_t*  tok_h;
  float*    norm2;
  float*    feat;
  int32_t*  counters;
} bscnav_store_t;

/* Reset a store to empty (replaces _init_cache, memory_2.py:708-722). */
int bscnav_store_reset(const bscnav_geom_t* g, const bscnav_store_t* st, void* stream);

/* Asynchronous read-out of counters[0..3] (slots in use = max_id, memory_2.py:894; overflow flag; colour mark; rows committed):
 * a one-warp kernel on `stream` writes them to host_dst, four int32 of page-locked (device-mapped) host memory; the caller
 * waits on an event recorded behind it.  No copy-engine operation enters the stream (a 16-byte cudaMemcpyAsync queued behind
 * the build's large host->device copies and stalled the scatter stream). */
int bscnav_store_publish_counts(const bscnav_store_t* st, int32_t* host_dst, void* stream);

/* Recompute norm2[0, n_rows) from feat (after filling token rows from a memory directory, memory_2.py:185-200);
 * same summation order as the commit kernel, so a reloaded store scores bit-identically. */
int bscnav_store_row_norms(const bscnav_geom_t* g, const bscnav_store_t* st, int64_t n_rows, void* stream);

/* Scratch needed by bscnav_fuse_frames for a batch of n_frames x n_points points. */
int64_t bscnav_fuse_scratch_bytes(int32_t n_frames, int32_t n_points);

/* Fuse a batch of frames into the store: depth unprojection, pose transform, voxel index, patch projection,
 * deterministic slot assignment and top-C winner selection, feature-row commit.
 * Replaces the per-point loop of obs2voxeltoken (memory_2.py:859-894) + update_memory_dist_base (:326-358).
 *   depth       float[n_frames*H*W]
 *   poses       double[n_frames*16]      pc_transform = tf @ base_transform @ base2cam_tf (memory_2.py:860)
 *   ord_base    uint32[n_frames]         ordinal of the frame's first visited point (frame-major, visiting order)
 *   sample_idx  int32[n_frames*n_points] visited pixel indices in visiting order (memory_2.py:746-749), or NULL
 *                                        for dense raster order (then n_points must equal H*W)
 *   tokens      float[n_frames*n_patch*D] patch tokens of each frame  (x_norm_patchtokens, memory_2.py:738-741)
 *   records_out optional uint4[n_frames*n_points] copy-out of the per-point records {cell | -1, patch | h<<16,
 *               ordinal, float32 bits of r^2} (parity tests), may be NULL
 */
int bscnav_fuse_frames(const bscnav_geom_t* g, const bscnav_store_t* st, const float* depth, const double* poses,
                       const uint32_t* ord_base, const int32_t* sample_idx, int32_t n_frames, int32_t n_points,
                       const float* tokens, void* scratch, int64_t scratch_bytes, void* records_out, void* stream);

/* Candidate filter of voxel_localized (memory_2.py:624-640). radius2 < 0 disables the sphere test.  The floor test is applied when
 * bit 0 of `flags` is set (an empty range floor_min > floor_max then selects nothing) or, for callers that leave flags 0, when
 * floor_min <= floor_max. */
typedef struct {
  int64_t radius2;                 /* region_radius^2 in cells^2 (compared as integers) */
  int32_t cur[3];                  /* curr_grid (row, col, h-minh)                       */
  int32_t floor_min, floor_max;    /* load_single_floor range, inclusive                 */
  int32_t reserved;                /* flags: bit 0 = floor test enabled                  */
} bscnav_filter_t;

int64_t bscnav_query_scratch_bytes(const bscnav_geom_t* g, int32_t n_query, int32_t k);

/* Cosine similarity of each query against every stored token, per-cell max, top-K (memory_2.py:642-671).
 *   q         float[n_query*D]
 *   out_sim   float[n_query*k]   descending; unused tail = -inf
 *   out_pos   int32[n_query*k*3] (row, col, h-minh); unused tail = -1
 *   out_count int32[n_query]     number of valid results (<= k);  k <= 1024
 * Ties are ordered by ascending (row, col, h). */
int bscnav_query_topk(const bscnav_geom_t* g, const bscnav_store_t* st, const float* q, int32_t n_query, int32_t k,
                      const bscnav_filter_t* filter, void* scratch, int64_t scratch_bytes,
                      float* out_sim, int32_t* out_pos, int32_t* out_count, void* stream);

/* Batched queries: the same result as bscnav_query_topk (bit-identical similarities and positions) for many queries per
 * pass over the store.  Candidates come from a tensor-core GEMM against a unit-normalised fp16 shadow of the token rows
 * (error bound 1.1e-3 on the cosine), are widened by a safety margin that provably contains the exact top-K, and are
 * re-scored with the exact fp32 arithmetic (csrc/query_batch.cu).
 *   shadow       half[align(max_slots*cache, 256) * dim], caller-owned; (re)build it with bscnav_store_build_shadow after the
 *                store changed (fuse / merge / load) and before querying
 *   n_rows_hint  upper bound on slots_in_use * cache known to the host (limits the GEMM to the used part of the store),
 *                or 0 for the full capacity
 *   out_count[i] = -1 flags a query whose candidate set exceeded the internal capacity (thousands of cells within 4e-3 of
 *                its K-th similarity): re-issue that query through bscnav_query_topk. */
int64_t bscnav_shadow_bytes(const bscnav_geom_t* g);
int bscnav_store_build_shadow(const bscnav_geom_t* g, const bscnav_store_t* st, void* shadow, void* stream);
int64_t bscnav_query_batch_scratch_bytes(const bscnav_geom_t* g, int32_t n_query, int32_t k, int64_t n_rows_hint);
int bscnav_query_topk_batch(const bscnav_geom_t* g, const bscnav_store_t* st, const void* shadow, const float* q,
                            int32_t n_query, int32_t k, const bscnav_filter_t* filter, int64_t n_rows_hint, void* scratch,
                            int64_t scratch_bytes, float* out_sim, int32_t* out_pos, int32_t* out_count, void* stream);

/* Cluster centres of each query's top-K cells (weighted_cluster_centers, BSCAgent.py:479-497): DBSCAN(eps, min_samples) over
 * the integer positions with scikit-learn's labelling (clusters numbered by their smallest core index, a border point joins
 * the lowest-numbered cluster that reaches it, noise = -1), then per cluster the similarity-weighted centre and the mean
 * similarity, clusters ordered by descending mean similarity.
 *   pos int32[n_query*k*3], sim float[n_query*k], count int32[n_query]   -- the outputs of bscnav_query_topk(_batch)
 *   out_centers double[n_query*max_clusters*3], out_mean_sim double[n_query*max_clusters], out_sizes int32[n_query*max_clusters],
 *   out_labels int32[n_query*k], out_n_clusters int32[n_query]; k <= 1024 */
int bscnav_cluster_centers(const int32_t* pos, const float* sim, const int32_t* count, int32_t n_query, int32_t k, double eps,
                           int32_t min_samples, int32_t max_clusters, double* out_centers, double* out_mean_sim,
                           int32_t* out_sizes, int32_t* out_labels, int32_t* out_n_clusters, void* stream);

/* Colour / occupancy side channels (memory_2.py:888-903; csrc/colour.cu), voxel keys only (levels > 1).  Optional: call
 * bscnav_colour_frames right after bscnav_fuse_frames with the same batch (same depth, poses, ord_base, sample_idx) from the
 * first batch on.  Results are those of the reference's sequential loop, bit for bit.
 *   grid_rgb uint8[max_slots*3]   alpha-weighted colour mean, cast to uint8 after every update   (memory_2.py:896-898)
 *   weight   float[max_slots]     accumulated alpha                                             (:899)
 *   cv_key   uint64[gs*gs]        0 = empty, else ((h-minh) << 32 | ordinal) + 1 of the highest point of the (row, col);
 *                                 max_height[row, col] = (cv_key - 1) >> 32, -inf where empty     (:901-902)
 *   cv_map   uint8[gs*gs*3]       its colour                                                     (:903)
 * rgb: uint8[n_frames*H*W*channels] (channels 3 or 4, alpha ignored).  counters[2] of the store tracks the slots whose colour
 * has been initialised. */
typedef struct {
  uint8_t*  grid_rgb;
  float*    weight;
  uint64_t* cv_key;
  uint8_t*  cv_map;
} bscnav_colour_t;
int bscnav_colour_reset(const bscnav_geom_t* g, const bscnav_colour_t* c, void* stream);
int64_t bscnav_colour_scratch_bytes(const bscnav_geom_t* g, int32_t n_frames, int32_t n_points);
int bscnav_colour_frames(const bscnav_geom_t* g, const bscnav_store_t* st, const bscnav_colour_t* c, const float* depth,
                         const uint8_t* rgb, int32_t channels, const double* poses, const uint32_t* ord_base,
                         const int32_t* sample_idx, int32_t n_frames, int32_t n_points, void* scratch, int64_t scratch_bytes,
                         void* stream);

/* Merge compacted deltas (e.g. all-gathered from other ranks) into the store under the same winner rule.
 *   d_cell int32[n], d_key uint64[n], d_h int32[n], d_feat float[n*D]; entries with d_cell < 0 are skipped.
 * Slot numbering of cells first seen through a merge follows delta order. */
int bscnav_merge_delta(const bscnav_geom_t* g, const bscnav_store_t* st, const int32_t* d_cell, const uint64_t* d_key,
                       const int32_t* d_h, const float* d_feat, int32_t n, void* scratch, int64_t scratch_bytes,
                       void* stream);

/* Export the rows whose key ordinal lies in [ord_lo, ord_hi) as a compact delta (for the all-gather merge).
 * out_n (device int32[2]): [0] = rows written (at most max_out), [1] = rows in range; [1] > max_out means the export was
 * truncated and must not be merged (the caller raises). */
int64_t bscnav_export_scratch_bytes(const bscnav_geom_t* g);
int bscnav_export_delta(const bscnav_geom_t* g, const bscnav_store_t* st, uint32_t ord_lo, uint32_t ord_hi,
                        int32_t* d_cell, uint64_t* d_key, int32_t* d_h, float* d_feat, int32_t max_out,
                        int32_t* out_n, void* scratch, int64_t scratch_bytes, void* stream);

/* The same delta as ONE packed buffer, the unit of the all-gather merge (one NCCL all-gather when a scene spans
 * ranks).  Row layout, bscnav_packed_row_bytes(g) = 16 + 4 D bytes: { int32 cell; int32 h; uint64 key; float feat[D] }.
 * out_n as in bscnav_export_delta.  bscnav_merge_packed merges a gathered buffer of n_segments x seg_capacity rows (segment s =
 * rank s's export, of which the first seg_counts[s * count_stride] rows are valid -- the all-gathered out_n arrays, read on the
 * device: no host round trip) under the store's winner rule in one pass; skip_segment (the caller's own rows; -1: none) is
 * ignored.  The result does not depend on the number or order of segments. */
int64_t bscnav_packed_row_bytes(const bscnav_geom_t* g);
int bscnav_export_packed(const bscnav_geom_t* g, const bscnav_store_t* st, uint32_t ord_lo, uint32_t ord_hi, void* packed,
                         int32_t max_out, int32_t* out_n, void* scratch, int64_t scratch_bytes, void* stream);
int64_t bscnav_merge_packed_scratch_bytes(int32_t n_segments, int32_t seg_capacity);
int bscnav_merge_packed(const bscnav_geom_t* g, const bscnav_store_t* st, const void* packed, int32_t n_segments,
                        int32_t seg_capacity, const int32_t* seg_counts, int32_t count_stride, int32_t skip_segment,
                        void* scratch, int64_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * ViT feature extractor (replaces dinov2.forward_features, memory_2.py:591-592,732-742). */
typedef struct {
  int32_t image, patch;           /* 224, 14                                           */
  int32_t dim, depth, heads, mlp; /* 1024, 24, 16, 4096                                */
  int32_t n_reg;                  /* register tokens after cls (DINOv2-reg: 4)         */
  int32_t act;                    /* 0 = GELU(erf), 1 = QuickGELU                      */
  int32_t ln_pre;                 /* 1 = CLIP pre-LayerNorm                            */
  int32_t patch_k_pad;            /* 3*patch*patch rounded up to a multiple of 64      */
  float   ln_eps;
  int32_t head_dim;               /* 0 or 64: dim == heads*64; 80: ViT-H/14 (mma.sync attention) */
  int32_t text_context;           /* 0: image tower; > 0: text tower over this many positions (causal attention,
                                     no patch embedding / cls / registers / pre-LN; image, patch, patch_k_pad unused) */
  int32_t precision;              /* 0: fp16 operands, fp32 accumulation (throughput path; range guard bscnav_vit_range_status);
                                     1: wide-range "bf16x3" path: every GEMM operand as a (hi, lo) pair of bf16, three tensor-core
                                     passes per product, fp32 attention -- fp32 range, ~1e-5 relative error, ~1/3 of the GEMM rate
                                     (head dim 64 only).  The half blob then holds, per weight matrix [N, K], rows of 2K bf16:
                                     [hi (K) | lo (K)] (vit.py:pack_weights(..., precision="bf16x3")) */
} bscnav_vit_cfg_t;

/* Element counts of the two packed weight blobs (layout documented in DESIGN.md / vit.py:pack_weights). */
int64_t bscnav_vit_half_elems(const bscnav_vit_cfg_t* c);
int64_t bscnav_vit_float_elems(const bscnav_vit_cfg_t* c);
int64_t bscnav_vit_workspace_bytes(const bscnav_vit_cfg_t* c, int32_t max_batch);

/* Create an engine bound to caller-owned weight blobs and workspace (TMA descriptors are encoded here). */
int bscnav_vit_create(const bscnav_vit_cfg_t* c, const void* w_half, const float* w_float, void* workspace,
                      int64_t workspace_bytes, int32_t max_batch, void** out_engine);
void bscnav_vit_destroy(void* engine);

/* Range guard of the fp16 activation path: the GEMM epilogues that write fp16 (qkv, fc1 + activation) set a device flag when a
 * rounded value is inf / nan, i.e. an activation left the fp16 range (trained ViTs have outlier channels; random-init weights
 * never do).  Reads the flag into *host_out (1 = overflow seen since the last reset) and synchronises `stream`; reset != 0
 * clears it.  The Python host raises on 1 instead of storing non-finite tokens. */
int bscnav_vit_range_status(void* engine, int32_t reset, int32_t* host_out, void* stream);

/* Optional per-kernel timing for the roofline report: when enabled, every launch of the next forward calls is
 * bracketed by CUDA events on the caller's stream.  profile_read synchronises on them and returns, per category
 * {0: tcgen05 GEMM, 1: attention, 2: LayerNorm, 3: pre-processing}, summed milliseconds, algorithmic work
 * (FLOPs for 0-1, bytes for 2-3) and launch counts (arrays of 4), then clears the record. */
int bscnav_vit_profile(void* engine, int32_t enable);
int bscnav_vit_profile_read(void* engine, double* out_ms, double* out_work, int32_t* out_launches);

/* Frames path (_get_patch_token, memory_2.py:732-742): uint8 HWC frames (channels = 3 or 4, alpha ignored)
 * -> /255 -> antialiased bilinear resize to image x image -> ImageNet normalise -> ViT -> patch tokens.
 *   out_tokens float[batch*n_patch*dim];  out_pre optional float[batch*3*image*image] (pre-processed image) */
int bscnav_vit_forward_frames(void* engine, const uint8_t* rgb, int32_t batch, int32_t height, int32_t width,
                              int32_t channels, float* out_tokens, float* out_pre, void* stream);

/* Query-image path (voxel_localized, memory_2.py:577-592): already normalised float[batch*3*image*image]. */
int bscnav_vit_forward_images(void* engine, const float* images, int32_t batch, float* out_tokens, void* stream);

/* Image-embedding path of the agent's view check (check_around, BSCAgent.py:700-731: clip_model.encode_image):
 * already normalised float[batch*3*image*image] -> ViT -> LayerNorm(post) of the cls token -> x proj (float[dim*embed_dim],
 * row-major [dim][embed_dim], no bias) -> out_embed float[batch*embed_dim]. */
int bscnav_vit_encode_images(void* engine, const float* images, int32_t batch, const float* proj, int32_t embed_dim,
                             float* out_embed, void* stream);

/* Text tower of the same CLIP (open_clip encode_text; goal text in check_around, BSCAgent.py:708-711, and long-term label
 * matching, :541-546): engine created from a cfg with text_context > 0.  token_ids int32[batch*context] (the caller's
 * tokenizer; ids outside [0, vocab) are clamped), token_embedding float[vocab*dim] -> + positional embedding -> causal
 * transformer -> LayerNorm(final) of the row at argmax(token_ids) (end-of-text) -> x proj float[dim*embed_dim]
 * -> out_embed float[batch*embed_dim]. */
int bscnav_text_encode(void* engine, const int32_t* token_ids, int32_t batch, const float* token_embedding, int32_t vocab,
                       const float* proj, int32_t embed_dim, float* out_embed, void* stream);

/* View scoring of check_around (BSCAgent.py:713-718, 726-731): L2-normalise the n view embeddings and the goal embedding,
 * similarities = views . goal, softmax over the n views -> out_prob float[n]; out_best int32[2] = {argmax (lowest index on
 * ties), 0}.  n <= 1024. */
int bscnav_score_views(const float* embeds, const float* goal, int32_t n, int32_t dim, float* out_prob, int32_t* out_best,
                       void* stream);

/* Gaussian centre-weighted pooling of query tokens (memory_2.py:596-608): float[batch*n_patch*dim] -> float[dim]. */
int bscnav_gaussian_pool(const float* tokens, int32_t batch, int32_t grid, int32_t dim, float* out, void* stream);

/* Stand-alone GEMM entry (tests/bench): C[M,N] = A[M,K] * W[N,K]^T + bias, fp16 operands, fp32 accumulate on
 * tcgen05; out_mode 0: fp16 out, 1: fp16 out with activation `act`, 2: fp32 out += (residual accumulate). */
int bscnav_gemm_f16(const void* a, const void* w, const float* bias, void* out, int32_t m, int32_t n, int32_t k,
                    int32_t out_mode, int32_t act, void* stream);

/* The stand-alone GEMM on the small-M path of the ViT engine (128-wide tiles, one CTA per tile), where the launcher may slice K
 * over idle SMs: every slice parks its fp32 partial tile in `scratch`, the slice arriving last sums them in slice order (the result
 * does not depend on arrival order) and runs the epilogue.  scratch: 16-byte aligned device memory, the first 1024 bytes (tile
 * counters) zero before the first call -- the kernel leaves them zero --, the rest partial tiles (64 KB each).  out_slices
 * (host, may be NULL): the slice count chosen for this shape (1 = not sliced). */
int bscnav_gemm_f16_sliced(const void* a, const void* w, const float* bias, void* out, int32_t m, int32_t n, int32_t k,
                           int32_t out_mode, int32_t act, void* scratch, int64_t scratch_bytes, int32_t* out_slices, void* stream);

/* Stand-alone wide-range GEMM (tests): a [M, 2K] and w [N, 2K] as (hi | lo) bf16 row pairs (hi = bf16(v), lo = bf16(v - hi)),
 * three tcgen05 passes a_hi w_hi + a_lo w_hi + a_hi w_lo, fp32 out [M, N] = a w^T + bias. */
int bscnav_gemm_bf16x3(const void* a, const void* w, const float* bias, float* out, int32_t m, int32_t n, int32_t k, void* stream);

/* Stand-alone attention entry (tests/bench): qkv fp16 [batch*T, 3*heads*64] (q pre-scaled by 1/sqrt(64)) ->
 * out fp16 [batch*T, heads*64].  impl 0: mma.sync kernel (any T); impl 1 / 2: first tcgen05 kernels; impl 3: the engine's default
 * (P in TMEM, SIMT tail rows, TMA-store epilogue; 17 <= T <= 272). */
int bscnav_attention_f16(const void* qkv, void* out, int32_t batch, int32_t T, int32_t heads, int32_t impl, void* stream);
/* The mma.sync kernel with an explicit head dim (64 | 80) and optional causal mask (causal = 1: keys <= query row; head dim 64):
 * qkv fp16 [batch*T, 3*heads*head_dim], q pre-scaled.  causal = 2: the same mask on the default tcgen05 kernel (head dim 64,
 * 17 <= T <= 272), the path the text tower takes. */
int bscnav_attention_f16_hd(const void* qkv, void* out, int32_t batch, int32_t T, int32_t heads, int32_t head_dim, int32_t causal,
                            void* stream);

/* Profiling hook of the default attention kernel (csrc/attention_ws2.cu): `flags` switch single stages off for timing
 * experiments (results are then wrong; 0 restores the product behaviour), `stamps` (device int64[512] or NULL) receives clock64
 * stamps of the first CTA's first 32 tiles.  Process-wide; not used by the product path. */
int bscnav_debug_attention_ws2(int32_t flags, void* stamps);

/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
int64_t bscnav_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BSCNAV_B200_H */
