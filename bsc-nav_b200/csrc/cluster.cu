// Cluster centres of a query's top-K cells on the device: DBSCAN(eps, min_samples) over the (row, col, h) positions,
// similarity-weighted centre and mean similarity per cluster, clusters ordered by mean similarity -- the tail of
// working_memory_retrival (BSCAgent.py:479-497, 640-641; twin objnav_benchmark.py:477-495), so that a retrieval needs no
// host round-trip between the top-K and the centres the agent navigates to.
//
// DBSCAN restated (scikit-learn 1.x `dbscan_inner`, the reference's dependency):
//   neighbours(i) = { j : |p_i - p_j|_2 <= eps }, i included;  core(i) <=> |neighbours(i)| >= min_samples
//   clusters are grown from unlabelled core points in index order, through core points only; a non-core point takes the
//   label of the first cluster that reaches it.  Equivalently: clusters = connected components of the core-core graph,
//   numbered by their smallest core index; a border point joins the lowest-numbered cluster owning one of its core
//   neighbours; everything else is noise (-1).
// Positions are integers, so the distance test is exact: d2 (int64) <= floor(eps^2).
// One block per query; K <= 1024.
#include "memory_kernels.cuh"

namespace bscnav {

static constexpr int kClThreads = 256;
static constexpr int kClMaxK = 1024;

__global__ void __launch_bounds__(kClThreads)
k_cluster_centers(const int32_t* __restrict__ pos, const float* __restrict__ sim, const int32_t* __restrict__ count, int k,
                  long long eps2, int min_samples, int max_clusters, double* __restrict__ centers, double* __restrict__ mean_sim,
                  int32_t* __restrict__ sizes, int32_t* __restrict__ labels, int32_t* __restrict__ n_clusters) {
  __shared__ int px[kClMaxK], py[kClMaxK], pz[kClMaxK];
  __shared__ int comp[kClMaxK];        // core: smallest core index of its component (converges); non-core: -1
  __shared__ int label[kClMaxK];
  __shared__ int n_roots;
  const int qi = blockIdx.x, tid = threadIdx.x;
  int n = count[qi];
  n = n < 0 ? 0 : (n > k ? k : n);
  for (int i = tid; i < n; i += kClThreads) {
    const int32_t* p = pos + ((long long)qi * k + i) * 3;
    px[i] = p[0]; py[i] = p[1]; pz[i] = p[2];
  }
  __syncthreads();
  auto near = [&](int i, int j) {
    const long long dx = px[i] - px[j], dy = py[i] - py[j], dz = pz[i] - pz[j];
    return dx * dx + dy * dy + dz * dz <= eps2;
  };
  for (int i = tid; i < n; i += kClThreads) {
    int deg = 0;
    for (int j = 0; j < n; ++j) deg += near(i, j) ? 1 : 0;
    comp[i] = deg >= min_samples ? i : -1;
  }
  __syncthreads();
  // components of the core-core graph by min-label propagation (K is small: a handful of sweeps)
  for (;;) {
    int changed = 0;
    for (int i = tid; i < n; i += kClThreads) {
      if (comp[i] < 0) continue;
      int m = comp[i];
      for (int j = 0; j < n; ++j)
        if (comp[j] >= 0 && comp[j] < m && near(i, j)) m = comp[j];
      if (m < comp[i]) { atomicMin(&comp[i], m); changed = 1; }
    }
    if (!__syncthreads_or(changed)) break;
  }
  // cluster number = rank of the component's smallest core index
  if (tid == 0) n_roots = 0;
  __syncthreads();
  for (int i = tid; i < n; i += kClThreads) {
    label[i] = -1;
    if (comp[i] == i) {
      int r = 0;
      for (int j = 0; j < i; ++j) r += comp[j] == j ? 1 : 0;
      label[i] = r;
      atomicAdd(&n_roots, 1);
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += kClThreads)
    if (comp[i] >= 0 && comp[i] != i) label[i] = label[comp[i]];
  __syncthreads();
  for (int i = tid; i < n; i += kClThreads) {
    if (comp[i] >= 0) continue;
    int best = -1;
    for (int j = 0; j < n; ++j)
      if (comp[j] >= 0 && near(i, j) && (best < 0 || label[j] < best)) best = label[j];
    label[i] = best;            // written after all reads of core labels (non-core labels are never read)
  }
  __syncthreads();
  for (int i = tid; i < k; i += kClThreads) labels[(long long)qi * k + i] = i < n ? label[i] : -1;
  // per cluster: mean similarity (pass A), then similarity-weighted centre (float64, members in index order), size and
  // the cluster's place in the output: descending mean similarity, ties keep cluster-number order (Python's stable sort)
  __shared__ double c_mean[kClMaxK];
  const int nr = n_roots;
  for (int c = tid; c < nr; c += kClThreads) {
    double sw = 0;
    int sz = 0;
    for (int i = 0; i < n; ++i)
      if (label[i] == c) { sw += (double)sim[(long long)qi * k + i]; ++sz; }
    c_mean[c] = sw / sz;
  }
  __syncthreads();
  for (int c = tid; c < nr; c += kClThreads) {
    int rank = 0;
    for (int o = 0; o < nr; ++o)
      if (c_mean[o] > c_mean[c] || (c_mean[o] == c_mean[c] && o < c)) ++rank;
    if (rank >= max_clusters) continue;
    double cx = 0, cy = 0, cz = 0, sw = 0;
    int sz = 0;
    for (int i = 0; i < n; ++i)
      if (label[i] == c) {
        const double w = (double)sim[(long long)qi * k + i];
        sw += w; cx += w * px[i]; cy += w * py[i]; cz += w * pz[i]; ++sz;
      }
    double* out = centers + ((long long)qi * max_clusters + rank) * 3;
    out[0] = cx / sw; out[1] = cy / sw; out[2] = cz / sw;
    mean_sim[(long long)qi * max_clusters + rank] = c_mean[c];
    sizes[(long long)qi * max_clusters + rank] = sz;
  }
  const int nc = nr < max_clusters ? nr : max_clusters;
  if (tid == 0) n_clusters[qi] = nc;
}

}  // namespace bscnav

using namespace bscnav;

extern "C" int bscnav_cluster_centers(const int32_t* pos, const float* sim, const int32_t* count, int32_t n_query, int32_t k,
                                      double eps, int32_t min_samples, int32_t max_clusters, double* out_centers,
                                      double* out_mean_sim, int32_t* out_sizes, int32_t* out_labels, int32_t* out_n_clusters,
                                      void* stream) {
  if (!pos || !sim || !count || !out_centers || !out_mean_sim || !out_sizes || !out_labels || !out_n_clusters) return BSCNAV_EINVAL;
  if (n_query <= 0 || k <= 0 || k > kClMaxK || !(eps >= 0) || min_samples <= 0 || max_clusters <= 0) return BSCNAV_EINVAL;
  const double e2 = eps * eps;
  const long long eps2 = e2 > 9.0e18 ? (long long)9.0e18 : (long long)e2;       // integer d2 <= eps^2  <=>  d2 <= floor(eps^2)
  BSC_LAUNCH(k_cluster_centers, (unsigned)n_query, kClThreads, 0, (cudaStream_t)stream, pos, sim, count, k, eps2, min_samples,
             max_clusters, out_centers, out_mean_sim, out_sizes, out_labels, out_n_clusters);
  return BSCNAV_OK;
}
