// radius.cu -- long-read eps sweep over the distance-sorted kNN graph and
// DBSCAN label expansion (SemiBin/long_read_cluster.py:115-120 ->
// sklearn/cluster/_dbscan.py:430-463 and _dbscan_inner.pyx).
//
// radius_sweep_kernel : one warp per row, all eps values in one pass
//                       (N*k*8 B read once, N*n_eps*5 B written)
// label kernels       : label(v) = rank of min{u core : u reaches v through core
//                       points} -- iterated atomicMin propagation along the
//                       eps-prefix of every core row, then a rank scan.  This is
//                       the order-free form of sklearn's DFS (oracle/dbscan.py).
#include "common.cuh"
#include "scan.cuh"

namespace sbk {

static constexpr int R_MAX_EPS = 16;
static constexpr int R_WARPS = 8;

struct EpsList { double v[R_MAX_EPS]; };

__global__ void __launch_bounds__(R_WARPS * 32)
radius_sweep_kernel(const float* __restrict__ dist, const int32_t* __restrict__ idx,
                    long long row_begin, long long n_rows, int k, EpsList eps, int n_eps,
                    const double* __restrict__ weight, double min_samples,
                    int32_t* __restrict__ prefix_len, uint8_t* __restrict__ core) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * R_WARPS + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  int cnt[R_MAX_EPS];
  double ws[R_MAX_EPS];
#pragma unroll
  for (int e = 0; e < R_MAX_EPS; ++e) { cnt[e] = 0; ws[e] = 0.0; }
  for (int c = lane; c < k; c += 32) {
    const double d = (double)dist[r * k + c];
    const double wj = weight ? weight[idx[r * k + c]] : 1.0;
#pragma unroll
    for (int e = 0; e < R_MAX_EPS; ++e) {
      if (e < n_eps && d <= eps.v[e]) { cnt[e] += 1; ws[e] += wj; }
    }
  }
  const double wself = weight ? weight[row_begin + r] : 1.0;
#pragma unroll
  for (int e = 0; e < R_MAX_EPS; ++e) {
    if (e >= n_eps) break;
    int c = cnt[e]; double s = ws[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      c += __shfl_xor_sync(0xffffffffu, c, o);
      s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    if (lane == 0) {
      prefix_len[(long long)e * n_rows + r] = c;
      core[(long long)e * n_rows + r] = (wself + s >= min_samples) ? 1 : 0;
    }
  }
}

static constexpr int L_NONE = 0x7fffffff;

__global__ void label_init_kernel(const uint8_t* __restrict__ core, long long total, long long n,
                                  int32_t* __restrict__ L, uint8_t* __restrict__ dirty) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) { L[i] = core[i] ? (int32_t)(i % n) : L_NONE; dirty[i] = core[i]; }
}

// One warp per (eps, row): push L[u] to the eps-prefix of a core row.  Only rows whose label changed since
// their last push are "dirty" and do any work: after the first sweep the frontier is a small part of the graph,
// so a sweep costs one flag byte per row instead of the row's k indices.
__global__ void __launch_bounds__(R_WARPS * 32)
label_prop_kernel(const int32_t* __restrict__ idx, long long n, int k,
                  const int32_t* __restrict__ prefix_len, const uint8_t* __restrict__ core,
                  int n_eps, int32_t* __restrict__ L, uint8_t* __restrict__ dirty, int* __restrict__ changed) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * R_WARPS + (threadIdx.x >> 5);
  if (wid >= n * n_eps) return;
  if (!dirty[wid]) return;                       // non-core rows are never dirty
  const long long e = wid / n, u = wid % n;
  int32_t* Le = L + e * n;
  uint8_t* De = dirty + e * n;
  // clear first, then read the label: a concurrent improvement of L[u] either is seen by the read below or
  // sets the flag again after this clear
  if (lane == 0) De[u] = 0;
  __threadfence();
  int32_t lu = ((volatile int32_t*)Le)[u];
  lu = __shfl_sync(0xffffffffu, lu, 0);          // the value read by the lane that cleared the flag
  // pointer jumping: L[x] = y means core y reaches x, and L[y] reaches y, so L[y] reaches x as well; following
  // the chain to its root before pushing makes the number of sweeps logarithmic in the path length
  {
    int32_t root = lu;
    for (int hop = 0; hop < 64; ++hop) { const int32_t nx = ((volatile int32_t*)Le)[root]; if (nx >= root) break; root = nx; }
    root = __shfl_sync(0xffffffffu, root, 0);
    if (root < lu) { if (lane == 0) atomicMin(&Le[u], root); lu = root; }
  }
  const int plen = prefix_len[wid];
  bool any = false;
  for (int c = lane; c < plen; c += 32) {
    const int32_t j = idx[u * k + c];
    if (Le[j] > lu) {
      const int32_t old = atomicMin(&Le[j], lu);
      if (old > lu) { if (core[e * n + j]) De[j] = 1; any = true; }
    }
  }
  if (__any_sync(0xffffffffu, any) && lane == 0) *changed = 1;
}

__global__ void label_seed_kernel(const int32_t* __restrict__ L, long long total, long long n,
                                  int32_t* __restrict__ is_seed) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) is_seed[i] = (L[i] == (int32_t)(i % n)) ? 1 : 0;
}

__global__ void label_rank_kernel(const int32_t* __restrict__ L, long long total, long long n,
                                  const long long* __restrict__ seed_off, long long* __restrict__ labels) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long e = i / n;
  const int32_t l = L[i];
  labels[i] = (l == L_NONE) ? -1 : (seed_off[e * n + l] - seed_off[e * n]);
}

}  // namespace sbk

using namespace sbk;

extern "C" int sbk_radius_sweep(const float* dist, const int32_t* idx, int64_t row_begin, int64_t n_rows,
                                int32_t k, const double* eps, int32_t n_eps,
                                const double* weight, double min_samples,
                                int32_t* prefix_len, uint8_t* core, void* cuda_stream) {
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  if (n_eps < 1 || n_eps > R_MAX_EPS) { set_error("radius sweep: n_eps=%d outside [1,%d]", n_eps, R_MAX_EPS); return SBK_ERR_INVALID; }
  if (k < 1) { set_error("radius sweep: k=%d", k); return SBK_ERR_INVALID; }
  if (n_rows <= 0) return SBK_OK;
  EpsList el;
  for (int e = 0; e < R_MAX_EPS; ++e) el.v[e] = e < n_eps ? eps[e] : -1.0;   // eps is a HOST array
  unsigned blocks = (unsigned)((n_rows + R_WARPS - 1) / R_WARPS);
  radius_sweep_kernel<<<blocks, R_WARPS * 32, 0, stream>>>(dist, idx, row_begin, n_rows, k, el, n_eps, weight,
                                                          min_samples, prefix_len, core);
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}

extern "C" size_t sbk_dbscan_labels_workspace_bytes(int64_t n, int32_t n_eps) {
  Arena a(nullptr, 0);
  size_t total = (size_t)n * n_eps;
  a.take<int32_t>(total);            // L
  a.take<int32_t>(total + 1);        // is_seed
  a.take<long long>(total + 1);      // seed_off
  a.take<long long>(scan_workspace_elems((long long)total));
  a.take<int>(64);
  a.take<uint8_t>(total);            // dirty
  return a.off + 256;
}

extern "C" int sbk_dbscan_labels(const int32_t* idx, int64_t n, int32_t k,
                                 const int32_t* prefix_len, const uint8_t* core, int32_t n_eps,
                                 long long* labels, void* workspace, size_t ws_bytes, void* cuda_stream) {
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  if (n <= 0 || n_eps < 1) return SBK_OK;
  Arena a(workspace, ws_bytes);
  const long long total = (long long)n * n_eps;
  int32_t* L = a.take<int32_t>((size_t)total);
  int32_t* is_seed = a.take<int32_t>((size_t)total + 1);
  long long* seed_off = a.take<long long>((size_t)total + 1);
  long long* bsum = a.take<long long>(scan_workspace_elems(total));
  int* changed = a.take<int>(64);
  uint8_t* dirty = a.take<uint8_t>((size_t)total);
  if (!a.ok()) { set_error("dbscan labels: workspace %zu < %zu", ws_bytes, a.off); return SBK_ERR_WORKSPACE; }
  const unsigned eb = (unsigned)((total + 255) / 256);
  label_init_kernel<<<eb, 256, 0, stream>>>(core, total, n, L, dirty);
  SBK_LAUNCH_CHECK();
  const unsigned wb = (unsigned)((total + R_WARPS - 1) / R_WARPS);
  // Propagate to the fixpoint.  The host polls `changed` every few sweeps; the
  // number of sweeps is bounded by the longest shortest path (+ polling slack).
  const int sweeps_per_poll = 2;
  for (long long it = 0;; ++it) {
    SBK_CUDA_CHECK(cudaMemsetAsync(changed, 0, sizeof(int), stream));
    for (int s = 0; s < sweeps_per_poll; ++s) {
      label_prop_kernel<<<wb, R_WARPS * 32, 0, stream>>>(idx, n, k, prefix_len, core, n_eps, L, dirty, changed);
      SBK_LAUNCH_CHECK();
    }
    int h = 0;
    SBK_CUDA_CHECK(cudaMemcpyAsync(&h, changed, sizeof(int), cudaMemcpyDeviceToHost, stream));
    SBK_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (!h) break;
    if (it > n) { set_error("dbscan labels: propagation did not converge"); return SBK_ERR_INTERNAL; }
  }
  label_seed_kernel<<<eb, 256, 0, stream>>>(L, total, n, is_seed);
  SBK_LAUNCH_CHECK();
  int rc = exclusive_scan(is_seed, total, seed_off, bsum, stream);
  if (rc) return rc;
  label_rank_kernel<<<eb, 256, 0, stream>>>(L, total, n, seed_off, labels);
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}
