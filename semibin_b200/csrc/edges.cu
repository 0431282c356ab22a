// edges.cu -- short-read edge list after the two kNN graphs
// (SemiBin/cluster.py:109-151 and cal_kl :13-60 evaluated at edges only).
// All kernels are HBM-bound streaming passes over the N x k neighbour rows.
#include "common.cuh"
#include "scan.cuh"
#include <algorithm>

namespace sbk {

static constexpr int E_THREADS = 128;
static constexpr int E_MAXK = 1024;

// Stage 1: one warp per row.  An embedding-graph entry survives the intersection iff its column is also in the
// row of the feature graph and neither distance is an explicit zero (eliminate_zeros, cluster.py:109,112); the
// survivors are recorded as ONE BIT per entry (keep[r][c/32] bit c%32: 25 bytes per row at k = 200) and only
// the row maximum of w = 1 - min(d_emb, 1) (cluster.py:115-119) leaves the kernel -- stage 2 recomputes w from
// the float32 distance it reads anyway, so the N x k float64 matrix of weights is never written or read back.
// The feature-graph columns of the row go into a per-warp open-addressing hash set in shared memory
// (load factor <= 0.5), so the intersection is O(k) per row and the kernel streams at HBM speed
// instead of doing k^2 compares.
static constexpr int E1_WARPS = 4;
__global__ void __launch_bounds__(E1_WARPS * 32)
edges_stage1_kernel(const float* __restrict__ emb_dist, const int32_t* __restrict__ emb_idx,
                    const double* __restrict__ kmer_dist, const int32_t* __restrict__ kmer_idx,
                    long long n_rows, int k, int tab_size, uint32_t* __restrict__ keep, double* __restrict__ rowmax) {
  extern __shared__ int32_t e1_tab[];                           // [E1_WARPS][tab_size], tab_size = power of two >= 2k
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const long long r = (long long)blockIdx.x * E1_WARPS + wp;
  if (r >= n_rows) return;
  int32_t* tab = e1_tab + wp * tab_size;
  const uint32_t tmask = (uint32_t)tab_size - 1u;
  for (int t = lane; t < tab_size; t += 32) tab[t] = -1;
  __syncwarp();
  for (int c = lane; c < k; c += 32) {
    const int32_t j = kmer_idx[r * k + c];
    if (kmer_dist && kmer_dist[r * k + c] == 0.0) continue;     // explicit zero -> eliminated
    uint32_t h = ((uint32_t)j * 2654435761u) & tmask;
    for (;;) {
      const int32_t prev = atomicCAS(&tab[h], -1, j);
      if (prev == -1 || prev == j) break;
      h = (h + 1) & tmask;
    }
  }
  __syncwarp();
  const int words = (k + 31) >> 5;
  double mx = 0.0;                                              // implicit zeros count
  for (int c0 = 0; c0 < k; c0 += 32) {                          // warp-uniform trip count (ballot inside)
    const int c = c0 + lane;
    bool kept = false;
    if (c < k) {
      const int32_t j = emb_idx[r * k + c];
      const double d = (double)emb_dist[r * k + c];
      bool found = false;
      uint32_t h = ((uint32_t)j * 2654435761u) & tmask;
      for (;;) {
        const int32_t t = tab[h];
        if (t == j) { found = true; break; }
        if (t == -1) break;
        h = (h + 1) & tmask;
      }
      kept = found && d != 0.0;
      if (kept) mx = fmax(mx, 1.0 - fmin(d, 1.0));
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, kept);
    if (lane == 0) keep[r * words + (c0 >> 5)] = bits;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) rowmax[r] = mx;
}

// counts[m] += #{r : rowmax[r] > thresholds[m]}   (cluster.py:122)
__global__ void __launch_bounds__(256)
threshold_count_kernel(const double* __restrict__ rowmax, long long n,
                       const double* __restrict__ thresholds, int n_thr,
                       unsigned long long* __restrict__ counts) {
  __shared__ double thr[SBK_MAX_THRESHOLDS];
  __shared__ unsigned int blk[SBK_MAX_THRESHOLDS];
  if (threadIdx.x < SBK_MAX_THRESHOLDS) {
    thr[threadIdx.x] = threadIdx.x < n_thr ? thresholds[threadIdx.x] : 0.0;
    blk[threadIdx.x] = 0;
  }
  __syncthreads();
  unsigned int local[SBK_MAX_THRESHOLDS];
#pragma unroll
  for (int m = 0; m < SBK_MAX_THRESHOLDS; ++m) local[m] = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double v = rowmax[i];
#pragma unroll
    for (int m = 0; m < SBK_MAX_THRESHOLDS; ++m) local[m] += (m < n_thr && v > thr[m]) ? 1u : 0u;
  }
#pragma unroll
  for (int m = 0; m < SBK_MAX_THRESHOLDS; ++m) {
    unsigned int v = local[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&blk[m], v);
  }
  __syncthreads();
  if (threadIdx.x < n_thr && blk[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)blk[threadIdx.x]);
}

// Per contig and sample: (max(m, 1e-6), max(v, 1), float32 log of that v) -- the clamps of cluster.py:24-29 and
// the one transcendental of cal_kl, evaluated once per contig and sample (N*S logs) instead of once per edge and
// sample.  The float32 log is taken as the correctly rounded value of the float64 log (the
// reference's own evaluators -- numpy's SIMD log, numexpr's -- are not correctly rounded and disagree with each
// other; oracle.edges.kl_factor_bound derives the resulting tolerance).
__global__ void __launch_bounds__(256)
edges_mvl_kernel(const float* __restrict__ depth, long long n_total, int n_sample, float* __restrict__ mvl) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_total * n_sample) return;
  const float m = fmaxf(depth[2 * t], 1e-6f);
  const float v = fmaxf(depth[2 * t + 1], 1.0f);
  mvl[3 * t] = m; mvl[3 * t + 1] = v; mvl[3 * t + 2] = (float)log((double)v);
}

// float32 KL term of cal_kl (cluster.py:13-60, numpy evaluator :49-60), row i / col j:
//   (log v_j - log v_i)/2 + ((m_j - m_i)^2 + v_i) / (2 v_j) - 0.5, clipped to [1e-6, 1-1e-6]
// every operation individually rounded to float32 (no FMA contraction); halving is exact.
__device__ __forceinline__ float kl_term(float mi, float vi, float lvi, float mj, float vj, float lvj) {
  float v_div = __fsub_rn(lvj, lvi);
  v_div = __fmul_rn(v_div, 0.5f);
  float m_dif = __fsub_rn(mj, mi);
  m_dif = __fmul_rn(m_dif, m_dif);
  m_dif = __fadd_rn(m_dif, vi);
  m_dif = __fdiv_rn(m_dif, __fmul_rn(2.0f, vj));
  v_div = __fadd_rn(v_div, m_dif);
  v_div = __fsub_rn(v_div, 0.5f);
  const float lo = 1e-6f, hi = (float)(1.0 - 1e-6);
  return fminf(fmaxf(v_div, lo), hi);
}

// Stage 2a: one CTA per row: the intersection bits of stage 1, w = 1 - min(d, 1) (float64, cluster.py:115-116),
// threshold prune, KL factor, <=1e-6 cut, upper triangle, sort by column; result staged row-strided + per-row count.
__global__ void __launch_bounds__(E_THREADS)
edges_stage2_kernel(const uint32_t* __restrict__ keep_bits, const float* __restrict__ emb_dist,
                    const int32_t* __restrict__ emb_idx,
                    long long row_begin, int k, double threshold,
                    const float* __restrict__ mvl, int n_sample, int is_combined,
                    int32_t* __restrict__ stage_y, double* __restrict__ stage_v,
                    int32_t* __restrict__ row_cnt) {
  __shared__ double skey[E_MAXK];
  __shared__ uint32_t sid[E_MAXK];
  __shared__ double sval[E_MAXK];
  __shared__ int s_n;
  extern __shared__ float row_mvl[];                          // [n_sample][3]: the row's own (m, v, log v), clamped
  const long long r = blockIdx.x;
  const long long gi = row_begin + r;
  const int tid = threadIdx.x;
  const int words = (k + 31) >> 5;
  if (tid == 0) s_n = 0;
  if (!is_combined)
    for (int t = tid; t < 3 * n_sample; t += E_THREADS) row_mvl[t] = mvl[gi * 3 * n_sample + t];
  __syncthreads();
  for (int c = tid; c < k; c += E_THREADS) {
    const int32_t j = emb_idx[r * k + c];
    // only the upper triangle is emitted (cluster.py:146-148): test it first, the weight and the KL factor of a
    // dropped entry are never looked at
    bool keep = ((keep_bits[r * words + (c >> 5)] >> (c & 31)) & 1u) && ((long long)j > gi);
    double ww = 0.0;
    if (keep) {
      ww = 1.0 - fmin((double)emb_dist[r * k + c], 1.0);       // :115-116
      keep = (ww != 0.0) && !(ww <= threshold);                // :112-113 (eliminate_zeros), :127-128
    }
    if (keep && !is_combined) {
      const float* dj = mvl + (long long)j * 3 * n_sample;
      float acc = 0.f;
      for (int s = 0; s < n_sample; ++s) {
        const float kl = kl_term(row_mvl[3 * s], row_mvl[3 * s + 1], row_mvl[3 * s + 2], dj[3 * s], dj[3 * s + 1], dj[3 * s + 2]);
        acc = (s == 0) ? -kl : __fsub_rn(acc, kl);            // kl *= -1 ; kl_matrix -= kl
      }
      acc = __fadd_rn(acc, (float)n_sample);                   // += n_sample
      acc = __fdiv_rn(acc, (float)n_sample);                   // /= n_sample
      ww = ww * (double)acc;                                   // cluster.py:141 (f64 * f32)
    }
    keep = keep && !(ww <= 1e-6);                              // :143
    if (keep) {
      int p = atomicAdd(&s_n, 1);
      skey[p] = (double)j; sid[p] = (uint32_t)p; sval[p] = ww;
    }
  }
  __syncthreads();
  const int n = s_n;
  if (tid == 0) row_cnt[r] = n;
  if (n == 0) return;
  if (n <= 2 * E_THREADS) {
    // rank sort by column (columns of a row are distinct): no barriers, each entry lands at its rank
    for (int t = tid; t < n; t += E_THREADS) {
      const double kt = skey[t];
      int rk = 0;
#pragma unroll 4
      for (int u = 0; u < n; ++u) rk += skey[u] < kt ? 1 : 0;
      stage_y[r * k + rk] = (int32_t)kt;
      stage_v[r * k + rk] = sval[t];
    }
    return;
  }
  int P = (int)pow2ceil((uint32_t)n);
  for (int t = n + tid; t < P; t += E_THREADS) { skey[t] = __longlong_as_double(0x7ff0000000000000LL); sid[t] = 0xffffffffu; }
  block_bitonic_sort<E_THREADS>(skey, sid, P, tid);
  for (int t = tid; t < n; t += E_THREADS) {
    stage_y[r * k + t] = (int32_t)skey[t];
    stage_v[r * k + t] = sval[sid[t]];
  }
}

// Stage 2b: compaction to the row-major edge list.
__global__ void __launch_bounds__(E_THREADS)
edges_compact_kernel(const int32_t* __restrict__ stage_y, const double* __restrict__ stage_v,
                     const int32_t* __restrict__ row_cnt, const long long* __restrict__ row_off,
                     long long row_begin, long long n_rows, int k, long long capacity,
                     int32_t* __restrict__ out_x, int32_t* __restrict__ out_y, double* __restrict__ out_v,
                     long long* __restrict__ n_edges) {
  const long long r = blockIdx.x;
  const int n = row_cnt[r];
  const long long o = row_off[r];
  for (int t = threadIdx.x; t < n; t += E_THREADS) {
    if (o + t < capacity) {
      out_x[o + t] = (int32_t)(row_begin + r);
      out_y[o + t] = stage_y[r * k + t];
      out_v[o + t] = stage_v[r * k + t];
    }
  }
  if (r == 0 && threadIdx.x == 0) *n_edges = row_off[n_rows];
}

}  // namespace sbk

using namespace sbk;

extern "C" int sbk_edges_stage1(const float* emb_dist, const int32_t* emb_idx,
                                const double* kmer_dist, const int32_t* kmer_idx,
                                int64_t n_rows, int32_t k,
                                const double* thresholds, int32_t n_thresholds,
                                uint32_t* keep_bits, double* rowmax, long long* counts, void* cuda_stream) {
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  if (k < 1 || k > E_MAXK) { set_error("edges: k=%d outside [1,%d]", k, E_MAXK); return SBK_ERR_INVALID; }
  if (n_thresholds < 0 || n_thresholds > SBK_MAX_THRESHOLDS) { set_error("edges: n_thresholds=%d", n_thresholds); return SBK_ERR_INVALID; }
  if (n_rows <= 0) return SBK_OK;
  const int tab_size = (int)pow2ceil((uint32_t)(2 * k));
  const size_t smem1 = (size_t)E1_WARPS * tab_size * sizeof(int32_t);
  edges_stage1_kernel<<<(unsigned)((n_rows + E1_WARPS - 1) / E1_WARPS), E1_WARPS * 32, smem1, stream>>>(
      emb_dist, emb_idx, kmer_dist, kmer_idx, n_rows, k, tab_size, keep_bits, rowmax);
  SBK_LAUNCH_CHECK();
  if (n_thresholds > 0) {
    int blocks = (int)((n_rows + 256 * 16 - 1) / (256 * 16));
    if (blocks > 148 * 8) blocks = 148 * 8;
    threshold_count_kernel<<<blocks, 256, 0, stream>>>(rowmax, n_rows, thresholds, n_thresholds,
                                                       (unsigned long long*)counts);
    SBK_LAUNCH_CHECK();
  }
  return SBK_OK;
}

extern "C" size_t sbk_edges_stage2_workspace_bytes(int64_t n_rows, int32_t k, int64_t n_total, int32_t n_sample) {
  Arena a(nullptr, 0);
  a.take<int32_t>((size_t)n_rows * k);
  a.take<double>((size_t)n_rows * k);
  a.take<int32_t>((size_t)n_rows + 1);
  a.take<long long>((size_t)n_rows + 1);
  a.take<long long>(scan_workspace_elems(n_rows));
  a.take<float>((size_t)std::max<int64_t>(n_total, 0) * (size_t)std::max(n_sample, 0) * 3);
  return a.off + 256;
}

extern "C" int sbk_edges_stage2(const uint32_t* keep_bits, const float* emb_dist, const int32_t* emb_idx,
                                int64_t row_begin, int64_t n_rows, int32_t k, double threshold,
                                const float* depth, int64_t n_total, int32_t n_sample, int is_combined,
                                int32_t* out_x, int32_t* out_y, double* out_v, int64_t capacity,
                                long long* n_edges, void* workspace, size_t ws_bytes, void* cuda_stream) {
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  if (k < 1 || k > E_MAXK) { set_error("edges: k=%d outside [1,%d]", k, E_MAXK); return SBK_ERR_INVALID; }
  if (!is_combined && (n_sample < 1 || depth == nullptr || n_total < row_begin + n_rows)) {
    set_error("edges: KL weighting needs depth[n_total, 2*n_sample] with n_sample >= 1 and n_total >= row_begin + n_rows");
    return SBK_ERR_INVALID;
  }
  if (n_sample > 1024) { set_error("edges stage2: n_sample=%d > 1024", n_sample); return SBK_ERR_INVALID; }
  Arena a(workspace, ws_bytes);
  int32_t* stage_y = a.take<int32_t>((size_t)n_rows * k);
  double* stage_v = a.take<double>((size_t)n_rows * k);
  int32_t* row_cnt = a.take<int32_t>((size_t)n_rows + 1);
  long long* row_off = a.take<long long>((size_t)n_rows + 1);
  long long* bsum = a.take<long long>(scan_workspace_elems(n_rows));
  float* mvl = a.take<float>(is_combined ? 0 : (size_t)n_total * n_sample * 3);
  if (!a.ok()) { set_error("edges stage2: workspace %zu < %zu", ws_bytes, a.off); return SBK_ERR_WORKSPACE; }
  if (n_rows <= 0) { SBK_CUDA_CHECK(cudaMemsetAsync(n_edges, 0, sizeof(long long), stream)); return SBK_OK; }
  if (!is_combined) {
    const long long cells = (long long)n_total * n_sample;
    edges_mvl_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, stream>>>(depth, n_total, n_sample, mvl);
    SBK_LAUNCH_CHECK();
  }
  edges_stage2_kernel<<<(unsigned)n_rows, E_THREADS, (size_t)(is_combined ? 0 : n_sample) * 3 * sizeof(float), stream>>>(
      keep_bits, emb_dist, emb_idx, row_begin, k, threshold, mvl, n_sample, is_combined, stage_y, stage_v, row_cnt);
  SBK_LAUNCH_CHECK();
  int rc = exclusive_scan(row_cnt, n_rows, row_off, bsum, stream);
  if (rc) return rc;
  edges_compact_kernel<<<(unsigned)n_rows, E_THREADS, 0, stream>>>(stage_y, stage_v, row_cnt, row_off, row_begin, n_rows, k,
                                                                  capacity, out_x, out_y, out_v, n_edges);
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}
