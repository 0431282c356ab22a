// kmeans.cu -- batched seeded, weighted Lloyd k-means: the re-clustering step of SemiBin/cluster.py:171-255
// (row N4 of SURVEY.md section 8).
//
// The reference loops over its bins on the host and calls, per bin with more than one seed,
//     KMeans(n_clusters=num_bin, init=seeds_embedding, n_init=1).fit(re_bin_features, sample_weight=length_weight)
// (cluster.py:236-241; scikit-learn 1.9.0 `lloyd`: _kmeans.py fit / _kmeans_single_lloyd, _k_means_lloyd.pyx,
// _k_means_common.pyx:167-211 relocation of empty clusters, :274-295 averaging).  A bin is a few hundred to a few
// thousand contigs x ~100 columns x a handful of clusters: far too small for a GPU one at a time, so ALL bins of a
// sample are solved by one launch: one CTA per bin runs the whole Lloyd loop (E-step, M-step, relocation, convergence
// tests, the final E-step) without leaving the device; bins only share the launch.
//
// Arithmetic: scikit-learn works on float32 data centred by the float32 column mean, with float32 accumulation in an
// order set by its chunking and thread count.  Here the same centred float32 values are used, every sum is FP64 in
// index order (deterministic), the E-step is the direct squared difference (first minimum wins ties, as the strict
// '<' scan of _k_means_lloyd.pyx).  Labels agree with scikit-learn except for points whose two nearest final centres
// tie to ~1e-4 relative (oracle/kmeans.py, tests/test_gpu_kmeans.py).
#include "common.cuh"
#include <math.h>
#include <algorithm>
#include <vector>

namespace sbk {

static constexpr int KM_THREADS = 256;
static constexpr int KM_MAX_K = 256;     // clusters per problem (seeds of one bin)

struct KmProblem {
  long long p0, p1;      // rows [p0, p1) of X
  int k0, k1;            // clusters [k0, k1) of centers_init / the centre workspace
  int id;                // the caller's problem index (the launch order is largest problem first)
};

__device__ __forceinline__ double km_block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int wp = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[wp] = v;
  __syncthreads();
  double t = 0.0;
  for (int q = 0; q < KM_THREADS / 32; ++q) t += s_red[q];
  return t;
}

// E-step for the CTA's points: labels[i] = argmin_j sum_c (x_ic - cen_jc)^2, first minimum.  Returns (block-wide) whether
// any label differs from labels_prev.
__device__ __forceinline__ bool km_assign(const float* __restrict__ X, long long ld, int d, long long p0, long long n,
                                          const float* __restrict__ meanf, const double* __restrict__ cen, int K,
                                          int32_t* __restrict__ labels, const int32_t* __restrict__ labels_prev) {
  int changed = 0;
  for (long long i = threadIdx.x; i < n; i += KM_THREADS) {
    const float* x = X + (p0 + i) * ld;
    double best = 0.0; int bj = 0;
    for (int j = 0; j < K; ++j) {
      const double* cj = cen + (long long)j * d;
      double a0 = 0.0, a1 = 0.0;
      int c = 0;
      for (; c + 1 < d; c += 2) {
        const double t0 = (double)(x[c] - meanf[c]) - cj[c], t1 = (double)(x[c + 1] - meanf[c + 1]) - cj[c + 1];
        a0 = fma(t0, t0, a0); a1 = fma(t1, t1, a1);
      }
      if (c < d) { const double t0 = (double)(x[c] - meanf[c]) - cj[c]; a0 = fma(t0, t0, a0); }
      const double s = a0 + a1;
      if (j == 0 || s < best) { best = s; bj = j; }
    }
    if (labels_prev && labels_prev[i] != bj) changed = 1;
    labels[i] = bj;
  }
  return __syncthreads_or(changed) != 0;
}

__global__ void __launch_bounds__(KM_THREADS)
kmeans_batch_kernel(const float* __restrict__ X, long long ld, int d, const float* __restrict__ weight,
                    const KmProblem* __restrict__ probs, const float* __restrict__ centers_init,
                    int max_iter, double tol_rel,
                    float* __restrict__ meanf_all, double* __restrict__ cen_a, double* __restrict__ cen_b,
                    double* __restrict__ wsum_all, int32_t* __restrict__ labels, int32_t* __restrict__ labels_old,
                    int32_t* __restrict__ n_iter_out, double* __restrict__ centers_out) {
  __shared__ double s_red[KM_THREADS / 32];
  __shared__ int s_flag;
  const KmProblem pr = probs[blockIdx.x];
  const long long p0 = pr.p0, n = pr.p1 - pr.p0;
  const int K = pr.k1 - pr.k0;
  float* meanf = meanf_all + (long long)blockIdx.x * d;
  double* cen = cen_a + (long long)pr.k0 * d;
  double* cen_new = cen_b + (long long)pr.k0 * d;
  double* wsum = wsum_all + pr.k0;
  int32_t* lab = labels + p0;
  int32_t* lab_old = labels_old + p0;
  if (n <= 0 || K <= 0) { if (threadIdx.x == 0) n_iter_out[pr.id] = 0; return; }

  // ---- fit(): centre the data and the seeds by the float32 column mean, tol = mean(var(X)) * tol_rel ----
  double var_sum = 0.0;
  for (int c = threadIdx.x; c < d; c += KM_THREADS) {
    double s = 0.0;
    for (long long i = 0; i < n; ++i) s += (double)X[(p0 + i) * ld + c];
    const float m = (float)(s / (double)n);
    meanf[c] = m;
    double mc = 0.0, v = 0.0;                       // variance of the CENTRED float32 column
    for (long long i = 0; i < n; ++i) mc += (double)(X[(p0 + i) * ld + c] - m);
    mc /= (double)n;
    for (long long i = 0; i < n; ++i) { const double t = (double)(X[(p0 + i) * ld + c] - m) - mc; v = fma(t, t, v); }
    var_sum += v / (double)n;
  }
  const double tol_abs = km_block_sum(var_sum, s_red) / (double)d * tol_rel;
  __syncthreads();
  for (int e = threadIdx.x; e < K * d; e += KM_THREADS) {
    const int c = e % d;
    cen[e] = (double)(centers_init[(long long)pr.k0 * d + e] - meanf[c]);
  }
  for (long long i = threadIdx.x; i < n; i += KM_THREADS) { lab[i] = -1; lab_old[i] = -1; }
  __syncthreads();

  bool strict = false;
  int it = 0;
  for (; it < max_iter; ++it) {
    // ---- E-step against `cen` ----
    const bool changed = km_assign(X, ld, d, p0, n, meanf, cen, K, lab, lab_old);
    // ---- M-step: weighted sums per (cluster, column), FP64, points in index order ----
    for (int e = threadIdx.x; e < K * d; e += KM_THREADS) {
      const int j = e / d, c = e % d;
      double s = 0.0;
      for (long long i = 0; i < n; ++i)
        if (lab[i] == j) s = fma((double)weight[p0 + i], (double)(X[(p0 + i) * ld + c] - meanf[c]), s);
      cen_new[e] = s;
    }
    for (int j = threadIdx.x; j < K; j += KM_THREADS) {
      double s = 0.0;
      for (long long i = 0; i < n; ++i) if (lab[i] == j) s += (double)weight[p0 + i];
      wsum[j] = s;
    }
    __syncthreads();
    // ---- empty clusters (_k_means_common.pyx:167-211): the farthest points from their centres seed them ----
    if (threadIdx.x == 0) { int ne = 0; for (int j = 0; j < K; ++j) ne += wsum[j] == 0.0 ? 1 : 0; s_flag = ne; }
    __syncthreads();
    if (s_flag > 0) {
      // serial and rare (a seed that lost every contig): thread 0 picks the farthest point for each empty cluster in turn
      if (threadIdx.x == 0) {
        const int n_empty = s_flag;
        double dmax_all = 0.0;
        for (long long i = 0; i < n; ++i) {
          const double* cj = cen + (long long)lab[i] * d;
          double s = 0.0;
          for (int c = 0; c < d; ++c) { const double t = (double)(X[(p0 + i) * ld + c] - meanf[c]) - cj[c]; s = fma(t, t, s); }
          dmax_all = fmax(dmax_all, s);
        }
        if (dmax_all > 0.0) {
          // the set of empty clusters is fixed up front (a cluster that gives its last point away is not refilled)
          for (int j = 0; j < K; ++j) if (wsum[j] == 0.0) wsum[j] = -1.0;
          double prev_d = INFINITY; long long prev_i = -1;
          int done = 0;
          for (int j = 0; j < K && done < n_empty; ++j) {
            if (wsum[j] != -1.0) continue;
            // next farthest point in (distance descending, index ascending) order
            double bd = -1.0; long long bi = -1;
            for (long long i = 0; i < n; ++i) {
              const double* cj = cen + (long long)lab[i] * d;
              double s = 0.0;
              for (int c = 0; c < d; ++c) { const double t = (double)(X[(p0 + i) * ld + c] - meanf[c]) - cj[c]; s = fma(t, t, s); }
              const bool after_prev = s < prev_d || (s == prev_d && i > prev_i);
              if (after_prev && (s > bd)) { bd = s; bi = i; }
            }
            if (bi < 0) break;
            prev_d = bd; prev_i = bi;
            const int old = lab[bi];
            const double w = (double)weight[p0 + bi];
            for (int c = 0; c < d; ++c) {
              const double xv = (double)(X[(p0 + bi) * ld + c] - meanf[c]) * w;
              cen_new[(long long)old * d + c] -= xv;
              cen_new[(long long)j * d + c] = xv;
            }
            wsum[j] = w; wsum[old] -= w;
            ++done;
          }
          for (int j = 0; j < K; ++j) if (wsum[j] == -1.0) wsum[j] = 0.0;
        }
      }
      __syncthreads();
    }
    // ---- average (sequential over clusters: an empty one copies the heaviest as it is at that moment, :274-295) ----
    if (s_flag > 0) {
      if (threadIdx.x == 0) {
        int amax = 0;
        for (int j = 1; j < K; ++j) if (wsum[j] > wsum[amax]) amax = j;
        for (int j = 0; j < K; ++j) {
          if (wsum[j] > 0.0) { const double a = 1.0 / wsum[j]; for (int c = 0; c < d; ++c) cen_new[(long long)j * d + c] *= a; }
          else for (int c = 0; c < d; ++c) cen_new[(long long)j * d + c] = cen_new[(long long)amax * d + c];
        }
      }
    } else {
      for (int e = threadIdx.x; e < K * d; e += KM_THREADS) cen_new[e] *= 1.0 / wsum[e / d];
    }
    __syncthreads();
    double sh = 0.0;
    for (int e = threadIdx.x; e < K * d; e += KM_THREADS) { const double t = cen_new[e] - cen[e]; sh = fma(t, t, sh); }
    const double shift_tot = km_block_sum(sh, s_red);
    { double* t = cen; cen = cen_new; cen_new = t; }
    if (!changed) { strict = true; ++it; break; }            // labels equal to the previous iteration's: converged
    if (shift_tot <= tol_abs) { ++it; break; }
    for (long long i = threadIdx.x; i < n; i += KM_THREADS) lab_old[i] = lab[i];
    __syncthreads();
  }
  if (!strict) km_assign(X, ld, d, p0, n, meanf, cen, K, lab, nullptr);      // labels that match the final centres
  if (threadIdx.x == 0) n_iter_out[pr.id] = it;
  if (centers_out)
    for (int e = threadIdx.x; e < K * d; e += KM_THREADS) centers_out[(long long)pr.k0 * d + e] = cen[e] + (double)meanf[e % d];
}

struct KmWs { KmProblem* probs; float* meanf; double* cen_a; double* cen_b; double* wsum; int32_t* labels_old; };
static void km_carve(Arena& a, long long n_total, int d, int n_prob, long long k_total, KmWs& w) {
  w.probs = a.take<KmProblem>((size_t)n_prob);
  w.meanf = a.take<float>((size_t)n_prob * d);
  w.cen_a = a.take<double>((size_t)k_total * d);
  w.cen_b = a.take<double>((size_t)k_total * d);
  w.wsum = a.take<double>((size_t)k_total);
  w.labels_old = a.take<int32_t>((size_t)n_total);
}

}  // namespace sbk

using namespace sbk;

extern "C" size_t sbk_kmeans_batch_workspace_bytes(int64_t n_total, int32_t d, int32_t n_problems, int64_t k_total) {
  if (n_total < 0 || d < 1 || n_problems < 0 || k_total < 0) return 0;
  Arena a(nullptr, 0);
  KmWs w;
  km_carve(a, n_total, d, std::max(n_problems, 1), std::max<int64_t>(k_total, 1), w);
  return a.off + 1024;
}

extern "C" int sbk_kmeans_batch(const float* X, int64_t ld, int32_t d, const float* weight, int64_t n_total,
                                const int64_t* point_offsets, const int32_t* cluster_offsets, int32_t n_problems,
                                const float* centers_init, int32_t max_iter, double tol,
                                int32_t* labels, int32_t* n_iter, double* centers_out,
                                void* workspace, size_t ws_bytes, void* cuda_stream) {
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  if (sbk_device_count() < 1) { set_error("kmeans: no CUDA device (there is no CPU fallback)"); return SBK_ERR_CUDA; }
  if (n_problems < 0 || d < 1 || ld < d || max_iter < 1 || !(tol >= 0.0)) { set_error("kmeans: bad arguments"); return SBK_ERR_INVALID; }
  if (n_problems == 0) return SBK_OK;
  if (!point_offsets || !cluster_offsets || !X || !weight || !centers_init || !labels || !n_iter) { set_error("kmeans: null argument"); return SBK_ERR_INVALID; }
  std::vector<KmProblem> hp((size_t)n_problems);
  for (int p = 0; p < n_problems; ++p) {
    hp[p].p0 = point_offsets[p]; hp[p].p1 = point_offsets[p + 1];
    hp[p].k0 = cluster_offsets[p]; hp[p].k1 = cluster_offsets[p + 1]; hp[p].id = p;
    if (hp[p].p1 < hp[p].p0 || hp[p].p1 > n_total || hp[p].k1 < hp[p].k0 || hp[p].k1 - hp[p].k0 > KM_MAX_K) {
      set_error("kmeans: problem %d has a bad range (points [%lld,%lld) of %lld, %d clusters, at most %d)", p, hp[p].p0, hp[p].p1,
                (long long)n_total, hp[p].k1 - hp[p].k0, KM_MAX_K);
      return SBK_ERR_INVALID;
    }
    // _kmeans.py:_check_params_vs_input: n_samples >= n_clusters
    if (hp[p].p1 - hp[p].p0 < hp[p].k1 - hp[p].k0) {
      set_error("n_samples=%lld should be >= n_clusters=%d.", hp[p].p1 - hp[p].p0, hp[p].k1 - hp[p].k0);
      return SBK_ERR_INVALID;
    }
  }
  const long long k_total = cluster_offsets[n_problems];
  // one CTA per problem and the CTAs start in launch order: the largest problems (points x clusters) go first so the
  // tail of the launch is made of small ones
  std::stable_sort(hp.begin(), hp.end(), [](const KmProblem& a, const KmProblem& b) {
    return (a.p1 - a.p0) * (long long)(a.k1 - a.k0) > (b.p1 - b.p0) * (long long)(b.k1 - b.k0);
  });
  Arena a(workspace, ws_bytes);
  KmWs w;
  km_carve(a, n_total, d, n_problems, std::max<long long>(k_total, 1), w);
  if (!a.ok() || workspace == nullptr) { set_error("kmeans: workspace %zu B < required %zu B", ws_bytes, a.off); return SBK_ERR_WORKSPACE; }
  SBK_CUDA_CHECK(cudaMemcpyAsync(w.probs, hp.data(), hp.size() * sizeof(KmProblem), cudaMemcpyHostToDevice, stream));
  SBK_CUDA_CHECK(cudaStreamSynchronize(stream));            // hp is a stack-lifetime host buffer
  kmeans_batch_kernel<<<(unsigned)n_problems, KM_THREADS, 0, stream>>>(X, ld, d, weight, w.probs, centers_init, max_iter, tol,
                                                                      w.meanf, w.cen_a, w.cen_b, w.wsum, labels, w.labels_old,
                                                                      n_iter, centers_out);
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}
