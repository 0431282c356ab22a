// knn_exact.cu -- FP64 CUDA-core pieces of the kNN graph build:
//   exact_filter_kernel : streams reference rows past a tile of query rows,
//                         FP64 direct-difference distances, bounded candidate
//                         buffers in shared memory with lazy pruning
//                         (the exact path, and the fallback for rows whose
//                         tensor-core certificate failed)
//   rerank_kernel       : per query row, FP64 distances of the candidate list,
//                         (d2, index) sort, exactness certificate, self removal
//                         and sklearn's rounding recipe
// Replaces sklearn ArgKmin32/64 (_argkmin.pyx.tp:163-307, _heap.pyx:46) reached
// from SemiBin/cluster.py:94-105 and long_read_cluster.py:108-113.
#include "common.cuh"
#include "knn.cuh"

namespace sbk {

static constexpr int XF_THREADS = 256;
static constexpr int XF_QT = 8;          // query rows per CTA

template <typename T>
__global__ void __launch_bounds__(XF_THREADS)
exact_filter_kernel(const T* __restrict__ X, long long n_ref, int d, long long ld,
                    const int32_t* __restrict__ rows, long long q_begin, long long n_q,
                    int n_split, int k1, int cap,
                    uint32_t* __restrict__ cand_idx, long long cand_stride,
                    int32_t* __restrict__ cand_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* qs = reinterpret_cast<double*>(smem_raw);                 // [d][QT]
  double* key = qs + (size_t)d * XF_QT;                             // [QT][cap]
  uint32_t* id = reinterpret_cast<uint32_t*>(key + (size_t)XF_QT * cap);   // [QT][cap]
  int* cnt = reinterpret_cast<int*>(id + (size_t)XF_QT * cap);      // [QT]
  double* thr = reinterpret_cast<double*>(cnt + XF_QT);             // [QT] (8B aligned: QT*4 = 32)

  const int tid = threadIdx.x;
  const long long slot0 = (long long)blockIdx.x * XF_QT;
  const int split = blockIdx.y;

  // query rows -> shared (double), dimension-major
  for (int e = tid; e < d * XF_QT; e += XF_THREADS) {
    int c = e / XF_QT, qi = e % XF_QT;
    long long slot = slot0 + qi;
    double v = 0.0;
    if (slot < n_q) {
      long long r = rows ? (long long)rows[slot] : (q_begin + slot);
      v = (double)X[r * ld + c];
    }
    qs[e] = v;
  }
  if (tid < XF_QT) { cnt[tid] = 0; thr[tid] = __longlong_as_double(0x7ff0000000000000LL); }
  __syncthreads();

  long long chunk = (n_ref + n_split - 1) / n_split;
  chunk = (chunk + XF_THREADS - 1) / XF_THREADS * XF_THREADS;
  const long long r_begin = (long long)split * chunk;
  const long long r_end = min(n_ref, r_begin + chunk);
  const double INF = __longlong_as_double(0x7ff0000000000000LL);

  for (long long base = r_begin; base < r_end; base += XF_THREADS) {
    const long long j = base + tid;
    const bool valid = j < r_end;
    double acc[XF_QT];
#pragma unroll
    for (int qi = 0; qi < XF_QT; ++qi) acc[qi] = 0.0;
    if (valid) {
      const T* y = X + j * ld;
      for (int c = 0; c < d; ++c) {
        const double yv = (double)__ldg(y + c);
        const double2* q2 = reinterpret_cast<const double2*>(qs + (size_t)c * XF_QT);
#pragma unroll
        for (int h = 0; h < XF_QT / 2; ++h) {
          double2 qq = q2[h];
          double t0 = qq.x - yv, t1 = qq.y - yv;
          acc[2 * h] = fma(t0, t0, acc[2 * h]);
          acc[2 * h + 1] = fma(t1, t1, acc[2 * h + 1]);
        }
      }
#pragma unroll
      for (int qi = 0; qi < XF_QT; ++qi) {
        if (acc[qi] < thr[qi]) {
          int pos = atomicAdd(&cnt[qi], 1);
          if (pos < cap) { key[(size_t)qi * cap + pos] = acc[qi]; id[(size_t)qi * cap + pos] = (uint32_t)j; }
        }
      }
    }
    __syncthreads();
    for (int qi = 0; qi < XF_QT; ++qi) {
      int n = cnt[qi];                               // uniform across the block
      if (n > cap - XF_THREADS) {
        n = min(n, cap);
        for (int t = n + tid; t < cap; t += XF_THREADS) { key[(size_t)qi * cap + t] = INF; id[(size_t)qi * cap + t] = 0xffffffffu; }
        block_bitonic_sort<XF_THREADS>(key + (size_t)qi * cap, id + (size_t)qi * cap, cap, tid);
        if (tid == 0) {
          cnt[qi] = min(n, k1);
          thr[qi] = (n >= k1) ? key[(size_t)qi * cap + k1 - 1] : INF;
        }
        __syncthreads();
      }
    }
  }
  // final prune + write-out: every split contributes exactly k1 slots
  for (int qi = 0; qi < XF_QT; ++qi) {
    long long slot = slot0 + qi;
    if (slot >= n_q) break;                          // uniform
    int n = min(cnt[qi], cap);
    __syncthreads();
    for (int t = n + tid; t < cap; t += XF_THREADS) { key[(size_t)qi * cap + t] = INF; id[(size_t)qi * cap + t] = 0xffffffffu; }
    block_bitonic_sort<XF_THREADS>(key + (size_t)qi * cap, id + (size_t)qi * cap, cap, tid);
    for (int t = tid; t < k1; t += XF_THREADS)
      cand_idx[slot * cand_stride + (long long)split * k1 + t] = id[(size_t)qi * cap + t];
    if (split == 0 && tid == 0) cand_count[slot] = n_split * k1;
  }
}

int exact_filter_cap(int k1) { return (int)pow2ceil((uint32_t)(k1 + XF_THREADS)); }

size_t exact_filter_smem(int d, int k1) {
  int cap = exact_filter_cap(k1);
  return (size_t)d * XF_QT * 8 + (size_t)XF_QT * cap * 12 + XF_QT * 4 + XF_QT * 8 + 16;
}

int launch_exact_filter(const void* X, int dtype, long long n_ref, int d, long long ld,
                        const int32_t* rows, long long q_begin, long long n_q, int n_split, int k1,
                        uint32_t* cand_idx, long long cand_stride, int32_t* cand_count,
                        cudaStream_t stream) {
  if (n_q <= 0) return SBK_OK;
  int cap = exact_filter_cap(k1);
  size_t smem = exact_filter_smem(d, k1);
  if (smem > 227 * 1024) { set_error("exact filter: k=%d d=%d needs %zu B shared memory", k1 - 1, d, smem); return SBK_ERR_INVALID; }
  dim3 grid((unsigned)((n_q + XF_QT - 1) / XF_QT), (unsigned)n_split);
  if (dtype == SBK_F32) {
    SBK_CUDA_CHECK(cudaFuncSetAttribute(exact_filter_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    exact_filter_kernel<float><<<grid, XF_THREADS, smem, stream>>>((const float*)X, n_ref, d, ld, rows, q_begin, n_q,
                                                                  n_split, k1, cap, cand_idx, cand_stride, cand_count);
  } else {
    SBK_CUDA_CHECK(cudaFuncSetAttribute(exact_filter_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    exact_filter_kernel<double><<<grid, XF_THREADS, smem, stream>>>((const double*)X, n_ref, d, ld, rows, q_begin, n_q,
                                                                    n_split, k1, cap, cand_idx, cand_stride, cand_count);
  }
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}

// ---------------------------------------------------------------------------
// rerank
// ---------------------------------------------------------------------------
static constexpr int RR_THREADS = 256;

// One CTA per candidate slot.
//   rows       : global row id per slot (NULL -> q_begin + slot)
//   bound      : per slot, every reference NOT in the candidate list is known to
//                have d2 > bound (NULL -> +inf, i.e. the list is a guaranteed superset)
//   fail_rows / n_fail : rows whose certificate failed are appended here (and
//                their output left untouched)
template <typename T, typename OutT>
__global__ void __launch_bounds__(RR_THREADS)
rerank_kernel(const T* __restrict__ X, long long n_ref, int d, long long ld,
              const int32_t* __restrict__ rows, long long q_begin,
              const uint32_t* __restrict__ cand_idx, long long cand_stride,
              const int32_t* __restrict__ cand_count, const double* __restrict__ bound,
              int k, int pmax,
              OutT* __restrict__ out_dist, int32_t* __restrict__ out_idx, long long out_row0,
              int32_t* __restrict__ fail_rows, int* __restrict__ n_fail,
              unsigned long long* __restrict__ cand_total) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* q = reinterpret_cast<double*>(smem_raw);           // [d]
  double* key = q + ((d + 1) & ~1);                          // [pmax]
  uint32_t* id = reinterpret_cast<uint32_t*>(key + pmax);    // [pmax]
  __shared__ int s_drop;

  const int tid = threadIdx.x;
  const long long slot = blockIdx.x;
  const long long row = rows ? (long long)rows[slot] : (q_begin + slot);
  const int k1 = k + 1;
  const double INF = __longlong_as_double(0x7ff0000000000000LL);

  int n = cand_count[slot];
  bool fail = (n > cand_stride) || (n < k1);
  if (fail) {
    if (tid == 0) { int p = atomicAdd(n_fail, 1); fail_rows[p] = (int32_t)row; }
    return;
  }
  for (int c = tid; c < d; c += RR_THREADS) q[c] = (double)X[row * ld + c];
  if (tid == 0) { s_drop = -1; if (cand_total) atomicAdd(cand_total, (unsigned long long)n); }
  int P = (int)pow2ceil((uint32_t)n);
  if (P < 2) P = 2;
  __syncthreads();
  const uint32_t* ci = cand_idx + slot * cand_stride;
  for (int c = tid; c < P; c += RR_THREADS) {
    double kk = INF; uint32_t jj = 0xffffffffu;
    if (c < n) {
      jj = ci[c];
      if (jj != 0xffffffffu) kk = sqdist_row<T>(q, X + (long long)jj * ld, d);
    }
    key[c] = kk; id[c] = jj;
  }
  block_bitonic_sort<RR_THREADS>(key, id, P, tid);
  // certificate: the (k+1)-th smallest exact d2 must be strictly below every
  // excluded reference's lower bound
  double kth = key[k];
  double b = bound ? bound[slot] : INF;
  if (!(kth < b) || id[k] == 0xffffffffu) {
    if (tid == 0) { int p = atomicAdd(n_fail, 1); fail_rows[p] = (int32_t)row; }
    return;
  }
  // self removal by index; if absent drop entry 0 (neighbors/_base.py:943-957)
  for (int t = tid; t < k1; t += RR_THREADS)
    if (id[t] == (uint32_t)row) s_drop = t;
  __syncthreads();
  const int drop = s_drop < 0 ? 0 : s_drop;
  const long long o = (row - out_row0) * (long long)k;
  for (int t = tid; t < k1; t += RR_THREADS) {
    if (t == drop) continue;
    int p = t - (t > drop ? 1 : 0);
    double d2 = key[t];
    if (sizeof(OutT) == 4) {
      // f64(f32(sqrt(f64(f32(d2)))))   sklearn _argkmin.pyx.tp:285-295
      float r = (float)d2;
      out_dist[o + p] = (OutT)(float)sqrt((double)r);
    } else {
      out_dist[o + p] = (OutT)sqrt(d2);
    }
    out_idx[o + p] = (int32_t)id[t];
  }
}

size_t rerank_smem(int d, int pmax) { return (size_t)((d + 1) & ~1) * 8 + (size_t)pmax * 12 + 16; }

int launch_rerank(const void* X, int dtype, long long n_ref, int d, long long ld,
                  const int32_t* rows, long long q_begin, long long n_slots,
                  const uint32_t* cand_idx, long long cand_stride, const int32_t* cand_count,
                  const double* bound, int k,
                  void* out_dist, int32_t* out_idx, long long out_row0,
                  int32_t* fail_rows, int* n_fail, unsigned long long* cand_total,
                  cudaStream_t stream) {
  if (n_slots <= 0) return SBK_OK;
  int pmax = (int)pow2ceil((uint32_t)cand_stride);
  size_t smem = rerank_smem(d, pmax);
  if (smem > 227 * 1024) { set_error("rerank: candidate capacity %lld too large", cand_stride); return SBK_ERR_INVALID; }
  dim3 grid((unsigned)n_slots);
  if (dtype == SBK_F32) {
    SBK_CUDA_CHECK(cudaFuncSetAttribute(rerank_kernel<float, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rerank_kernel<float, float><<<grid, RR_THREADS, smem, stream>>>(
        (const float*)X, n_ref, d, ld, rows, q_begin, cand_idx, cand_stride, cand_count, bound, k, pmax,
        (float*)out_dist, out_idx, out_row0, fail_rows, n_fail, cand_total);
  } else {
    SBK_CUDA_CHECK(cudaFuncSetAttribute(rerank_kernel<double, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rerank_kernel<double, double><<<grid, RR_THREADS, smem, stream>>>(
        (const double*)X, n_ref, d, ld, rows, q_begin, cand_idx, cand_stride, cand_count, bound, k, pmax,
        (double*)out_dist, out_idx, out_row0, fail_rows, n_fail, cand_total);
  }
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}

}  // namespace sbk
