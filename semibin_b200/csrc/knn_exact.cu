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
#include <algorithm>
#include <stdlib.h>
#include <cub/device/device_radix_sort.cuh>

namespace sbk {

static constexpr int XF_THREADS = 256;
// query rows per CTA: 8 (each streamed reference row is reused for 8 queries), or 1 for the handful of rows
// the tensor path hands back (the 7 idle query slots cost 8x the FP64 work and 4x the shared-memory reads)
template <typename T, int XF_QT>
__global__ void __launch_bounds__(XF_THREADS)
exact_filter_kernel(const T* __restrict__ X, long long n_ref, int d, long long ld,
                    const int32_t* __restrict__ rows, long long q_begin, long long n_q,
                    int n_split, int k1, int cap,
                    uint32_t* __restrict__ cand_idx, long long cand_stride,
                    int32_t* __restrict__ cand_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* qs = reinterpret_cast<double*>(smem_raw);                 // [d][QT]
  double* key = qs + (size_t)d * XF_QT;                             // [QT][cap]
  double* thr = key + (size_t)XF_QT * cap;                          // [QT]
  uint32_t* id = reinterpret_cast<uint32_t*>(thr + XF_QT);          // [QT][cap]
  int* cnt = reinterpret_cast<int*>(id + (size_t)XF_QT * cap);      // [QT]

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
      // STEP columns per step: the loads of a step are independent (memory-level parallelism), 16-byte vector
      // loads when the row is aligned.  With a single query row per CTA the arithmetic is negligible and the loop
      // is pure load latency: 48 (24) columns in flight per thread instead of 16.
      constexpr int STEP = XF_QT == 1 ? (sizeof(T) == 4 ? 48 : 24) : 16;
      constexpr int VEC = 16 / sizeof(T);
      const bool vec_ok = ((((uintptr_t)y) & 15) == 0);
      for (int c0 = 0; c0 < d; c0 += STEP) {
        T yv[STEP];
        const int nc = min(STEP, d - c0);
        if (vec_ok && nc == STEP) {
          if (sizeof(T) == 4) {
            const float4* y4 = reinterpret_cast<const float4*>(y + c0);
#pragma unroll
            for (int u = 0; u < STEP / VEC; ++u) {
              const float4 f = __ldg(y4 + u);
              yv[4 * u] = (T)f.x; yv[4 * u + 1] = (T)f.y; yv[4 * u + 2] = (T)f.z; yv[4 * u + 3] = (T)f.w;
            }
          } else {
            const double2* y2 = reinterpret_cast<const double2*>(y + c0);
#pragma unroll
            for (int u = 0; u < STEP / VEC; ++u) {
              const double2 f = __ldg(y2 + u);
              yv[2 * u] = (T)f.x; yv[2 * u + 1] = (T)f.y;
            }
          }
        } else {
#pragma unroll
          for (int u = 0; u < STEP; ++u) yv[u] = (u < nc) ? __ldg(y + c0 + u) : (T)0;
        }
#pragma unroll
        for (int u = 0; u < STEP; ++u) {
          if (u < nc) {
            const double yu = (double)yv[u];
            if (XF_QT == 1) {
              const double t0 = qs[c0 + u] - yu;
              acc[0] = fma(t0, t0, acc[0]);
            } else {
              const double2* q2 = reinterpret_cast<const double2*>(qs + (size_t)(c0 + u) * XF_QT);
#pragma unroll
              for (int h = 0; h < XF_QT / 2; ++h) {
                double2 qq = q2[h];
                double t0 = qq.x - yu, t1 = qq.y - yu;
                acc[2 * h] = fma(t0, t0, acc[2 * h]);
                acc[2 * h + 1] = fma(t1, t1, acc[2 * h + 1]);
              }
            }
          }
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
    // every thread snapshots the counts BETWEEN two barriers: a warp that finds nothing to prune must not run
    // ahead into the next batch's atomicAdd while a slower warp still reads cnt[] (it could then enter the sort,
    // which contains barriers, alone)
    int n_snap[XF_QT];
#pragma unroll
    for (int qi = 0; qi < XF_QT; ++qi) n_snap[qi] = cnt[qi];
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < XF_QT; ++qi) {
      int n = n_snap[qi];                            // uniform across the block
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

static constexpr int XF_QT_MAX = 8;
static constexpr long long XF_FEW_ROWS = 32;     // up to this many query rows: one row per CTA
size_t exact_filter_smem(int d, int k1) {
  int cap = exact_filter_cap(k1);
  return (size_t)d * XF_QT_MAX * 8 + (size_t)XF_QT_MAX * cap * 12 + XF_QT_MAX * 4 + XF_QT_MAX * 8 + 16;
}

int launch_exact_filter(const void* X, int dtype, long long n_ref, int d, long long ld,
                        const int32_t* rows, long long q_begin, long long n_q, int n_split, int k1,
                        uint32_t* cand_idx, long long cand_stride, int32_t* cand_count,
                        cudaStream_t stream) {
  if (n_q <= 0) return SBK_OK;
  int cap = exact_filter_cap(k1);
  size_t smem = exact_filter_smem(d, k1);
  if (smem > 227 * 1024) { set_error("exact filter: k=%d d=%d needs %zu B shared memory", k1 - 1, d, smem); return SBK_ERR_INVALID; }
#define SBK_XF_LAUNCH(T, QT)                                                                                           \
  do {                                                                                                                 \
    dim3 grid((unsigned)((n_q + QT - 1) / QT), (unsigned)n_split);                                                     \
    SBK_CUDA_CHECK(cudaFuncSetAttribute(exact_filter_kernel<T, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    exact_filter_kernel<T, QT><<<grid, XF_THREADS, smem, stream>>>((const T*)X, n_ref, d, ld, rows, q_begin, n_q,      \
                                                                   n_split, k1, cap, cand_idx, cand_stride, cand_count); \
  } while (0)
  const bool few = n_q <= XF_FEW_ROWS;
  if (dtype == SBK_F32) { if (few) SBK_XF_LAUNCH(float, 1); else SBK_XF_LAUNCH(float, 8); }
  else { if (few) SBK_XF_LAUNCH(double, 1); else SBK_XF_LAUNCH(double, 8); }
#undef SBK_XF_LAUNCH
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}

// ---------------------------------------------------------------------------
// rerank
// ---------------------------------------------------------------------------
static constexpr int RR_THREADS = 256;

// order-preserving map float -> uint32
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu));
}

// exact d2 of survivors id[lo..hi) -> key[lo..hi).  REGQ: LPC lanes per survivor, query slice in registers
// (RegQuery); otherwise 8 lanes per survivor, query in shared memory (sqdist_g8).  Entries 0xffffffff
// (padding of the exact path) get +inf.
template <typename T, bool REGQ>
__device__ __forceinline__ void rr_distances(const T* __restrict__ X, long long ld, int d, const double* q,
                                             const RegQuery<T>& rq, const uint32_t* id, double* key, int lo, int hi, int tid) {
  constexpr int LPC = REGQ ? DistCfg<T>::LPC : 8;
  const int sub = tid & (LPC - 1);
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  for (int c0 = lo; c0 < hi; c0 += RR_THREADS / LPC) {     // warp-uniform trip count (shuffles inside)
    const int c = c0 + tid / LPC;
    const bool ok = c < hi;
    const uint32_t jj = ok ? id[c] : 0xffffffffu;
    const bool real = jj != 0xffffffffu;
    const T* y = X + (long long)(real ? jj : 0u) * ld;
    const double v = REGQ ? rq.eval(y, sub) : sqdist_g8<T>(q, y, d, sub);
    if (ok && sub == 0) key[c] = real ? v : INF;
  }
}

// One CTA per candidate slot.
//   rows       : global row id per slot (NULL -> q_begin + slot)
//   cand_idx   : [slot][n_seg][seg_cap] candidate columns, cand_count[slot*n_seg+seg] valid per segment
//   cand_pair  : (column, tensor-core LOWER bound) pairs, same segmentation (tensor path; else NULL):
//                lower bounds L_ij <= S^2 d2_ij - nrm_i + sigma_i of the candidates (NULL on
//                the exact path, where every candidate is evaluated).  Tensor path, two phases:
//                  1. T = the candidates with the k+1 smallest L (radix select); exact FP64 d2 of T;
//                     dmax = max d2 over T  >=  the (k+1)-th smallest exact d2 of the row
//                  2. a candidate outside T can only matter if its lower bound allows d2 <= dmax, i.e.
//                     L_ij <= S^2 dmax - nrm_i + sigma_i; those are evaluated too, the rest are dropped
//   orig       : tensor path: candidate columns are reference POSITIONS of the staged (permuted) order; orig maps
//                them to rows when a candidate becomes a survivor
//   nrm / norms4 / scale : exact ||z_i||^2, (.., .., .., sigma_i) and S of the staging (tensor path)
//   bound      : per slot, every reference NOT in the candidate list is known to have d2 > bound
//                (NULL -> +inf, i.e. the list is a guaranteed superset)
//   fail_rows / n_fail : rows whose certificate failed are appended here (output left untouched)
template <typename T, typename OutT, bool REGQ>
__global__ void __launch_bounds__(RR_THREADS, REGQ ? 4 : 2)
rerank_kernel(const T* __restrict__ X, long long n_ref, int d, long long ld,
              const int32_t* __restrict__ rows, long long q_begin,
              const uint32_t* __restrict__ cand_idx, const uint2* __restrict__ cand_pair,
              long long cand_stride, int n_seg, const int32_t* __restrict__ cand_count,
              const double* __restrict__ bound, const double* __restrict__ nrm,
              const float4* __restrict__ norms4, const double* __restrict__ scale,
              const int32_t* __restrict__ orig, int k, int ncand_cap, int pmax,
              OutT* __restrict__ out_dist, int32_t* __restrict__ out_idx, long long out_row0,
              int32_t* __restrict__ fail_rows, int* __restrict__ n_fail,
              unsigned long long* __restrict__ totals,
              const int32_t* __restrict__ slot_list, const int* __restrict__ n_list, int list_cap, const PeerOut peers) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* q = reinterpret_cast<double*>(smem_raw);           // [d]
  double* key = q + ((d + 1) & ~1);                          // [pmax]  survivors: exact d2
  uint32_t* id = reinterpret_cast<uint32_t*>(key + pmax);    // [pmax]  survivors: column
  uint32_t* cid = id + pmax;                                 // [ncand_cap]  all candidates (tensor path only)
  float* cval = reinterpret_cast<float*>(cid + ncand_cap);   // [ncand_cap]  lower bounds L
  __shared__ uint32_t hist[512];
  __shared__ __align__(16) uint32_t s_khi[2 * RR_THREADS];   // rank sort: high words of the survivors' d2
  __shared__ uint16_t s_own[2 * RR_THREADS];                 // rank sort: who claimed each position
  __shared__ uint32_t s_prefix, s_remaining;
  __shared__ int s_drop, s_m, s_m2;
  __shared__ unsigned long long s_dmax;

  const int tid = threadIdx.x;
  long long slot = blockIdx.x;
  if (slot_list) {                               // retry mode: the slots the warp kernel handed back (device-side count)
    if ((int)blockIdx.x >= min(*n_list, list_cap)) return;
    slot = slot_list[blockIdx.x];
  }
  const long long row = rows ? (long long)rows[slot] : (q_begin + slot);
  const int k1 = k + 1;
  const int seg_cap = (int)(cand_stride / n_seg);
  const double INF = __longlong_as_double(0x7ff0000000000000LL);

  // segment sizes (uniform across the block)
  int seg_n[4], seg_off[4], n = 0;
  bool fail = false;
#pragma unroll
  for (int sgm = 0; sgm < 4; ++sgm) {
    seg_n[sgm] = 0; seg_off[sgm] = n;
    if (sgm < n_seg) {
      const int c = cand_count[slot * n_seg + sgm];
      fail |= (c > seg_cap);
      seg_n[sgm] = min(c, seg_cap);
      n += seg_n[sgm];
    }
  }
  if (fail || n < k1) {
    if (tid == 0) {
      int p = atomicAdd(n_fail, 1); fail_rows[p] = (int32_t)row;
      if (totals) atomicAdd(&totals[fail ? 2 : 3], 1ull);        // [2] segment overflow, [3] too few candidates
    }
    return;
  }
  RegQuery<T> rq;
  if (REGQ) {
    const bool vec_ok = ((((uintptr_t)X) & 15) == 0) && (((ld * (long long)sizeof(T)) & 15) == 0);
    rq.load(X + row * ld, d, tid & (DistCfg<T>::LPC - 1), vec_ok);
  } else {
    for (int c = tid; c < d; c += RR_THREADS) q[c] = (double)X[row * ld + c];
  }
  if (tid == 0) { s_drop = -1; s_m = 0; s_m2 = 0; s_dmax = 0ull; }
#pragma unroll
  for (int sgm = 0; sgm < 4; ++sgm) {
    if (cand_pair) {
      const uint2* cp = cand_pair + slot * cand_stride + (long long)sgm * seg_cap;
      for (int c = tid; c < seg_n[sgm]; c += RR_THREADS) {
        const uint2 e = cp[c];
        cid[seg_off[sgm] + c] = e.x; cval[seg_off[sgm] + c] = __uint_as_float(e.y);
      }
    } else {                                     // exact path: candidates ARE the survivors
      const uint32_t* ci = cand_idx + slot * cand_stride + (long long)sgm * seg_cap;
      for (int c = tid; c < seg_n[sgm]; c += RR_THREADS) id[seg_off[sgm] + c] = ci[c];
    }
  }
  hist[tid] = 0; hist[256 + tid] = 0;            // radix-select histograms (zeroed before the barrier below)
  __syncthreads();

  int m;
  if (cand_pair) {
    // ---- phase 1: the k+1 smallest lower bounds by radix select on the top 24 bits of the ordered float
    // (three 8-bit passes, two alternating histograms -> two barriers per pass).  Candidates that agree with
    // the (k+1)-th in those 24 bits are all taken: any superset of the k+1 smallest is a valid T.
    uint32_t prefix = 0, mask = 0, remaining = (uint32_t)k1;
    for (int shift = 24; shift >= 8; shift -= 8) {
      uint32_t* h = hist + ((shift >> 3) & 1) * 256;
      for (int c = tid; c < n; c += RR_THREADS) {
        const uint32_t u = f2ord(cval[c]);
        if ((u & mask) == prefix) atomicAdd(&h[(u >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid < 32) {
        uint32_t loc[8], sum = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) { loc[b] = h[tid * 8 + b]; sum += loc[b]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
          if (tid >= o) incl += y;
        }
        const uint32_t excl = incl - sum;
        if (excl < remaining && remaining <= incl) {
          uint32_t run = excl;
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            if (run < remaining && remaining <= run + loc[b]) {
              s_prefix = prefix | ((uint32_t)(tid * 8 + b) << shift);
              s_remaining = remaining - run;
            }
            run += loc[b];
          }
        }
      }
      __syncthreads();
      prefix = s_prefix; remaining = s_remaining; mask |= 0xffu << shift;
      h[tid] = 0;                                // reused two passes later, after two more barriers
    }
    const uint32_t kth_ord = prefix | 0xffu;     // every candidate sharing the (k+1)-th's top 24 bits is in T
    for (int c = tid; c < n; c += RR_THREADS) {
      if (f2ord(cval[c]) <= kth_ord) { const int p = atomicAdd(&s_m, 1); if (p < pmax) id[p] = (uint32_t)orig[cid[c]]; }
    }
    __syncthreads();
    const int m1 = s_m;                          // >= k+1 (more only on equal lower bounds)
    if (m1 > pmax) {                             // more survivors than the sort buffer: exact filter takes the row
      if (tid == 0) { int p = atomicAdd(n_fail, 1); fail_rows[p] = (int32_t)row; if (totals) atomicAdd(&totals[4], 1ull); }
      return;
    }
    rr_distances<T, REGQ>(X, ld, d, q, rq, id, key, 0, m1, tid);
    __syncthreads();
    {
      unsigned long long mx = 0ull;              // d2 >= 0: the bit pattern is monotone
      for (int c = tid; c < m1; c += RR_THREADS) mx = max(mx, (unsigned long long)__double_as_longlong(key[c]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if ((tid & 31) == 0) atomicMax(&s_dmax, mx);
    }
    __syncthreads();
    // ---- phase 2: candidates whose lower bound does not exclude d2 <= dmax ----
    {
      const double S = scale[0];
      const double dmax = __longlong_as_double((long long)s_dmax);
      // L_ij <= S^2 d2_ij - nrm_i + sigma_i; the inflation covers the rounding of dmax and of this expression
      const double cut = S * S * dmax * (1.0 + 1e-12) - nrm[row] * (1.0 - 1e-12) + (double)norms4[row].w + 1e-300;
      for (int c = tid; c < n; c += RR_THREADS) {
        const float l = cval[c];
        if (f2ord(l) > kth_ord && (double)l <= cut) { const int p = atomicAdd(&s_m, 1); if (p < pmax) id[p] = (uint32_t)orig[cid[c]]; }
      }
    }
    __syncthreads();
    m = s_m;
    if (m > pmax) {
      if (tid == 0) { int p = atomicAdd(n_fail, 1); fail_rows[p] = (int32_t)row; if (totals) atomicAdd(&totals[4], 1ull); }
      return;
    }
    rr_distances<T, REGQ>(X, ld, d, q, rq, id, key, m1, m, tid);
    __syncthreads();
    // T alone holds k+1 rows within dmax, so a phase-2 survivor farther than dmax cannot be among the k+1
    // nearest: compact [m1, m) down to those within dmax (keeps the sort at 256 entries for almost every row)
    {
      double kk[4]; uint32_t ii[4]; int nk = 0;
      bool lost = false;                           // pmax = 2048 (k >= 512): a thread may own up to 6 entries of [m1, m)
      const double dmax = __longlong_as_double((long long)s_dmax);
      for (int c = m1 + tid; c < m; c += RR_THREADS) {
        const double v = key[c];
        if (v <= dmax) { if (nk < 4) { kk[nk] = v; ii[nk] = id[c]; ++nk; } else lost = true; }
      }
      if (__syncthreads_or(lost ? 1 : 0)) {        // never drop a survivor silently: the exact filter takes the row
        if (tid == 0) { int p = atomicAdd(n_fail, 1); fail_rows[p] = (int32_t)row; if (totals) atomicAdd(&totals[4], 1ull); }
        return;
      }
      if (nk) {
        const int p0 = atomicAdd(&s_m2, nk);
#pragma unroll
        for (int u = 0; u < 4; ++u) if (u < nk) { key[m1 + p0 + u] = kk[u]; id[m1 + p0 + u] = ii[u]; }
      }
      __syncthreads();
      m = m1 + s_m2;
    }
  } else {
    m = n;
    // ---- exact path: FP64 distances of every candidate (padding entries 0xffffffff stay +inf) ----
    rr_distances<T, REGQ>(X, ld, d, q, rq, id, key, 0, m, tid);
  }
  if (tid == 0 && totals) { atomicAdd(&totals[0], (unsigned long long)n); atomicAdd(&totals[1], (unsigned long long)m); }

  if (m <= 2 * RR_THREADS) {
    // ---- rank sort: every thread counts how many survivors precede its own and drops its pair at that
    // position: two or three block barriers against the ~40 of a bitonic network.  Only the first k+1
    // positions are needed.  Fast pass: compare the HIGH WORDS of the d2 bit patterns only (d2 >= 0, so
    // the pattern is monotone; one broadcast 16-byte load and two integer instructions per pair instead of
    // two FP64 compares, an index compare and their selects).  Equal high words give equal counts, i.e. two
    // survivors claim the same position: detected through the `s_own` table, and only such rows (a few %)
    // repeat the count with the full (d2, index) comparison.  +inf padding of the exact path ranks last.
    __syncthreads();
    for (int c = tid; c < ((m + 3) & ~3); c += RR_THREADS)
      s_khi[c] = c < m ? (uint32_t)__double2hiint(key[c]) : 0x7ff00000u;   // pad: never "less"
    double kt[2]; uint32_t it[2]; int rk[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int c = tid + u * RR_THREADS;
      rk[u] = -1;
      if (c < m) { kt[u] = key[c]; it[u] = id[c]; rk[u] = 0; }
    }
    __syncthreads();
    if (rk[0] == 0) {
      const uint32_t h0 = (uint32_t)__double2hiint(kt[0]);
      const uint4* kh4 = reinterpret_cast<const uint4*>(s_khi);
      if (rk[1] == 0) {
        const uint32_t h1 = (uint32_t)__double2hiint(kt[1]);
        uint32_t r0 = 0, r1 = 0;
#pragma unroll 2
        for (int j = 0; j < (m + 3) / 4; ++j) {
          const uint4 hj = kh4[j];                 // high words < 2^31: the borrow of (hj - h) is "hj < h"
          r0 += ((hj.x - h0) >> 31) + ((hj.y - h0) >> 31) + ((hj.z - h0) >> 31) + ((hj.w - h0) >> 31);
          r1 += ((hj.x - h1) >> 31) + ((hj.y - h1) >> 31) + ((hj.z - h1) >> 31) + ((hj.w - h1) >> 31);
        }
        rk[0] = (int)r0; rk[1] = (int)r1;
      } else {
        uint32_t r0 = 0;
#pragma unroll 2
        for (int j = 0; j < (m + 3) / 4; ++j) {
          const uint4 hj = kh4[j];
          r0 += ((hj.x - h0) >> 31) + ((hj.y - h0) >> 31) + ((hj.z - h0) >> 31) + ((hj.w - h0) >> 31);
        }
        rk[0] = (int)r0;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (rk[u] >= 0 && rk[u] < k1) s_own[rk[u]] = (uint16_t)(tid + u * RR_THREADS);
    __syncthreads();
    bool clash = false;
#pragma unroll
    for (int u = 0; u < 2; ++u)
      clash |= (rk[u] >= 0 && rk[u] < k1 && s_own[rk[u]] != (uint16_t)(tid + u * RR_THREADS));
    if (__syncthreads_or(clash ? 1 : 0)) {
      // rare: full comparison (key/id are still untouched)
#pragma unroll
      for (int u = 0; u < 2; ++u) if (rk[u] >= 0) rk[u] = 0;
      if (rk[0] == 0) {
        if (rk[1] == 0) {
#pragma unroll 4
          for (int j = 0; j < m; ++j) {
            const double kj = key[j]; const uint32_t ij = id[j];
            rk[0] += (kj < kt[0] || (kj == kt[0] && ij < it[0])) ? 1 : 0;
            rk[1] += (kj < kt[1] || (kj == kt[1] && ij < it[1])) ? 1 : 0;
          }
        } else {
#pragma unroll 4
          for (int j = 0; j < m; ++j) {
            const double kj = key[j]; const uint32_t ij = id[j];
            rk[0] += (kj < kt[0] || (kj == kt[0] && ij < it[0])) ? 1 : 0;
          }
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (rk[u] >= 0 && rk[u] < k1) { key[rk[u]] = kt[u]; id[rk[u]] = it[u]; }
    __syncthreads();
  } else {
    int P = (int)pow2ceil((uint32_t)m);
    for (int c = m + tid; c < P; c += RR_THREADS) { key[c] = INF; id[c] = 0xffffffffu; }
    block_bitonic_sort<RR_THREADS>(key, id, P, tid);
  }
  // certificate: the (k+1)-th smallest exact d2 must be strictly below every
  // excluded reference's lower bound
  const double kth = (m >= k1) ? key[k] : INF;
  const double b = bound ? bound[slot] : INF;
  if (m < k1 || !(kth < b) || id[k] == 0xffffffffu) {
    if (tid == 0) { int p = atomicAdd(n_fail, 1); fail_rows[p] = (int32_t)row; if (totals) atomicAdd(&totals[5], 1ull); }
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
    OutT dv;
    if (sizeof(OutT) == 4) {
      // f64(f32(sqrt(f64(f32(d2)))))   sklearn _argkmin.pyx.tp:285-295
      float r = (float)d2;
      dv = (OutT)(float)sqrt((double)r);
    } else {
      dv = (OutT)sqrt(d2);
    }
    out_dist[o + p] = dv;
    out_idx[o + p] = (int32_t)id[t];
    for (int u = 0; u < peers.n; ++u) {          // the same row into the other GPUs' full result buffers (P2P stores)
      reinterpret_cast<OutT*>(peers.dist[u])[row * (long long)k + p] = dv;
      peers.idx[u][row * (long long)k + p] = (int32_t)id[t];
    }
  }
}

// ---------------------------------------------------------------------------
// warp-per-row rerank (tensor path, the common case)
// ---------------------------------------------------------------------------
// One warp owns one query row from its candidate list to the last store: no block barriers, 24 rows in flight per
// SM, the gathers of two candidate groups outstanding per lane.  Same two phases as rerank_kernel:
//   select  the lower bounds of the row's candidates (<= 24 per lane, in registers) are quantised to 16 bits over
//           their range and the (k+1)-th smallest is found bit by bit from warp-wide counts; A = everything up to
//           it (a superset of the k+1 smallest lower bounds); positions -> rows
//   phase 1 FP64 direct-difference d2 of A by groups of LPC lanes (RegQuery: the lane's slice of the query lives
//           in registers); dmax = max d2 over A >= the (k+1)-th smallest exact d2 of the row
//   phase 2 the candidates outside A whose lower bound still allows d2 <= dmax (a few per row) are evaluated too
//   sort    (d2, row) in REGISTERS by a bitonic network over the warp (32 * EPL elements, EPL per lane, blocked:
//           strides below EPL are register-to-register, the others one shuffle per word); d2 >= 0, so its bit
//           pattern orders like the value and (pattern, row) compares as integers
//   certificate, self removal and sklearn's rounding recipe as in rerank_kernel.
// Rows with more candidates than the registers hold or more survivors than the sort holds are handed to
// rerank_kernel through a device-side list (no host round trip).
static constexpr int RW_WARPS = 8;
__device__ __forceinline__ bool rw_less(unsigned long long ka, uint32_t ia, unsigned long long kb, uint32_t ib) {
  return (ka < kb) | ((ka == kb) & (ia < ib));      // bitwise on purpose: no short-circuit branches
}
// One compare-exchange layer (SIZE, STRIDE) of the bitonic network over NEL = 32 * EPL elements, element index
// i = lane * EPL + u; templates so that every register index is a compile-time constant.
template <int EPL, int SIZE, int STRIDE>
struct RwLayer {
  static __device__ __forceinline__ void run(unsigned long long (&kt)[EPL], uint32_t (&it)[EPL], int lane) {
    if (STRIDE >= EPL) {
      constexpr int lmask = STRIDE / EPL > 0 ? STRIDE / EPL : 1;
      const bool lower = (lane & lmask) == 0;
      const bool up = ((lane * EPL) & SIZE) == 0;                  // SIZE > STRIDE >= EPL: the bit lies in the lane part
#pragma unroll
      for (int u = 0; u < EPL; ++u) {
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, kt[u], lmask);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, it[u], lmask);
        // keep the minimum (lower == up) or the maximum; identical pairs (padding) may be exchanged freely
        const bool take = rw_less(ok, oi, kt[u], it[u]) == (lower == up);
        kt[u] = take ? ok : kt[u]; it[u] = take ? oi : it[u];
      }
    } else {
#pragma unroll
      for (int u = 0; u < EPL; ++u) {
        if ((u & STRIDE) == 0) {
          constexpr int vmask = STRIDE < EPL ? STRIDE : 0;
          const int v = u | vmask;
          const bool up = (((lane * EPL + u) & SIZE) == 0);
          const bool swap = rw_less(kt[v], it[v], kt[u], it[u]) == up;
          const unsigned long long lo_k = swap ? kt[v] : kt[u], hi_k = swap ? kt[u] : kt[v];
          const uint32_t lo_i = swap ? it[v] : it[u], hi_i = swap ? it[u] : it[v];
          kt[u] = lo_k; kt[v] = hi_k; it[u] = lo_i; it[v] = hi_i;
        }
      }
    }
    RwLayer<EPL, SIZE, STRIDE / 2>::run(kt, it, lane);
  }
};
template <int EPL, int SIZE>
struct RwLayer<EPL, SIZE, 0> {
  static __device__ __forceinline__ void run(unsigned long long (&)[EPL], uint32_t (&)[EPL], int) {}
};
template <int EPL, int SIZE>
struct RwSort {
  static __device__ __forceinline__ void run(unsigned long long (&kt)[EPL], uint32_t (&it)[EPL], int lane) {
    RwLayer<EPL, SIZE, SIZE / 2>::run(kt, it, lane);
    RwSort<EPL, SIZE * 2>::run(kt, it, lane);
  }
};
template <> struct RwSort<2, 128> { static __device__ __forceinline__ void run(unsigned long long (&)[2], uint32_t (&)[2], int) {} };
template <> struct RwSort<4, 256> { static __device__ __forceinline__ void run(unsigned long long (&)[4], uint32_t (&)[4], int) {} };
template <> struct RwSort<8, 512> { static __device__ __forceinline__ void run(unsigned long long (&)[8], uint32_t (&)[8], int) {} };
template <> struct RwSort<16, 1024> { static __device__ __forceinline__ void run(unsigned long long (&)[16], uint32_t (&)[16], int) {} };

// Lean FP64 row distance for the warp kernel: a group of LPC lanes evaluates one candidate row, lane `sub` owns the
// 16-byte chunks sub, sub + LPC, ... of every row; its slice of the query lives in registers as double.  Compile-time
// VEC (16-byte loads: rows 16-byte aligned) and precomputed per-chunk activity bits keep the inner loop free of
// branches and address arithmetic: one predicated LDG.128 per chunk from a per-lane base pointer.  Two rows per
// call so that both rows' loads are in flight before either is consumed.  Fixed xor-shuffle tree: deterministic,
// and bit-identical to RegQuery::eval / sqdist_g8 (same per-lane partial sums, same combination order).
template <typename T, bool VEC, bool FULL>
struct WarpDist {
  using C = DistCfg<T>;
  static constexpr int NV = C::NCH * C::V;
  double q[NV];
  double qt;
  int d, nfull, sub;
  unsigned act;                                        // bit j: chunk sub + LPC * j exists
  bool tail;                                           // this lane owns a tail column (d not a multiple of V)

  __device__ __forceinline__ void load(const T* __restrict__ x, int d_, int sub_) {
    d = d_; nfull = d_ / C::V; sub = sub_; act = 0u;
#pragma unroll
    for (int j = 0; j < C::NCH; ++j) {
      const int ch = sub_ + C::LPC * j;
      if (ch < nfull) act |= 1u << j;
#pragma unroll
      for (int e = 0; e < C::V; ++e) q[j * C::V + e] = ch < nfull ? (double)x[ch * C::V + e] : 0.0;
    }
    const int ct = nfull * C::V + sub_;
    tail = sub_ < C::V && ct < d_;
    qt = tail ? (double)x[ct] : 0.0;
  }

  // FULL: every chunk but the last exists for every lane (d > LPC * V * (NCH - 1)): those loads are unconditional
  // and need no zero fill -- the shapes of the path (D = 100..128 float32, 136 / 146 float64) all qualify
  __device__ __forceinline__ void fetch(const T* __restrict__ y, T (&v)[NV], T& vt) const {
#pragma unroll
    for (int j = 0; j < C::NCH; ++j) {
      const bool always = FULL && j < C::NCH - 1;
      if (!always) {
#pragma unroll
        for (int e = 0; e < C::V; ++e) v[j * C::V + e] = (T)0;
      }
      if (always || ((act >> j) & 1u)) {
        if (VEC) {
          if (sizeof(T) == 4) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(y) + sub + C::LPC * j);
            v[j * C::V] = (T)f.x; v[j * C::V + 1] = (T)f.y; v[j * C::V + 2] = (T)f.z; v[j * C::V + 3] = (T)f.w;
          } else {
            const double2 f = __ldg(reinterpret_cast<const double2*>(y) + sub + C::LPC * j);
            v[j * C::V] = (T)f.x; v[j * C::V + 1] = (T)f.y;
          }
        } else {
#pragma unroll
          for (int e = 0; e < C::V; ++e) v[j * C::V + e] = __ldg(y + (sub + C::LPC * j) * C::V + e);
        }
      }
    }
    vt = (T)0;
    if (tail) vt = __ldg(y + nfull * C::V + sub);
  }

  __device__ __forceinline__ double reduce(const T (&v)[NV], T vt) const {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (sizeof(T) == 4) {
#pragma unroll
      for (int j = 0; j < C::NCH; ++j) {               // inactive chunks hold zeros on both sides
        const double t0 = q[4 * j] - (double)v[4 * j], t1 = q[4 * j + 1] - (double)v[4 * j + 1];
        const double t2 = q[4 * j + 2] - (double)v[4 * j + 2], t3 = q[4 * j + 3] - (double)v[4 * j + 3];
        a0 = fma(t0, t0, a0); a1 = fma(t1, t1, a1); a2 = fma(t2, t2, a2); a3 = fma(t3, t3, a3);
      }
    } else {
#pragma unroll
      for (int j = 0; j < C::NCH; ++j) {
        const double t0 = q[2 * j] - (double)v[2 * j], t1 = q[2 * j + 1] - (double)v[2 * j + 1];
        if (j & 1) { a2 = fma(t0, t0, a2); a3 = fma(t1, t1, a3); }
        else { a0 = fma(t0, t0, a0); a1 = fma(t1, t1, a1); }
      }
    }
    { const double t = qt - (double)vt; a0 = fma(t, t, a0); }      // zero on both sides where there is no tail
    double s = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = C::LPC / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
  }
};

// FP64 squared distances of the rows id_s[lo..hi) -> key_s[lo..hi); returns the lane's running maximum.
// Two candidate groups per turn: 2 * NCH loads in flight per lane.
template <typename T, bool VEC, bool FULL>
__device__ __forceinline__ double rw_distances(const T* __restrict__ X, long long ld, const WarpDist<T, VEC, FULL>& wd,
                                               const uint32_t* id_s, double* key_s, int lo, int hi, int g, double mx) {
  constexpr int LPC = DistCfg<T>::LPC, CPI = 32 / LPC, NV = WarpDist<T, VEC, FULL>::NV;
#pragma unroll 1
  for (int e0 = lo; e0 < hi; e0 += 2 * CPI) {          // warp-uniform trip count (shuffles in reduce)
    const int ea = e0 + g, eb = e0 + CPI + g;
    const bool ra = ea < hi, rb = eb < hi;
    const uint32_t ja = id_s[ra ? ea : lo], jb = id_s[rb ? eb : lo];      // an idle group re-reads a valid row
    T va[NV], vb[NV]; T ta, tb;
    wd.fetch(X + (long long)ja * ld, va, ta);
    wd.fetch(X + (long long)jb * ld, vb, tb);
    const double da = wd.reduce(va, ta);
    const double db = wd.reduce(vb, tb);
    if (ra) { mx = fmax(mx, da); if (wd.sub == 0) key_s[ea] = da; }
    if (rb) { mx = fmax(mx, db); if (wd.sub == 0) key_s[eb] = db; }
  }
  return mx;
}
// Phase 2: the rows list[0..cnt) are evaluated the same way, but only those within dmax are appended to
// (key_s, id_s) at m (ballot compaction over the group leaders); returns the new m (may exceed cap: caller checks).
template <typename T, bool VEC, bool FULL>
__device__ __forceinline__ int rw_distances_keep(const T* __restrict__ X, long long ld, const WarpDist<T, VEC, FULL>& wd,
                                                 const uint32_t* list, int cnt, double dmax, double* key_s, uint32_t* id_s,
                                                 int m, int cap, int g, int lane) {
  constexpr int LPC = DistCfg<T>::LPC, CPI = 32 / LPC, NV = WarpDist<T, VEC, FULL>::NV;
#pragma unroll 1
  for (int e0 = 0; e0 < cnt; e0 += 2 * CPI) {
    const int ea = e0 + g, eb = e0 + CPI + g;
    const bool ra = ea < cnt, rb = eb < cnt;
    const uint32_t ja = list[ra ? ea : 0], jb = list[rb ? eb : 0];
    T va[NV], vb[NV]; T ta, tb;
    wd.fetch(X + (long long)ja * ld, va, ta);
    wd.fetch(X + (long long)jb * ld, vb, tb);
    const double da = wd.reduce(va, ta);
    const double db = wd.reduce(vb, tb);
    const bool ka = ra && wd.sub == 0 && da <= dmax, kb = rb && wd.sub == 0 && db <= dmax;
    const uint32_t bala = __ballot_sync(0xffffffffu, ka), balb = __ballot_sync(0xffffffffu, kb);
    const int ata = m + __popc(bala & ((1u << lane) - 1u));
    const int atb = m + __popc(bala) + __popc(balb & ((1u << lane) - 1u));
    if (ka && ata < cap) { key_s[ata] = da; id_s[ata] = ja; }
    if (kb && atb < cap) { key_s[atb] = db; id_s[atb] = jb; }
    m += __popc(bala) + __popc(balb);
  }
  return m;
}
__device__ __forceinline__ double rw_warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static constexpr int RW_QCAP = 3072;                   // candidates per row the select phase holds (16-bit keys in shared memory)
static constexpr int RW_P2CAP = 512;                   // phase-2 candidates per row (more: the block kernel takes the row)
static constexpr size_t rw_smem_bytes(int nel) { return (size_t)RW_WARPS * ((size_t)nel * 12 + RW_QCAP * 2 + RW_P2CAP * 4); }
template <typename T, int EPL, int MINB, bool VEC, bool FULL>
__global__ void __launch_bounds__(RW_WARPS * 32, MINB)
rerank_warp_kernel(const T* __restrict__ X, int d, long long ld, long long q_begin, long long n_rows,
                   const int32_t* __restrict__ perm, const int32_t* __restrict__ row_slot,
                   const uint2* __restrict__ cand, const int32_t* __restrict__ cand_count, int cand_cap,
                   const float* __restrict__ thr, const double* __restrict__ bound, const double* __restrict__ nrm,
                   const float4* __restrict__ norms4, const double* __restrict__ scale, const int32_t* __restrict__ orig, int k,
                   T* __restrict__ out_dist, int32_t* __restrict__ out_idx, long long out_row0,
                   int32_t* __restrict__ fail_rows, int* __restrict__ n_fail,
                   int32_t* __restrict__ retry_slots, int* __restrict__ n_retry, int retry_cap,
                   unsigned long long* __restrict__ totals, const PeerOut peers) {
  constexpr int NEL = 32 * EPL;
  constexpr int LPC = DistCfg<T>::LPC;
  extern __shared__ __align__(16) unsigned char rw_smem[];
  const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* key_s = reinterpret_cast<double*>(rw_smem) + (size_t)wp * NEL;                                       // [RW_WARPS][NEL]
  uint32_t* id_base = reinterpret_cast<uint32_t*>(reinterpret_cast<double*>(rw_smem) + RW_WARPS * NEL);
  uint32_t* id_s = id_base + (size_t)wp * NEL;                                                                 // [RW_WARPS][NEL]
  uint32_t* p2_base = id_base + RW_WARPS * NEL;
  uint32_t* p2_s = p2_base + (size_t)wp * RW_P2CAP;                                                            // [RW_WARPS][RW_P2CAP]
  uint16_t* q_s = reinterpret_cast<uint16_t*>(p2_base + RW_WARPS * RW_P2CAP) + (size_t)wp * RW_QCAP;           // [RW_WARPS][RW_QCAP]
  const long long widx = (long long)blockIdx.x * RW_WARPS + wp;
  if (widx >= n_rows) return;
  // rows are visited in the locality order of launch_row_order (neighbouring warps work on neighbouring rows of the
  // feature space: their gathers hit the same lines of L2 / L1); the lists and the outputs stay indexed by slot / row
  const long long rloc = perm ? (long long)perm[widx] : widx;
  const long long row = q_begin + rloc;
  // lists, thresholds and bounds are indexed by slot: the row's offset in the chunk, or (symmetric filter) its staged position
  const long long slot = row_slot ? (long long)row_slot[row] : rloc;
  const int k1 = k + 1;
  const int seg_cap = cand_cap >> 2;
  // the four segments of the row as one list: element e lives at base[e + adj(segment of e)]
  int n, p1, p2, p3; bool bad = false;
  {
    int c0 = cand_count[slot * 4], c1 = cand_count[slot * 4 + 1], c2 = cand_count[slot * 4 + 2], c3 = cand_count[slot * 4 + 3];
    bad = c0 > seg_cap || c1 > seg_cap || c2 > seg_cap || c3 > seg_cap;
    c0 = min(c0, seg_cap); c1 = min(c1, seg_cap); c2 = min(c2, seg_cap); c3 = min(c3, seg_cap);
    p1 = c0; p2 = c0 + c1; p3 = p2 + c2; n = p3 + c3;
  }
  const int adj1 = seg_cap - p1, adj2 = 2 * seg_cap - p2, adj3 = 3 * seg_cap - p3;
  auto elem = [&](int e) { return e + (e >= p3 ? adj3 : e >= p2 ? adj2 : e >= p1 ? adj1 : 0); };
  // exits: overflow / too few candidates -> exact fallback; too many for this kernel -> block-per-row rerank
  auto hand_back = [&](int why) {                        // why: 2 overflow, 3 too few, -1 retry
    if (lane == 0) {
      bool to_fail = why >= 0;
      if (why < 0) {
        const int p = atomicAdd(n_retry, 1);
        if (p < retry_cap) { retry_slots[p] = (int32_t)slot; if (totals) atomicAdd(&totals[6], 1ull); }
        else { to_fail = true; why = 4; }
      }
      if (to_fail) { const int p = atomicAdd(n_fail, 1); fail_rows[p] = (int32_t)row; if (totals) atomicAdd(&totals[why], 1ull); }
    }
  };
  if (bad || n < k1) { hand_back(bad ? 2 : 3); return; }
  if (n > RW_QCAP) { hand_back(-1); return; }
  const uint2* base = cand + slot * (long long)cand_cap;
  // ---- select: lower bounds quantised to 13 bits between the row's own score and its final threshold ----
  // Every candidate of the later launches has L <= thr (the thresholds only ever dropped), the k+1 smallest with
  // the planned confidence; candidates above it (admitted by an earlier, looser threshold) clamp to the top level.
  // The self score  -||z_i||^2  is the smallest possible lower bound up to the slack.  8192 levels over that range
  // are several per neighbour; a coarser level only puts a few more rows into A.
  float lo_f, sc_f;
  {
    const float hi_f = thr[slot];
    lo_f = (float)(-nrm[row]) - 0.125f * fabsf(hi_f + (float)nrm[row]) - 1.0f;
    const float range = hi_f - lo_f;
    sc_f = range > 0.f ? 8191.0f / range : 0.f;
  }
  const int npad = (n + 127) & ~127;                     // four keys per lane and turn; RW_QCAP is a multiple of 128
  for (int e = lane; e < npad; e += 32) {
    uint32_t qv = 8191u;                                 // padding: never below any test value
    if (e < n) qv = min(8191u, __float2uint_rz(fmaxf(__uint_as_float(base[elem(e)].y) - lo_f, 0.f) * sc_f));
    q_s[e] = (uint16_t)qv;
  }
  __syncwarp();
  uint32_t res = 0;                                      // the (k+1)-th smallest quantised value, bit by bit
  {
    // four 13-bit keys per 8-byte load; "key >= t" for both halves of a word at once: with the guard bit 15 set,
    // (half - t) keeps bit 15 exactly when half >= t and never borrows from the other half
    const uint2* q4 = reinterpret_cast<const uint2*>(q_s);
    const int nw = npad >> 2;                            // multiple of 32
    for (int bit = 12; bit >= 0; --bit) {
      const uint32_t t = res | (1u << bit);
      const uint32_t tt = t | (t << 16);
      uint32_t ge = 0;                                   // packed counters (both halves), at most 96 * 2 each
#pragma unroll 2
      for (int i = lane; i < nw; i += 32) {
        const uint2 w2 = q4[i];
        ge += ((((w2.x | 0x80008000u) - tt) >> 15) & 0x00010001u) + ((((w2.y | 0x80008000u) - tt) >> 15) & 0x00010001u);
      }
      int cnt_ge = (int)((ge & 0xffffu) + (ge >> 16));
      cnt_ge = __reduce_add_sync(0xffffffffu, cnt_ge);
      if (npad - cnt_ge < k1) res = t;                   // fewer than k+1 keys below t
    }
  }
  // A = {q <= res}: at least k+1 members, compacted as positions, then mapped to rows in one parallel pass
  int mA = 0;
  for (int e0 = 0; e0 < n; e0 += 32) {
    const int e = e0 + lane;
    const bool inA = e < n && (uint32_t)q_s[e] <= res;
    const uint32_t bal = __ballot_sync(0xffffffffu, inA);
    const int at = mA + __popc(bal & ((1u << lane) - 1u));
    if (inA && at < NEL) id_s[at] = base[elem(e)].x;
    mA += __popc(bal);
  }
  if (mA > NEL) { hand_back(-1); return; }
  __syncwarp();
  for (int i = lane; i < mA; i += 32) id_s[i] = (uint32_t)orig[id_s[i]];
  __syncwarp();
  const int sub = lane & (LPC - 1), g = lane / LPC;
  WarpDist<T, VEC, FULL> wd;
  wd.load(X + row * ld, d, sub);
  // ---- phase 1 ----
  const double dmax = rw_warp_max(rw_distances<T, VEC, FULL>(X, ld, wd, id_s, key_s, 0, mA, g, 0.0));
  // ---- phase 2: candidates outside A whose lower bound does not exclude d2 <= dmax ----
  // Dense neighbourhoods put dozens of candidates inside the slack band: they are listed (positions, one scan with
  // independent loads), mapped to rows in one parallel pass, evaluated like A, and only those within dmax are kept
  // (A alone holds k+1 rows within dmax, so nothing farther can be among the k+1 nearest): the sort sees ~k+1 entries.
  int m = mA, m2 = 0;
  {
    const double S = scale[0];
    // L_ij <= S^2 d2_ij - nrm_i + sigma_i; the inflation covers the rounding of dmax and of this expression
    const double cut = S * S * dmax * (1.0 + 1e-12) - nrm[row] * (1.0 - 1e-12) + (double)norms4[row].w + 1e-300;
#pragma unroll 2
    for (int e0 = 0; e0 < n; e0 += 32) {
      const int e = e0 + lane;
      bool take = false; uint32_t pos = 0;
      if (e < n && (uint32_t)q_s[e] > res) {
        const uint2 pr = base[elem(e)];
        take = (double)__uint_as_float(pr.y) <= cut; pos = pr.x;
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, take);
      const int at = m2 + __popc(bal & ((1u << lane) - 1u));
      if (take && at < RW_P2CAP) p2_s[at] = pos;
      m2 += __popc(bal);
    }
  }
  if (m2 > RW_P2CAP) { hand_back(-1); return; }
  if (m2 > 0) {
    __syncwarp();
    for (int i = lane; i < m2; i += 32) p2_s[i] = (uint32_t)orig[p2_s[i]];
    __syncwarp();
    m = rw_distances_keep<T, VEC, FULL>(X, ld, wd, p2_s, m2, dmax, key_s, id_s, mA, NEL, g, lane);
    if (m > NEL) { hand_back(-1); return; }
  }
  const long long n_eval = (long long)mA + m2;
  if (lane == 0 && totals) { atomicAdd(&totals[0], (unsigned long long)n); atomicAdd(&totals[1], (unsigned long long)n_eval); }
  __syncwarp();
  // ---- (d2 pattern, row) pairs into registers; the initial placement is irrelevant to a sort ----
  unsigned long long kt[EPL]; uint32_t it[EPL];
#pragma unroll
  for (int u = 0; u < EPL; ++u) {
    const int e = u * 32 + lane;
    kt[u] = e < m ? (unsigned long long)__double_as_longlong(key_s[e]) : 0x7ff0000000000000ull;
    it[u] = e < m ? id_s[e] : 0xffffffffu;
  }
  // ---- bitonic network, element index i = lane * EPL + u ----
  RwSort<EPL, 2>::run(kt, it, lane);
  // sorted pairs back to shared memory at their rank: the certificate reads rank k, the stores below walk the
  // ranks lane-strided (coalesced) and the FP64 sqrt stays out of an unrolled register loop
  __syncwarp();
#pragma unroll
  for (int u = 0; u < EPL; ++u) {
    const int r = lane * EPL + u;
    key_s[r] = __longlong_as_double((long long)kt[u]);
    id_s[r] = it[u];
  }
  __syncwarp();
  // ---- certificate: the (k+1)-th smallest exact d2 strictly below every excluded reference's lower bound ----
  const double kth = key_s[k];
  const uint32_t kth_id = id_s[k];
  const double b = bound[slot];
  if (m < k1 || kth_id == 0xffffffffu || !(kth < b)) {
    if (lane == 0) { const int p = atomicAdd(n_fail, 1); fail_rows[p] = (int32_t)row; if (totals) atomicAdd(&totals[5], 1ull); }
    return;
  }
  // ---- self removal by index; if absent drop entry 0 (neighbors/_base.py:943-957) ----
  int self_rank = 0x7fffffff;
#pragma unroll
  for (int u = 0; u < EPL; ++u) if (it[u] == (uint32_t)row && lane * EPL + u < k1) self_rank = lane * EPL + u;
  self_rank = __reduce_min_sync(0xffffffffu, self_rank);
  const int drop = self_rank == 0x7fffffff ? 0 : self_rank;
  const long long o = (row - out_row0) * (long long)k;
#pragma unroll 1
  for (int r = lane; r < k1; r += 32) {
    if (r == drop) continue;
    const int p = r - (r > drop ? 1 : 0);
    const double d2 = key_s[r];
    T dv;
    if (sizeof(T) == 4) dv = (T)(float)sqrt((double)(float)d2);   // f64(f32(sqrt(f64(f32(d2)))))  _argkmin.pyx.tp:285-295
    else dv = (T)sqrt(d2);
    const int32_t iv = (int32_t)id_s[r];
    out_dist[o + p] = dv;
    out_idx[o + p] = iv;
    for (int u = 0; u < peers.n; ++u) {          // the same row into the other GPUs' full result buffers (P2P stores)
      reinterpret_cast<T*>(peers.dist[u])[row * (long long)k + p] = dv;
      peers.idx[u][row * (long long)k + p] = iv;
    }
  }
}

#ifdef SBK_ENABLE_PROBES
static int rw_env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
#else
static int rw_env_int(const char*, int dflt) { return dflt; }
#endif

int rerank_warp_cap(int k) {          // sort capacity (power of two >= k+1 with room for the phase-2 survivors), 0 = not supported
  const int need = k + 1 + std::max(8, (k + 1) / 8);
  for (int c = 64; c <= 512; c <<= 1) if (c >= need) return c;
  return 0;
}

template <typename T, int EPL, int MINB, bool VEC, bool FULL>
static int rw_launch(unsigned grid, cudaStream_t stream, const T* X, int d, long long ld, long long q_begin, long long n_rows,
                     const int32_t* perm, const int32_t* row_slot,
                     const uint2* cand, const int32_t* cand_count, int cand_cap, const float* thr, const double* bound,
                     const double* nrm, const float4* norms4, const double* scale, const int32_t* orig, int k,
                     T* out_dist, int32_t* out_idx, long long out_row0, int32_t* fail_rows, int* n_fail,
                     int32_t* retry_slots, int* n_retry, int retry_cap, unsigned long long* totals, const PeerOut& peers) {
  const size_t smem = rw_smem_bytes(32 * EPL);
  SBK_CUDA_CHECK(cudaFuncSetAttribute(rerank_warp_kernel<T, EPL, MINB, VEC, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rerank_warp_kernel<T, EPL, MINB, VEC, FULL><<<grid, RW_WARPS * 32, smem, stream>>>(X, d, ld, q_begin, n_rows, perm, row_slot, cand, cand_count, cand_cap,
      thr, bound, nrm, norms4, scale, orig, k, out_dist, out_idx, out_row0, fail_rows, n_fail, retry_slots, n_retry, retry_cap, totals, peers);
  return SBK_OK;
}

int launch_rerank_warp(const void* X, int dtype, int d, long long ld, long long q_begin, long long n_rows, const int32_t* perm,
                       const int32_t* row_slot,
                       const uint2* cand, const int32_t* cand_count, int cand_cap, const float* thr, const double* bound,
                       const double* nrm, const float4* norms4, const double* scale, const int32_t* orig, int k,
                       void* out_dist, int32_t* out_idx, long long out_row0, int32_t* fail_rows, int* n_fail,
                       int32_t* retry_slots, int* n_retry, int retry_cap, unsigned long long* totals, cudaStream_t stream,
                       const PeerOut* peers) {
  if (n_rows <= 0) return SBK_OK;
  PeerOut po; po.n = 0;
  if (peers) po = *peers;
  const int cap = rerank_warp_cap(k);
  if (cap == 0) { set_error("rerank: k=%d not supported by the warp kernel", k); return SBK_ERR_INVALID; }
  const unsigned grid = (unsigned)((n_rows + RW_WARPS - 1) / RW_WARPS);
  const size_t esz = dtype == SBK_F32 ? 4 : 8;
  const bool vec = ((((uintptr_t)X) & 15) == 0) && (((ld * (long long)esz) & 15) == 0);
  // every chunk but the last exists for every lane: d above LPC * V * (NCH - 1) columns (96 float32 / 128 float64)
  const bool full = dtype == SBK_F32 ? d > DistCfg<float>::LPC * DistCfg<float>::V * (DistCfg<float>::NCH - 1)
                                     : d > DistCfg<double>::LPC * DistCfg<double>::V * (DistCfg<double>::NCH - 1);
  // 2 CTAs (16 warps) per SM at up to 128 registers by default; the probe build can ask for 3 CTAs at 80 registers
  static const int occ = rw_env_int("SBK_RW_OCC", 2);
  int rc = SBK_OK;
#define SBK_RW_ARGS(T) grid, stream, (const T*)X, d, ld, q_begin, n_rows, perm, row_slot, cand, cand_count, cand_cap, thr, bound, nrm, norms4, scale, orig, k, \
      (T*)out_dist, out_idx, out_row0, fail_rows, n_fail, retry_slots, n_retry, retry_cap, totals, po
#define SBK_RW_LAUNCH2(T, EPL, MINB)                                                                                    \
  do {                                                                                                                  \
    if (vec && full) rc = rw_launch<T, EPL, MINB, true, true>(SBK_RW_ARGS(T));                                          \
    else if (vec) rc = rw_launch<T, EPL, MINB, true, false>(SBK_RW_ARGS(T));                                            \
    else rc = rw_launch<T, EPL, MINB, false, false>(SBK_RW_ARGS(T));                                                    \
  } while (0)
#define SBK_RW_LAUNCH(T, EPL) do { if (occ == 3) SBK_RW_LAUNCH2(T, EPL, 3); else SBK_RW_LAUNCH2(T, EPL, 2); } while (0)
#define SBK_RW_DISPATCH(T)                                                                                              \
  do {                                                                                                                  \
    if (cap <= 64) SBK_RW_LAUNCH(T, 2); else if (cap <= 128) SBK_RW_LAUNCH(T, 4);                                       \
    else if (cap <= 256) SBK_RW_LAUNCH(T, 8); else SBK_RW_LAUNCH(T, 16);                                                \
  } while (0)
  if (dtype == SBK_F32) SBK_RW_DISPATCH(float); else SBK_RW_DISPATCH(double);
#undef SBK_RW_DISPATCH
#undef SBK_RW_LAUNCH
#undef SBK_RW_LAUNCH2
#undef SBK_RW_ARGS
  if (rc) return rc;
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}

// ---------------------------------------------------------------------------
// locality order of the query rows for the rerank
// ---------------------------------------------------------------------------
// The rerank gathers ~k+1 reference rows per query; queries that sit close together in the feature space gather
// (nearly) the same rows.  Visiting the queries grouped by a 16-bit random-hyperplane signature of the centred row
// turns most of that traffic into L2 / L1 hits instead of 80+ GB of DRAM gathers in input order.  Any order is
// correct; this one only has to be cheap: one pass over the chunk's rows (16 sign-sums per row) + a 16-bit radix sort.
__global__ void __launch_bounds__(256)
row_signature_kernel(const void* __restrict__ Xv, int is_f64, int d, long long ld, long long q_begin, long long n_rows,
                     const double* __restrict__ centre, uint32_t* __restrict__ keys, int32_t* __restrict__ vals,
                     const int32_t* __restrict__ rows) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_rows) return;
  float acc[16];
#pragma unroll
  for (int b = 0; b < 16; ++b) acc[b] = 0.f;
  const long long r = rows ? (long long)rows[s] : q_begin + s;      // a row list: the order is over those rows, values = row ids - q_begin
  for (int c = 0; c < d; ++c) {
    const float v = is_f64 ? (float)(reinterpret_cast<const double*>(Xv)[r * ld + c] - centre[c])
                           : (float)((double)reinterpret_cast<const float*>(Xv)[r * ld + c] - centre[c]);
    uint32_t h = (uint32_t)c * 2654435761u;               // 16 pseudo-random signs per column
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
#pragma unroll
    for (int b = 0; b < 16; ++b) acc[b] += ((h >> b) & 1u) ? v : -v;
  }
  uint32_t key = 0;
#pragma unroll
  for (int b = 0; b < 16; ++b) key |= (acc[b] > 0.f ? 1u : 0u) << b;
  keys[s] = key; vals[s] = rows ? (int32_t)(r - q_begin) : (int32_t)s;
}

size_t row_order_workspace_bytes(long long n_rows) {
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)std::max<long long>(n_rows, 1), 0, 16, (cudaStream_t)0);
  return align_up((size_t)n_rows * 4, 256) * 4 + align_up(tmp, 256) + 1024;
}

// perm_out[i] = slot visited i-th.  `ws` holds row_order_workspace_bytes(n_rows).
int launch_row_order(const void* X, int dtype, int d, long long ld, long long q_begin, long long n_rows,
                     const double* centre, void* ws, size_t ws_bytes, int32_t** perm_out, cudaStream_t stream,
                     const int32_t* rows) {
  if (n_rows <= 0) { *perm_out = nullptr; return SBK_OK; }
  Arena a(ws, ws_bytes);
  uint32_t* keys_in = a.take<uint32_t>((size_t)n_rows);
  uint32_t* keys_out = a.take<uint32_t>((size_t)n_rows);
  int32_t* vals_in = a.take<int32_t>((size_t)n_rows);
  int32_t* vals_out = a.take<int32_t>((size_t)n_rows);
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys_in, keys_out, vals_in, vals_out, (int)n_rows, 0, 16, stream);
  void* cub_tmp = a.take<char>(tmp);
  if (!a.ok()) { set_error("row order: workspace %zu < %zu", ws_bytes, a.off); return SBK_ERR_WORKSPACE; }
  row_signature_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, stream>>>(X, dtype == SBK_F64 ? 1 : 0, d, ld, q_begin, n_rows, centre,
                                                                           keys_in, vals_in, rows);
  SBK_LAUNCH_CHECK();
  SBK_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(cub_tmp, tmp, keys_in, keys_out, vals_in, vals_out, (int)n_rows, 0, 16, stream));
  *perm_out = vals_out;
  return SBK_OK;
}

bool rerank_warp_supported(int dtype, int d, int k) {
  return rerank_warp_cap(k) != 0 && d <= (dtype == SBK_F32 ? RegQuery<float>::MAX_D : RegQuery<double>::MAX_D);
}

size_t rerank_smem(int d, int pmax, int ncand_cap) {
  return (size_t)((d + 1) & ~1) * 8 + (size_t)pmax * 12 + (size_t)ncand_cap * 8 + 16;
}

int launch_rerank(const void* X, int dtype, long long n_ref, int d, long long ld,
                  const int32_t* rows, long long q_begin, long long n_slots,
                  const uint32_t* cand_idx, const uint2* cand_pair, long long cand_stride, int n_seg,
                  const int32_t* cand_count, const double* bound, const double* nrm, const float4* norms4,
                  const double* scale, const int32_t* orig, int k,
                  void* out_dist, int32_t* out_idx, long long out_row0,
                  int32_t* fail_rows, int* n_fail, unsigned long long* totals,
                  cudaStream_t stream, const int32_t* slot_list, const int* n_list, int list_cap, const PeerOut* peers) {
  PeerOut po; po.n = 0;
  if (peers) po = *peers;
  if (cand_pair && !orig) { set_error("rerank: tensor-path candidates need the position -> row map"); return SBK_ERR_INVALID; }
  if (n_slots <= 0) return SBK_OK;
  // tensor path: all candidates staged (ncand_cap), survivors of the two-phase cut sorted in a
  // buffer of pmax entries; exact path: every candidate is a survivor
  int ncand_cap = cand_pair ? (int)cand_stride : 0;
  int pmax = (int)pow2ceil((uint32_t)cand_stride);
  if (cand_pair) pmax = std::min(pmax, std::max(1024, (int)pow2ceil((uint32_t)(2 * (k + 1)))));
  size_t smem = rerank_smem(d, pmax, ncand_cap);
  if (smem > 227 * 1024) { set_error("rerank: candidate capacity %lld too large", cand_stride); return SBK_ERR_INVALID; }
  if (n_seg < 1 || n_seg > 4 || cand_stride % n_seg) { set_error("rerank: bad segmentation"); return SBK_ERR_INVALID; }
  dim3 grid((unsigned)(slot_list ? std::min<long long>(n_slots, list_cap) : n_slots));
#define SBK_RR_LAUNCH(T, REGQ)                                                                                          \
  do {                                                                                                                  \
    SBK_CUDA_CHECK(cudaFuncSetAttribute(rerank_kernel<T, T, REGQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    rerank_kernel<T, T, REGQ><<<grid, RR_THREADS, smem, stream>>>(                                                      \
        (const T*)X, n_ref, d, ld, rows, q_begin, cand_idx, cand_pair, cand_stride, n_seg, cand_count, bound, nrm, norms4, scale, \
        orig, k, ncand_cap, pmax, (T*)out_dist, out_idx, out_row0, fail_rows, n_fail, totals, slot_list, n_list, list_cap, po);                               \
  } while (0)
  if (dtype == SBK_F32) {
    if (d <= RegQuery<float>::MAX_D) SBK_RR_LAUNCH(float, true); else SBK_RR_LAUNCH(float, false);
  } else {
    if (d <= RegQuery<double>::MAX_D) SBK_RR_LAUNCH(double, true); else SBK_RR_LAUNCH(double, false);
  }
#undef SBK_RR_LAUNCH
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}

}  // namespace sbk
