// common.cuh -- shared helpers for libsbk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/sbk.h"

namespace sbk {

void set_error(const char* fmt, ...);

#define SBK_CUDA_CHECK(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      sbk::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SBK_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define SBK_LAUNCH_CHECK() SBK_CUDA_CHECK(cudaGetLastError())

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
  char* base; size_t cap; size_t off;
  Arena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
  template <class T> T* take(size_t count) {
    off = align_up(off, 256);
    T* r = (T*)(base ? base + off : nullptr);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};

__host__ __device__ static inline uint32_t pow2ceil(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ---------------------------------------------------------------------------
// (key, id) lexicographic order and a block-wide bitonic sort over shared memory
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool kv_less(double ka, uint32_t ia, double kb, uint32_t ib) {
  return (ka < kb) || (ka == kb && ia < ib);
}

// Sorts key[0..P) / id[0..P) ascending by (key, id).  P is a power of two.
// All `nthreads` threads of the group must call; `sync` is __syncthreads-like.
template <int NTHREADS>
__device__ __forceinline__ void block_bitonic_sort(double* key, uint32_t* id, int P, int tid) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = tid; t < (P >> 1); t += NTHREADS) {
        int lo = 2 * t - (t & (stride - 1));      // index with bit `stride` clear
        int hi = lo + stride;
        bool up = ((lo & size) == 0);
        double ka = key[lo], kb = key[hi];
        uint32_t ia = id[lo], ib = id[hi];
        bool swap = up ? kv_less(kb, ib, ka, ia) : kv_less(ka, ia, kb, ib);
        if (swap) { key[lo] = kb; key[hi] = ka; id[lo] = ib; id[hi] = ia; }
      }
    }
  }
  __syncthreads();
}

// FP64 direct-difference squared distance between a query held in shared memory (as double)
// and a row y of X, evaluated by a group of 8 consecutive lanes: lane `sub` (0..7) of the group
// takes the 16-byte chunks sub, sub+8, ... of the row (one coalesced 128-byte request per
// group and step) and the partial sums are combined by a fixed xor-shuffle tree, so the value
// does not depend on the launch shape or on the alignment of the row (the unaligned path reads
// the same columns with scalar loads).  All 32 lanes of the warp must call (shuffles); every
// lane of a group returns the group's sum.
template <typename T>
__device__ __forceinline__ double sqdist_g8(const double* __restrict__ q, const T* __restrict__ y, int d, int sub) {
  constexpr int V = 16 / (int)sizeof(T);             // columns per 16-byte chunk
  const int nfull = d / V;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if ((((uintptr_t)y) & 15) == 0) {
#pragma unroll 4
    for (int ch = sub; ch < nfull; ch += 8) {
      const int c = ch * V;
      if (sizeof(T) == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(y) + ch);
        const double t0 = q[c] - (double)v.x, t1 = q[c + 1] - (double)v.y;
        const double t2 = q[c + 2] - (double)v.z, t3 = q[c + 3] - (double)v.w;
        a0 = fma(t0, t0, a0); a1 = fma(t1, t1, a1); a2 = fma(t2, t2, a2); a3 = fma(t3, t3, a3);
      } else {
        const double2 v = __ldg(reinterpret_cast<const double2*>(y) + ch);
        const double t0 = q[c] - v.x, t1 = q[c + 1] - v.y;
        if ((ch >> 3) & 1) { a2 = fma(t0, t0, a2); a3 = fma(t1, t1, a3); }
        else { a0 = fma(t0, t0, a0); a1 = fma(t1, t1, a1); }
      }
    }
  } else {
    for (int ch = sub; ch < nfull; ch += 8) {
      const int c = ch * V;
      if (sizeof(T) == 4) {
        const double t0 = q[c] - (double)__ldg(y + c), t1 = q[c + 1] - (double)__ldg(y + c + 1);
        const double t2 = q[c + 2] - (double)__ldg(y + c + 2), t3 = q[c + 3] - (double)__ldg(y + c + 3);
        a0 = fma(t0, t0, a0); a1 = fma(t1, t1, a1); a2 = fma(t2, t2, a2); a3 = fma(t3, t3, a3);
      } else {
        const double t0 = q[c] - (double)__ldg(y + c), t1 = q[c + 1] - (double)__ldg(y + c + 1);
        if ((ch >> 3) & 1) { a2 = fma(t0, t0, a2); a3 = fma(t1, t1, a3); }
        else { a0 = fma(t0, t0, a0); a1 = fma(t1, t1, a1); }
      }
    }
  }
  // tail columns (d not a multiple of the chunk width): column nfull*V + t goes to lane t
  const int ct = nfull * V + sub;
  if (sub < V && ct < d) { const double t = q[ct] - (double)__ldg(y + ct); a0 = fma(t, t, a0); }
  double s = (a0 + a1) + (a2 + a3);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  return s;
}

// Register-resident query row for the common narrow case: a group of LPC consecutive lanes evaluates one
// candidate row, lane `sub` owning the 16-byte chunks sub, sub+LPC, ... (NCH of them at most) of every row.
// The lane's slice of the query lives in registers as double, so the inner loop is one 16-byte global
// load per chunk and FP64 arithmetic -- no shared-memory traffic (the load/store pipe was the busiest unit
// of the rerank when the query sat in shared memory).  Fixed xor-shuffle tree -> deterministic.
template <typename T> struct DistCfg;
template <> struct DistCfg<float>  { static constexpr int LPC = 8,  NCH = 4, V = 4; };   // d <= 128
template <> struct DistCfg<double> { static constexpr int LPC = 16, NCH = 5, V = 2; };   // d <= 160

template <typename T>
struct RegQuery {
  using C = DistCfg<T>;
  static constexpr int MAX_D = C::LPC * C::NCH * C::V;
  double q[C::NCH * C::V];
  double qt;
  int nfull, d;
  bool vec;

  __device__ __forceinline__ void load(const T* __restrict__ x, int d_, int sub, bool vec_ok) {
    d = d_; nfull = d_ / C::V; vec = vec_ok;
#pragma unroll
    for (int j = 0; j < C::NCH; ++j) {
      const int ch = sub + C::LPC * j;
#pragma unroll
      for (int e = 0; e < C::V; ++e) q[j * C::V + e] = ch < nfull ? (double)x[ch * C::V + e] : 0.0;
    }
    const int ct = nfull * C::V + sub;
    qt = (sub < C::V && ct < d_) ? (double)x[ct] : 0.0;
  }

  // all 32 lanes call; every lane of a group returns the group's squared distance
  __device__ __forceinline__ double eval(const T* __restrict__ y, int sub) const {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (sizeof(T) == 4) {
      float4 v[C::NCH];
#pragma unroll
      for (int j = 0; j < C::NCH; ++j) {
        const int ch = sub + C::LPC * j;
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ch < nfull) {
          if (vec) v[j] = __ldg(reinterpret_cast<const float4*>(y) + ch);
          else v[j] = make_float4((float)__ldg(y + ch * 4), (float)__ldg(y + ch * 4 + 1), (float)__ldg(y + ch * 4 + 2), (float)__ldg(y + ch * 4 + 3));
        }
      }
#pragma unroll
      for (int j = 0; j < C::NCH; ++j) {      // inactive chunks hold zeros on both sides
        const double t0 = q[4 * j] - (double)v[j].x, t1 = q[4 * j + 1] - (double)v[j].y;
        const double t2 = q[4 * j + 2] - (double)v[j].z, t3 = q[4 * j + 3] - (double)v[j].w;
        a0 = fma(t0, t0, a0); a1 = fma(t1, t1, a1); a2 = fma(t2, t2, a2); a3 = fma(t3, t3, a3);
      }
    } else {
      double2 v[C::NCH];
#pragma unroll
      for (int j = 0; j < C::NCH; ++j) {
        const int ch = sub + C::LPC * j;
        v[j] = make_double2(0.0, 0.0);
        if (ch < nfull) {
          if (vec) v[j] = __ldg(reinterpret_cast<const double2*>(y) + ch);
          else v[j] = make_double2((double)__ldg(y + ch * 2), (double)__ldg(y + ch * 2 + 1));
        }
      }
#pragma unroll
      for (int j = 0; j < C::NCH; ++j) {
        const double t0 = q[2 * j] - v[j].x, t1 = q[2 * j + 1] - v[j].y;
        if (j & 1) { a2 = fma(t0, t0, a2); a3 = fma(t1, t1, a3); }
        else { a0 = fma(t0, t0, a0); a1 = fma(t1, t1, a1); }
      }
    }
    const int ct = nfull * C::V + sub;
    if (sub < C::V && ct < d) { const double t = qt - (double)__ldg(y + ct); a0 = fma(t, t, a0); }
    double s = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = C::LPC / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
  }
};

}  // namespace sbk
