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

// FP64 direct-difference squared distance between a query held in shared memory
// (as double) and row j of X.
template <typename T>
__device__ __forceinline__ double sqdist_row(const double* __restrict__ q, const T* __restrict__ y, int d);

template <>
__device__ __forceinline__ double sqdist_row<float>(const double* __restrict__ q,
                                                    const float* __restrict__ y, int d) {
  // four independent accumulators (column c goes to accumulator c mod 4) so the FMA
  // latency chain is 4x shorter; the final combination order is fixed -> deterministic
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int c = 0;
  if ((((uintptr_t)y) & 15) == 0) {
    const float4* y4 = reinterpret_cast<const float4*>(y);
#pragma unroll 4
    for (; c + 4 <= d; c += 4) {
      float4 v = __ldg(y4 + (c >> 2));
      double t0 = q[c] - (double)v.x, t1 = q[c + 1] - (double)v.y;
      double t2 = q[c + 2] - (double)v.z, t3 = q[c + 3] - (double)v.w;
      a0 = fma(t0, t0, a0); a1 = fma(t1, t1, a1);
      a2 = fma(t2, t2, a2); a3 = fma(t3, t3, a3);
    }
  }
  for (; c < d; ++c) { double t = q[c] - (double)__ldg(y + c); a0 = fma(t, t, a0); }
  return (a0 + a1) + (a2 + a3);
}

template <>
__device__ __forceinline__ double sqdist_row<double>(const double* __restrict__ q,
                                                     const double* __restrict__ y, int d) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int c = 0;
  if ((((uintptr_t)y) & 15) == 0) {
    const double2* y2 = reinterpret_cast<const double2*>(y);
#pragma unroll 4
    for (; c + 4 <= d; c += 4) {
      double2 v = __ldg(y2 + (c >> 1)), w = __ldg(y2 + (c >> 1) + 1);
      double t0 = q[c] - v.x, t1 = q[c + 1] - v.y, t2 = q[c + 2] - w.x, t3 = q[c + 3] - w.y;
      a0 = fma(t0, t0, a0); a1 = fma(t1, t1, a1);
      a2 = fma(t2, t2, a2); a3 = fma(t3, t3, a3);
    }
  }
  for (; c < d; ++c) { double t = q[c] - __ldg(y + c); a0 = fma(t, t, a0); }
  return (a0 + a1) + (a2 + a3);
}

}  // namespace sbk
