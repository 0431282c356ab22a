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

}  // namespace sbk
