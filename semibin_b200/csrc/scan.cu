#include "scan.cuh"

namespace sbk {

__global__ void __launch_bounds__(SCAN_BLOCK)
scan_local_kernel(const int32_t* __restrict__ cnt, long long n, long long* __restrict__ off,
                  long long* __restrict__ block_sums) {
  __shared__ long long warp_tot[32];
  const long long i = (long long)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long v = (i < n) ? (long long)cnt[i] : 0;
  long long x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[warp] = x;
  __syncthreads();
  if (warp == 0) {
    long long t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  long long incl = x + (warp ? warp_tot[warp - 1] : 0);
  if (i < n) off[i] = incl - v;                 // exclusive, block-local
  if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = incl;
}

__global__ void scan_blocks_kernel(long long* __restrict__ block_sums, long long nb) {
  // single thread block, sequential over chunks of 1024 (nb is small)
  __shared__ long long carry;
  __shared__ long long warp_tot[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long base = 0; base < nb; base += 1024) {
    long long i = base + threadIdx.x;
    long long v = (i < nb) ? block_sums[i] : 0;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      long long t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    long long incl = x + (warp ? warp_tot[warp - 1] : 0) + carry;
    if (i < nb) block_sums[i] = incl - v;       // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sums[nb] = carry; // grand total
}

__global__ void __launch_bounds__(SCAN_BLOCK)
scan_add_kernel(long long* __restrict__ off, long long n, const long long* __restrict__ block_sums,
                long long nb) {
  const long long i = (long long)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  if (i < n) off[i] += block_sums[blockIdx.x];
  if (i == 0) off[n] = block_sums[nb];
}

size_t scan_workspace_elems(long long n) { return (size_t)((n + SCAN_BLOCK - 1) / SCAN_BLOCK + 2); }

int exclusive_scan(const int32_t* cnt, long long n, long long* off, long long* block_sums,
                   cudaStream_t stream) {
  if (n <= 0) {
    SBK_CUDA_CHECK(cudaMemsetAsync(off, 0, sizeof(long long), stream));
    return SBK_OK;
  }
  long long nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  scan_local_kernel<<<(unsigned)nb, SCAN_BLOCK, 0, stream>>>(cnt, n, off, block_sums);
  SBK_LAUNCH_CHECK();
  scan_blocks_kernel<<<1, 1024, 0, stream>>>(block_sums, nb);
  SBK_LAUNCH_CHECK();
  scan_add_kernel<<<(unsigned)nb, SCAN_BLOCK, 0, stream>>>(off, n, block_sums, nb);
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}

}  // namespace sbk
