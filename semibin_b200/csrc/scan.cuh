// scan.cuh -- exclusive prefix sum of int32 counts into int64 offsets
// (three small kernels; inputs are at most a few million rows).
#pragma once
#include "common.cuh"

namespace sbk {
static constexpr int SCAN_BLOCK = 1024;
// off[0..n] (n+1 entries) = exclusive scan of cnt[0..n); block_sums needs
// ceil(n/SCAN_BLOCK)+1 int64 entries.
size_t scan_workspace_elems(long long n);
int exclusive_scan(const int32_t* cnt, long long n, long long* off, long long* block_sums,
                   cudaStream_t stream);
}  // namespace sbk
