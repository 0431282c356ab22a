// integrate.cu -- the long-read ensemble integration loop on the GPU.
//
// Replaces SemiBin/long_read_cluster.py:12-49 (get_best_bin) and the extraction loop :122-143, which
// rebuild every bin of every DBSCAN result from Python lists for every extracted bin (O(bins * R * N)
// interpreter work plus list.index/pop per contig).  Here:
//   * per result r the contigs are sorted by label once (stable radix sort): a bin (r, label) is a
//     contiguous member list in ascending contig order;
//   * per bin: total length of its live members, marker count, distinct markers (OR of per-contig
//     marker bit masks), first live member, live count  (one CTA per bin);
//   * every extraction: arg-max over the bins that can still qualify (length >= minfasta and at least
//     one marker: both only ever shrink, so the list is built once) under the reference's order
//       smallest contamination step, then largest F1, then smallest length, then LAST in the
//       reference's iteration order (result index, position of the bin's first live contig)
//     (:16-47: a later bin replaces the incumbent when F1 is equal and its length is <=);
//   * the winner's live members are retired and only the bins they belonged to are re-reduced;
//   * the whole extraction loop (select -> retire -> re-reduce, :122-143) is ONE persistent cooperative kernel, one
//     CTA per SM, three grid barriers per extracted bin and no host round trip (as three launches per extraction the
//     loop was pure launch latency: 70 us per bin, twice the cost of the kNN graph and the DBSCAN sweep it follows).
// F1 / contamination are evaluated in FP64 with the reference's operation order and no contraction, so
// comparisons agree with the Python floats bit for bit; lengths and counts are integers.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cooperative_groups.h>
#include <algorithm>
#include <vector>
#include <string.h>

namespace sbk {

static constexpr int IG_MAX_R = 32;
static constexpr int IG_MAX_WORDS = 4;          // marker bit mask words per contig (<= 256 distinct markers)
static constexpr int IG_SEL_BLOCKS = 64;         // partial[] holds up to 4x this many CTAs' results

struct IgBin { long long weight; int total, distinct, first, count; };

struct IgStatus {                               // device-resident loop state
  int best;            // bin chosen by the last select: >= 0 bin id, -1 none (loop over), -2 "one contig left" (:124-126)
  int cur_label;       // extraction index of `best`
  int n_extracted;
  int done;
  long long remaining_weight;
  long long remaining_count;
  int n_dirty;
  unsigned int sel_ticket;
};

struct IgConst {
  int R; long long N; int words; int n_markers; long long minfasta;
  long long base[IG_MAX_R + 1];                 // first global bin id of each result
};

struct IgKey { int cls; double f1; long long w; int r; int first; int bin; };

__device__ __forceinline__ bool ig_better(const IgKey& a, const IgKey& b) {   // a strictly preferred over b
  if (b.bin < 0) return a.bin >= 0;
  if (a.bin < 0) return false;
  if (a.cls != b.cls) return a.cls < b.cls;
  if (a.f1 != b.f1) return a.f1 > b.f1;
  if (a.w != b.w) return a.w < b.w;
  if (a.r != b.r) return a.r > b.r;
  return a.first > b.first;
}

__device__ __forceinline__ int ig_result_of(const IgConst& c, long long g) {
  int r = 0;
  while (r + 1 < c.R && g >= c.base[r + 1]) ++r;
  return r;
}

__global__ void ig_keys_kernel(const long long* __restrict__ labels, long long n, int32_t* __restrict__ keys,
                               int32_t* __restrict__ vals, int* __restrict__ maxlab) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int m = -1;
  if (i < n) {
    const long long l = labels[i];
    keys[i] = l >= 0 ? (int32_t)l : 0x7fffffff;                    // noise sorts last
    vals[i] = (int32_t)i;
    m = l >= 0 ? (int)l : -1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m >= 0) atomicMax(maxlab, m);
}

__global__ void ig_bounds_kernel(const int32_t* __restrict__ skeys, long long n, long long base,
                                 int32_t* __restrict__ bstart, int32_t* __restrict__ bend) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int32_t k = skeys[p];
  if (k == 0x7fffffff) return;
  if (p == 0 || skeys[p - 1] != k) bstart[base + k] = (int32_t)p;
  if (p == n - 1 || skeys[p + 1] != k) bend[base + k] = (int32_t)(p + 1);
}

// One CTA per bin (all bins, or the dirty list).  Clears the bin's dirty flag.  Grid-stride over the items, so the
// same body serves the initial pass (one launch) and the re-reduction inside the persistent loop.
__device__ __forceinline__ void ig_stats_body(const IgConst& c, const int32_t* __restrict__ order, const int32_t* __restrict__ bstart,
                                              const int32_t* __restrict__ bend, const unsigned char* __restrict__ alive,
                                              const long long* __restrict__ lengths, const unsigned long long* __restrict__ masks,
                                              const int32_t* __restrict__ list, long long n_items,
                                              IgBin* __restrict__ bins, unsigned int* __restrict__ dirty,
                                              bool whole_list = false) {
  __shared__ long long s_w[8];
  __shared__ unsigned long long s_m[8][IG_MAX_WORDS];
  __shared__ int s_t[8], s_f[8], s_c[8];
  const int nwarp = blockDim.x >> 5;
  // whole_list: this CTA re-reduces every item of the (CTA-private) list; else the items are dealt round-robin to the CTAs
  for (long long it = whole_list ? 0 : blockIdx.x; it < n_items; it += whole_list ? 1 : gridDim.x) {
    const long long g = list ? (long long)list[it] : it;
    const int r = ig_result_of(c, g);
    const int32_t* mem = order + (long long)r * c.N;
    const int lo = bstart[g], hi = bend[g];
    long long w = 0; int tot = 0, first = 0x7fffffff, cnt = 0;
    unsigned long long m[IG_MAX_WORDS] = {0ull, 0ull, 0ull, 0ull};
    for (int p = lo + threadIdx.x; p < hi; p += blockDim.x) {
      const int i = mem[p];
      if (alive[i]) {
        w += lengths[i]; ++cnt; first = min(first, i);
        for (int u = 0; u < c.words; ++u) { const unsigned long long mi = masks[(long long)i * c.words + u]; m[u] |= mi; tot += __popcll(mi); }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      w += __shfl_xor_sync(0xffffffffu, w, o); tot += __shfl_xor_sync(0xffffffffu, tot, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o); first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
#pragma unroll
      for (int u = 0; u < IG_MAX_WORDS; ++u) m[u] |= __shfl_xor_sync(0xffffffffu, m[u], o);
    }
    const int wp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
      s_w[wp] = w; s_t[wp] = tot; s_f[wp] = first; s_c[wp] = cnt;
      for (int u = 0; u < IG_MAX_WORDS; ++u) s_m[wp][u] = m[u];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      IgBin b; b.weight = 0; b.total = 0; b.first = 0x7fffffff; b.count = 0;
      unsigned long long mm[IG_MAX_WORDS] = {0ull, 0ull, 0ull, 0ull};
      for (int q = 0; q < nwarp; ++q) {
        b.weight += s_w[q]; b.total += s_t[q]; b.first = min(b.first, s_f[q]); b.count += s_c[q];
        for (int u = 0; u < IG_MAX_WORDS; ++u) mm[u] |= s_m[q][u];
      }
      b.distinct = 0;
      for (int u = 0; u < IG_MAX_WORDS; ++u) b.distinct += __popcll(mm[u]);
      bins[g] = b;
      dirty[g] = 0u;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
ig_stats_kernel(IgConst c, const int32_t* __restrict__ order, const int32_t* __restrict__ bstart,
                const int32_t* __restrict__ bend, const unsigned char* __restrict__ alive,
                const long long* __restrict__ lengths, const unsigned long long* __restrict__ masks,
                long long n_bins_all, IgBin* __restrict__ bins, unsigned int* __restrict__ dirty) {
  ig_stats_body(c, order, bstart, bend, alive, lengths, masks, nullptr, n_bins_all, bins, dirty);
}

// Bins that can ever be chosen (long_read_cluster.py:26-34): both quantities only shrink as contigs retire.
__global__ void ig_eligible_kernel(const IgBin* __restrict__ bins, long long n_bins, long long minfasta,
                                   int32_t* __restrict__ elig, int* __restrict__ n_elig, int32_t* __restrict__ elig_pos) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_bins) return;
  const IgBin b = bins[g];
  int pos = -1;
  if (b.weight >= minfasta && b.total > 0) { pos = atomicAdd(n_elig, 1); elig[pos] = (int32_t)g; }
  elig_pos[g] = pos;                             // where the bin sits in the eligible list (-1: it can never be chosen)
}

__device__ __forceinline__ IgKey ig_key_of(const IgConst& c, const IgBin& b, int g) {
  IgKey k; k.bin = -1; k.cls = 9; k.f1 = 0.0; k.w = 0; k.r = 0; k.first = 0;
  if (b.weight < c.minfasta || b.total <= 0) return k;            // :27-28, :32-33
  // :34-40 in the reference's operation order, every operation rounded separately
  const double recall = __ddiv_rn((double)b.distinct, (double)c.n_markers);
  const double cont = __ddiv_rn((double)(b.total - b.distinct), (double)b.total);
  const double om = __dsub_rn(1.0, cont);
  k.f1 = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, recall), om), __dadd_rn(recall, om));
  k.cls = cont <= 0.1 ? 0 : cont <= 0.2 ? 1 : cont <= 0.3 ? 2 : cont <= 0.4 ? 3 : cont <= 0.5 ? 4 : 5;   // :16, :37
  k.w = b.weight; k.r = ig_result_of(c, g); k.first = b.first; k.bin = g;
  return k;
}

__device__ __forceinline__ IgKey ig_shfl_key(const IgKey& k, int o) {
  IgKey x;
  x.cls = __shfl_xor_sync(0xffffffffu, k.cls, o); x.f1 = __shfl_xor_sync(0xffffffffu, k.f1, o);
  x.w = __shfl_xor_sync(0xffffffffu, k.w, o); x.r = __shfl_xor_sync(0xffffffffu, k.r, o);
  x.first = __shfl_xor_sync(0xffffffffu, k.first, o); x.bin = __shfl_xor_sync(0xffffffffu, k.bin, o);
  return x;
}

static constexpr int IG_OWN_MAX = 256;    // touched bins of one extraction a CTA lists in shared memory

// Grid barrier of the persistent loop (all CTAs co-resident: cooperative launch, one CTA per SM).  One arrival counter
// and a generation word; against cooperative_groups' grid.sync() it took 9 ms off the 85 ms of the 500k-contig case (two
// barriers per extracted bin, 2336 bins).
__device__ __forceinline__ void ig_grid_barrier(unsigned int* bar, unsigned int n_blocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int gen;
    asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
    __threadfence();
    if (atomicAdd(bar, 1u) == n_blocks - 1u) {
      bar[0] = 0u;
      __threadfence();
      asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen + 1u) : "memory");
    } else {
      unsigned int now;
      do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(now) : "l"(bar + 1) : "memory"); } while (now == gen);
    }
  }
  __syncthreads();
}
// The extraction loop of long_read_cluster.py:122-143 as one persistent cooperative kernel (one CTA per SM).
// Every CTA OWNS a contiguous slice of the eligible-bin list: it keeps the best key of its slice, re-reduces the bins of
// its slice that the last extraction touched and rescans the slice only then (a few dozen bins change per extraction,
// most CTAs have nothing to do).  Per extracted bin, two grid barriers:
//   1. owners: re-reduce their touched bins, rescan their slice if anything changed -> partial[blockIdx]   | grid barrier
//   2. every CTA combines ALL partials itself (the decision and the loop state are replicated), evaluates the loop
//      conditions of :123-134, retires the winner's live members and lists the bins they belong to (:136-141);
//      the lists of consecutive extractions alternate between two buffers                                     | grid barrier
__global__ void __launch_bounds__(256)
ig_loop_kernel(IgConst c, const long long* __restrict__ labels, const int32_t* __restrict__ order,
               const int32_t* __restrict__ bstart, const int32_t* __restrict__ bend,
               unsigned char* __restrict__ alive, const long long* __restrict__ lengths,
               const unsigned long long* __restrict__ masks, IgBin* __restrict__ bins, unsigned int* __restrict__ dirty,
               int32_t* __restrict__ dirty_list, long long dirty_cap, const int32_t* __restrict__ elig, const int* __restrict__ n_elig,
               const int32_t* __restrict__ elig_pos, IgKey* __restrict__ partial, IgStatus* __restrict__ status,
               int* __restrict__ n_dirty2, long long* __restrict__ out_labels, long long max_iter) {
  unsigned int* gbar = reinterpret_cast<unsigned int*>(n_dirty2 + 2);      // [arrivals, generation], zeroed by the host
  __shared__ IgKey s_k[8];
  __shared__ int s_best;
  __shared__ int s_own[IG_OWN_MAX];
  __shared__ int s_nown;
  // loop state, replicated in every thread of every CTA (identical by construction)
  long long remaining_weight = status->remaining_weight, remaining_count = status->remaining_count;
  int n_extracted = 0;
  const int n = *n_elig;
  const int per = (n + (int)gridDim.x - 1) / (int)gridDim.x;
  const int e_lo = min(n, (int)blockIdx.x * per), e_hi = min(n, e_lo + per);
  bool done = false;
  for (long long iter = 0; iter < max_iter && !done; ++iter) {
    const int cur = (int)(iter & 1);                       // dirty-list buffer this extraction fills
    // ---- 1. owners: re-reduce what the previous extraction touched, rescan the slice if anything changed ----
    bool rescan = iter == 0;
    if (iter > 0) {
      // which of the touched bins are mine: one list entry per thread (independent loads), owned ones into shared memory
      const int32_t* dl = dirty_list + (long long)(cur ^ 1) * dirty_cap;
      const int nd = n_dirty2[cur ^ 1];
      if (threadIdx.x == 0) s_nown = 0;
      __syncthreads();
      for (int it = threadIdx.x; it < nd; it += blockDim.x) {
        const int g = dl[it];
        const int e = elig_pos[g];
        if (e >= e_lo && e < e_hi) { const int at = atomicAdd(&s_nown, 1); if (at < IG_OWN_MAX) s_own[at] = g; }
      }
      __syncthreads();
      const int nown = s_nown;
      if (nown > IG_OWN_MAX) {                             // more than the shared list holds (never seen): walk the global list
        for (int it = 0; it < nd; ++it) {
          const int g = dl[it];
          const int e = elig_pos[g];
          if (e < e_lo || e >= e_hi) continue;             // block-uniform
          ig_stats_body(c, order, bstart, bend, alive, lengths, masks, dl + it, 1, bins, dirty, true);
        }
      } else if (nown > 0) {
        ig_stats_body(c, order, bstart, bend, alive, lengths, masks, s_own, (long long)nown, bins, dirty, true);
      }
      rescan = nown > 0;
    }
    if (rescan) {                                          // block-uniform
      IgKey best; best.bin = -1; best.cls = 9; best.f1 = 0; best.w = 0; best.r = 0; best.first = 0;
      for (int e = e_lo + threadIdx.x; e < e_hi; e += blockDim.x) {
        const int g = elig[e];
        const IgKey kk = ig_key_of(c, bins[g], g);
        if (ig_better(kk, best)) best = kk;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const IgKey x = ig_shfl_key(best, o); if (ig_better(x, best)) best = x; }
      if ((threadIdx.x & 31) == 0) s_k[threadIdx.x >> 5] = best;
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int q = 1; q < (int)(blockDim.x >> 5); ++q) if (ig_better(s_k[q], best)) best = s_k[q];
        partial[blockIdx.x] = best;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) n_dirty2[cur] = 0;     // its last readers passed two barriers ago
    ig_grid_barrier(gbar, gridDim.x);
    // ---- 2. decision (replicated: one warp per CTA combines the partials), retire the winner ----
    if (threadIdx.x < 32) {
      IgKey b; b.bin = -1; b.cls = 9; b.f1 = 0; b.w = 0; b.r = 0; b.first = 0;
      for (int q = threadIdx.x; q < (int)gridDim.x; q += 32) { const IgKey x = partial[q]; if (ig_better(x, b)) b = x; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const IgKey x = ig_shfl_key(b, o); if (ig_better(x, b)) b = x; }
      if (threadIdx.x == 0) {
        int choice;
        if (remaining_weight < c.minfasta) choice = -1;                                    // :123
        else if (remaining_count == 1) choice = -2;                                        // :124-126
        else if (b.bin < 0) choice = -1;                                                   // :133-134
        else choice = b.bin;
        s_best = choice;
      }
    }
    __syncthreads();
    const int choice = s_best;
    if (choice == -1) { done = true; break; }
    const int cur_label = n_extracted;
    n_extracted += 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (choice == -2) {                                      // the single remaining contig becomes a bin of its own
      for (long long i = t0; i < c.N; i += stride)
        if (alive[i]) { alive[i] = 0; out_labels[i] = cur_label; }
      done = true;
      break;
    }
    remaining_weight -= bins[choice].weight;                 // bins[] is not written before the next barrier
    remaining_count -= bins[choice].count;
    {
      const int r = ig_result_of(c, choice);
      const int32_t* mem = order + (long long)r * c.N;
      const int lo = bstart[choice], hi = bend[choice];
      int32_t* dl = dirty_list + (long long)cur * dirty_cap;
      for (long long p = lo + t0; p < hi; p += stride) {
        const int i = mem[p];
        if (!alive[i]) continue;
        alive[i] = 0;
        out_labels[i] = cur_label;
        // the member's bins in all results, eight results at a time: the label, position and flag loads of a batch are
        // independent of each other (one result after the other this chain of dependent loads was most of an extraction)
        for (int q0 = 0; q0 < c.R; q0 += 8) {
          long long gg[8]; int ep[8]; unsigned int fl[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int q = q0 + u;
            const long long l = q < c.R ? labels[(long long)q * c.N + i] : -1;
            gg[u] = l >= 0 ? c.base[q < c.R ? q : 0] + l : -1;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) ep[u] = gg[u] >= 0 ? elig_pos[gg[u]] : -1;      // bins that can never be chosen are not listed
#pragma unroll
          for (int u = 0; u < 8; ++u) fl[u] = ep[u] >= 0 ? dirty[gg[u]] : 1u;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (fl[u] == 0u && atomicExch(&dirty[gg[u]], 1u) == 0u) dl[atomicAdd(&n_dirty2[cur], 1)] = (int32_t)gg[u];
        }
      }
    }
    ig_grid_barrier(gbar, gridDim.x);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { status->n_extracted = n_extracted; status->done = done ? 1 : 0; }
}


struct IgWs {
  int32_t* order; int32_t* keys_in; int32_t* keys_out; int32_t* vals_in; int* maxlab;
  int32_t* bstart; int32_t* bend; IgBin* bins; unsigned int* dirty; int32_t* dirty_list; int32_t* elig; int* n_elig;
  unsigned char* alive; IgKey* partial; IgStatus* status; void* cub_tmp; size_t cub_bytes;
  int32_t* elig_pos; int* n_dirty2; long long dirty_cap;
};

static size_t ig_cub_bytes(long long n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)n, 0, 32, (cudaStream_t)0);
  return bytes;
}

static void ig_carve(Arena& a, long long n, int R, IgWs& w) {
  const size_t max_bins = (size_t)n * R;
  w.order = a.take<int32_t>((size_t)n * R);
  w.keys_in = a.take<int32_t>((size_t)n); w.keys_out = a.take<int32_t>((size_t)n); w.vals_in = a.take<int32_t>((size_t)n);
  w.maxlab = a.take<int>(IG_MAX_R);
  w.bstart = a.take<int32_t>(max_bins); w.bend = a.take<int32_t>(max_bins);
  w.bins = a.take<IgBin>(max_bins);
  w.dirty = a.take<unsigned int>(max_bins);
  w.dirty_cap = (long long)max_bins;                       // two alternating lists (consecutive extractions)
  w.dirty_list = a.take<int32_t>(2 * max_bins);
  w.elig_pos = a.take<int32_t>(max_bins); w.n_dirty2 = a.take<int>(4);
  w.elig = a.take<int32_t>(max_bins); w.n_elig = a.take<int>(4);
  w.alive = a.take<unsigned char>((size_t)n);
  w.partial = a.take<IgKey>(IG_SEL_BLOCKS * 4);
  w.status = a.take<IgStatus>(2);
  w.cub_bytes = ig_cub_bytes(n);
  w.cub_tmp = a.take<char>(w.cub_bytes);
}

}  // namespace sbk

using namespace sbk;

extern "C" size_t sbk_integrate_workspace_bytes(int64_t n, int32_t n_results) {
  if (n < 1 || n >= (1ll << 31) - 1 || n_results < 1 || n_results > IG_MAX_R) return 0;
  Arena a(nullptr, 0);
  IgWs w;
  ig_carve(a, n, n_results, w);
  return a.off + 1024;
}

extern "C" int sbk_integrate(const long long* labels, int32_t n_results, int64_t n, const long long* lengths,
                             const unsigned long long* marker_masks, int32_t mask_words, int32_t n_markers,
                             int64_t minfasta, long long* out_labels, int32_t* n_extracted,
                             void* workspace, size_t ws_bytes, void* cuda_stream) {
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  if (sbk_device_count() < 1) { set_error("integrate: no CUDA device (there is no CPU fallback)"); return SBK_ERR_CUDA; }
  if (n < 1 || n >= (1ll << 31) - 1) { set_error("integrate: n=%lld outside [1, 2^31)", (long long)n); return SBK_ERR_INVALID; }
  if (n_results < 1 || n_results > IG_MAX_R) { set_error("integrate: %d clustering results (1..%d supported)", n_results, IG_MAX_R); return SBK_ERR_INVALID; }
  if (mask_words < 1 || mask_words > IG_MAX_WORDS) { set_error("integrate: %d marker mask words (1..%d supported)", mask_words, IG_MAX_WORDS); return SBK_ERR_INVALID; }
  if (n_markers < 1) { set_error("integrate: n_markers must be positive"); return SBK_ERR_INVALID; }
  Arena a(workspace, ws_bytes);
  IgWs w;
  ig_carve(a, n, n_results, w);
  if (!a.ok() || workspace == nullptr) { set_error("integrate: workspace %zu B < required %zu B", ws_bytes, a.off); return SBK_ERR_WORKSPACE; }

  IgConst c; memset(&c, 0, sizeof(c));
  c.R = n_results; c.N = n; c.words = mask_words; c.n_markers = n_markers; c.minfasta = minfasta;
  const unsigned nb = (unsigned)((n + 255) / 256);
  // ---- members of every bin: contigs sorted by label, per result ----
  SBK_CUDA_CHECK(cudaMemsetAsync(w.maxlab, 0xff, IG_MAX_R * sizeof(int), stream));      // -1
  std::vector<int> maxlab(n_results, -1);
  for (int r = 0; r < n_results; ++r) {
    ig_keys_kernel<<<nb, 256, 0, stream>>>(labels + (long long)r * n, n, w.keys_in, w.vals_in, w.maxlab + r);
    SBK_LAUNCH_CHECK();
    size_t bytes = w.cub_bytes;
    SBK_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, bytes, w.keys_in, w.keys_out, w.vals_in,
                                                   w.order + (long long)r * n, (int)n, 0, 32, stream));
    // bin boundaries need the global bin base of this result = number of labels of the earlier ones
    SBK_CUDA_CHECK(cudaMemcpyAsync(&maxlab[r], w.maxlab + r, sizeof(int), cudaMemcpyDeviceToHost, stream));
    SBK_CUDA_CHECK(cudaStreamSynchronize(stream));
    if (maxlab[r] >= n) { set_error("integrate: label %d >= n in result %d", maxlab[r], r); return SBK_ERR_INVALID; }
    c.base[r + 1] = c.base[r] + (long long)(maxlab[r] + 1);
    if (r == 0) {
      // start/end of labels that never occur stay empty ranges
      SBK_CUDA_CHECK(cudaMemsetAsync(w.bstart, 0, (size_t)n * n_results * sizeof(int32_t), stream));
      SBK_CUDA_CHECK(cudaMemsetAsync(w.bend, 0, (size_t)n * n_results * sizeof(int32_t), stream));
    }
    ig_bounds_kernel<<<nb, 256, 0, stream>>>(w.keys_out, n, c.base[r], w.bstart, w.bend);
    SBK_LAUNCH_CHECK();
  }
  const long long n_bins = c.base[n_results];
  // ---- initial state ----
  SBK_CUDA_CHECK(cudaMemsetAsync(w.alive, 1, (size_t)n, stream));
  SBK_CUDA_CHECK(cudaMemsetAsync(out_labels, 0xff, (size_t)n * sizeof(long long), stream));           // -1
  SBK_CUDA_CHECK(cudaMemsetAsync(w.n_elig, 0, 4 * sizeof(int), stream));
  SBK_CUDA_CHECK(cudaMemsetAsync(w.n_dirty2, 0, 4 * sizeof(int), stream));
  if (n_bins > 0) SBK_CUDA_CHECK(cudaMemsetAsync(w.dirty, 0, (size_t)n_bins * sizeof(unsigned int), stream));
  // total length of all contigs (loop condition :123): reuse the reduction of a one-bin view is overkill -> host sum
  std::vector<long long> hl((size_t)n);
  SBK_CUDA_CHECK(cudaMemcpyAsync(hl.data(), lengths, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, stream));
  SBK_CUDA_CHECK(cudaStreamSynchronize(stream));
  IgStatus st; memset(&st, 0, sizeof(st));
  st.best = -1;
  for (long long i = 0; i < n; ++i) st.remaining_weight += hl[i];
  st.remaining_count = n;
  SBK_CUDA_CHECK(cudaMemcpyAsync(w.status, &st, sizeof(st), cudaMemcpyHostToDevice, stream));
  if (n_bins > 0) {
    const unsigned sg = (unsigned)std::min<long long>(n_bins, 148 * 64);
    ig_stats_kernel<<<sg, 256, 0, stream>>>(c, w.order, w.bstart, w.bend, w.alive, lengths, marker_masks, n_bins, w.bins, w.dirty);
    SBK_LAUNCH_CHECK();
    ig_eligible_kernel<<<(unsigned)((n_bins + 255) / 256), 256, 0, stream>>>(w.bins, n_bins, minfasta, w.elig, w.n_elig, w.elig_pos);
    SBK_LAUNCH_CHECK();
  }
  // ---- extraction loop: one persistent cooperative kernel, one CTA per SM ----
  {
    int dev = 0, sms = 0, occ = 0;
    SBK_CUDA_CHECK(cudaGetDevice(&dev));
    SBK_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SBK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ig_loop_kernel, 256, 0));
    if (occ < 1) { set_error("integrate: the loop kernel does not fit on an SM"); return SBK_ERR_INTERNAL; }
    const int grid = std::min(sms, IG_SEL_BLOCKS * 4);
    long long max_iter = n + 2;
    const long long* labels_p = labels; const long long* lengths_p = lengths; const unsigned long long* masks_p = marker_masks;
    void* args[] = {(void*)&c, (void*)&labels_p, (void*)&w.order, (void*)&w.bstart, (void*)&w.bend, (void*)&w.alive,
                    (void*)&lengths_p, (void*)&masks_p, (void*)&w.bins, (void*)&w.dirty, (void*)&w.dirty_list, (void*)&w.dirty_cap,
                    (void*)&w.elig, (void*)&w.n_elig, (void*)&w.elig_pos, (void*)&w.partial, (void*)&w.status, (void*)&w.n_dirty2,
                    (void*)&out_labels, (void*)&max_iter};
    SBK_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)ig_loop_kernel, dim3(grid), dim3(256), args, 0, stream));
  }
  SBK_CUDA_CHECK(cudaMemcpyAsync(&st, w.status, sizeof(st), cudaMemcpyDeviceToHost, stream));
  SBK_CUDA_CHECK(cudaStreamSynchronize(stream));
  if (!st.done) { set_error("integrate: extraction loop did not terminate"); return SBK_ERR_INTERNAL; }
  if (n_extracted) *n_extracted = st.n_extracted;
  return SBK_OK;
}
