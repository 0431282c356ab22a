// knn_tc.cu -- tcgen05 (5th-gen tensor core) candidate filter of the kNN graph.
//
// Replaces the 256x256-chunk dgemm + heap loop of sklearn's EuclideanArgKmin
// (_middle_term_computer.pyx.tp:401-442, _argkmin.pyx.tp:490-510) reached from
// SemiBin/cluster.py:94-105 and long_read_cluster.py:108-113.
//
// Data flow (all per GPU, references replicated):
//   tc_stage_kernel   X (f32/f64) -> centred, power-of-two scaled FP16 operands in
//                     core-matrix tiled layout + per-row norms / rounding-error norms
//   tc_kernel<SAMPLE> every query tile x a pseudo-random reference sample:
//                     per-row group minima -> order-statistic threshold thr_i and
//                     the certificate bound_i
//   tc_kernel<FILTER> every query tile x ALL reference tiles: emit column j as a
//                     candidate iff s_ij <= thr_i
// where   s_ij = ||z_j||^2 - 2 <zh_i, zh_j>   is produced DIRECTLY by the MMA: the
// three spare K columns carry an FP16 hi/mid/lo split of ||z_j||^2 against ones.
// ||z_i||^2 is constant per row and irrelevant for the ranking.
//
// Exactness: two more K columns make the MMA subtract a rigorous PER-PAIR error bound
// E_ij (bilinear in per-row norms, see tc_pair_bound in knn.cuh), so the accumulator
// holds a LOWER bound  L_ij <= S^2 d2_ij - ||z_i||^2 + sigma_i  of the exact score.
// Every reference that is NOT emitted (L_ij > thr_i) has d2 > bound_i; the FP64 rerank
// (knn_exact.cu) certifies  d2_(k+1) < bound_i  per row or sends the row to the
// exact filter.  The neighbour sets are therefore exact, never approximate.
//
// Kernel anatomy: a CTA PAIR (cluster of 2, the two SMs of a TPC) per 256-query tile, 1 CTA / SM,
// 20 warps per CTA.  tcgen05.mma.cta_group::2 multiplies the pair's 256 query rows (128 per CTA)
// with a 256-reference tile of which each CTA stages only its 128-row half: per SM the operand
// traffic from L2 is 32 B/clk at the MMA floor instead of the 64 B/clk (= the SM's L2 port) of a
// single-CTA M=128 tile, which is what held the single-CTA version at 61 % tensor-pipe activity.
//   warp 0     : bulk-async (TMA engine, UBLKCP) producer: own A tile once, then a ring of
//                128-row reference half tiles, mbarrier complete_tx
//   warp 1     : leader CTA: single-thread tcgen05.mma issuer, M=256 N=256 K=16, FP16 in / FP32
//                accumulate in TMEM (128 lanes x 256 columns in each CTA), two accumulator stages;
//                completion is multicast to both CTAs (tcgen05.commit ... multicast::cluster)
//                peer CTA: relay -- forwards "my half tile has landed" to the leader's barrier
//   warp 2     : TMEM allocator
//   warps 4-19 : epilogue (4 warps per TMEM lane quadrant, 64 tile columns each):
//                tcgen05.ld 32x32b.x32, FMNMX3 min-tree over groups of 8 against
//                the per-row threshold, rare-path append to the thread's own
//                candidate segment (register counter, no atomics)
#include "common.cuh"
#include "knn.cuh"
#include <math.h>
#include <algorithm>
#include <vector>
#include <string>
#include <time.h>
#include <type_traits>
#include <string.h>
#include <stdlib.h>

namespace sbk {

// Warp roles.  The SMSP arbiter prefers the HIGHEST warp id among eligible warps (B300_MICROARCH.md), so the
// single-thread producer and MMA issuer sit above the 16 epilogue warps: they must never queue behind
// epilogue instructions, or the tensor pipe idles between tiles.
static constexpr int TC_WARP_TMA = 16, TC_WARP_MMA = 17, TC_WARP_TMEM = 18;
static constexpr int TC_EPI_WARPS = 16;         // 4 per TMEM lane quadrant, 64 of the 256 tile columns each
static constexpr int TC_THREADS = 32 * (TC_EPI_WARPS + 4);
static constexpr int TC_NGROUP = 128;           // group minima per row (4 column quarters x 32 columns)
static constexpr int TC_MAX_STAGES = 6;
static constexpr int TC_DEFAULT_CG = 2;
static constexpr int TC_DEFAULT_NH = 1;         // accumulator stages per tile (2 = four half-tile stages); see tc_kernel
static constexpr int TC_MAX_KPAD = 176;
static constexpr int TC_MAX_SPLIT = 12;         // columns that may get an FP16 hi+lo representation
// instruction descriptor, kind::f16: A,B = F16 (0), D = F32 (1<<4), K-major, N>>3 @17, M>>4 @24 (built in the kernel)

static constexpr long long TC_L2_CHUNK_BYTES = 32ll << 20;   // reference operand bytes kept hot in L2 per filter launch
enum TcMode { TC_SAMPLE = 0, TC_FILTER = 1, TC_DUMP = 2,
              // profiling aids (SBK_TC_PROBE=3|4 replaces the filter pass; results are then meaningless):
              TC_PROBE_NOLD = 3,      // epilogue releases the accumulator without reading it: MMA + operand pipeline alone
              TC_PROBE_LDONLY = 4,    // epilogue reads the accumulator (tcgen05.ld) and discards it
              TC_PROBE_NOAPPEND = 5,  // filter epilogue without the candidate append (min tree + threshold test only)
              TC_FILTER_SYM = 6 };    // symmetric self-join filter: row test AND column test on every tile (see tc_kernel)
// symmetric filter: ring of per-tile column thresholds, one slot = 256 column thresholds t_j + their 32 group-of-8 maxima
static constexpr int TC_T8_SLOTS = 10;          // > nstage + 3 (see tc_kernel)
static constexpr int TC_TT_FLOATS = TC_BN / 8;                  // floats per tile: the maxima of its 32 groups of 8 column thresholds (128 B)
// pool of group records (symmetric filter): chunks of 512 uint4 = 8 KB: [0] = (record count, ..), then 170 records of
// three uint4: (row position, first column position, -, -), the row's 8 accumulator values of the group
static constexpr int TC_POOL_CH = 170;
static constexpr int TC_POOL_STRIDE = 512;                      // uint4 per chunk

// bytes of the reference-tile ring; at least the SAMPLE epilogue's scratch
// ([TC_BM][TC_NGROUP + 1] floats) which reuses it once the pipeline has drained
__host__ __device__ static inline size_t tc_ring_bytes(int kpad, int nstage, int cg) {
  const size_t ring = (size_t)nstage * (TC_BN / cg) * kpad * 2;
  const size_t scratch = (size_t)(TC_NGROUP + 1) * TC_BM * 4;
  return ring > scratch ? ring : scratch;
}

// device scalars produced by the centre/selection kernel
struct TcScalars {
  int n_split;                     // number of split ("precision") columns chosen for this matrix
  int split_cols[TC_MAX_SPLIT];
};

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(100000u) : "memory");   // suspend-time hint (ns)
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {      // non-blocking
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU.  The loop lives in one asm
// block (try_wait, branch, counter): a waiting warp is woken by every barrier event of the CTA, so the
// instructions per wake-up are what the waiters cost the working warps in issue slots.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // four try_waits per turn of the counter: a wake-up costs four instructions instead of seven.  When the counter
  // runs out the wall clock decides: four seconds without progress is a protocol bug -> trap.
  unsigned long long t0 = 0;
  for (;;) {
    uint32_t timed_out;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " .reg .u32 n;\n"
        " mov.u32 n, 0;\n"
        "WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        " @p bra DONE_%=;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        " @p bra DONE_%=;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        " @p bra DONE_%=;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        " @p bra DONE_%=;\n"
        " add.u32 n, n, 1;\n"
        " setp.lt.u32 p, n, 0x1000;\n"
        " @p bra WAIT_%=;\n"
        " mov.u32 %0, 1;\n"
        " bra END_%=;\n"
        "DONE_%=:\n"
        " mov.u32 %0, 0;\n"
        "END_%=:\n"
        "}"
        : "=r"(timed_out) : "r"(smem_u32(bar)), "r"(parity), "r"(100000u) : "memory");
    if (!timed_out) return;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0) t0 = now;
    else if (now - t0 > 4000000000ull) {
      printf("sbk: mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// Pure polling wait (no suspend) for the single-warp roles on the tensor pipe's critical path.
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  unsigned long long t0 = 0;
  for (;;) {
#pragma unroll 1
    for (int n = 0; n < 0x10000; ++n) if (mbar_test(bar, parity)) return;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0) t0 = now;
    else if (now - t0 > 4000000000ull) { printf("sbk: mbarrier spin timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n .reg .b32 ra;\n mapa.shared::cluster.u32 ra, %0, %1;\n mbarrier.arrive.shared::cluster.b64 _, [ra];\n}"
      ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ bool elect_one() {      // one lane of the (converged) warp
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major, no swizzle ("interleave") shared-memory matrix descriptor:
//   core matrix = 8 rows x 16 B stored contiguously (128 B)
//   LBO = byte stride between core matrices adjacent in K
//   SBO = byte stride between core matrices adjacent in M/N
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;                               // descriptor version (sm_100)
  return d;                                      // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------
// staging
// ---------------------------------------------------------------------------
static constexpr int TC_COLSTAT_ROWS = 128;
template <typename T>
__global__ void __launch_bounds__(256)
tc_colstat_kernel(const T* __restrict__ X, long long n, int d, long long ld,
                  double* __restrict__ colsum, double* __restrict__ colsumsq,
                  unsigned long long* __restrict__ maxabs_bits) {
  // block handles TC_COLSTAT_ROWS rows; thread t accumulates columns t, t+256, ...; eight rows per step so that
  // eight independent loads are in flight per thread (with one the kernel ran at 0.85 TB/s)
  const long long r0 = (long long)blockIdx.x * TC_COLSTAT_ROWS;
  const long long r1 = min(n, r0 + TC_COLSTAT_ROWS);
  double mx = 0.0;
  for (int c = threadIdx.x; c < d; c += 256) {
    double s = 0.0, s2 = 0.0;
    long long r = r0;
    for (; r + 8 <= r1; r += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (double)X[(r + u) * ld + c];
#pragma unroll
      for (int u = 0; u < 8; ++u) { s += v[u]; s2 = fma(v[u], v[u], s2); mx = fmax(mx, fabs(v[u])); }
    }
    for (; r < r1; ++r) { double v = (double)X[r * ld + c]; s += v; s2 = fma(v, v, s2); mx = fmax(mx, fabs(v)); }
    atomicAdd(&colsum[c], s);
    atomicAdd(&colsumsq[c], s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(maxabs_bits, (unsigned long long)__double_as_longlong(mx));
}

// Single block: centre, scale, and the choice of "precision" columns.  A column whose
// energy (sum of squared centred values) is far above the average dominates the FP16
// rounding error of every score; such columns get an FP16 hi+lo representation in two
// extra K columns each (cross terms zl_i*zh_j + zh_i*zl_j), which removes their share
// of the error bound.  Homogeneous matrices (the embedding) select none.
__global__ void tc_centre_kernel(const double* __restrict__ colsum, const double* __restrict__ colsumsq,
                                 const unsigned long long* __restrict__ maxabs_bits,
                                 long long n, int d, int max_split, double* __restrict__ centre,
                                 double* __restrict__ scale, TcScalars* __restrict__ sc) {
  __shared__ double cmax;
  __shared__ double energy[4096];
  __shared__ double total;
  __shared__ int n_sel;
  if (threadIdx.x == 0) { cmax = 0.0; total = 0.0; n_sel = 0; }
  __syncthreads();
  double m = 0.0, esum = 0.0;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    double cv = colsum[c] / (double)n;
    cv = (double)(float)cv;                       // keep the centre float-representable
    centre[c] = cv;
    m = fmax(m, fabs(cv));
    // sum (x - cv)^2 = sumsq - 2 cv sum + n cv^2
    const double e = fmax(0.0, colsumsq[c] - 2.0 * cv * colsum[c] + (double)n * cv * cv);
    energy[c] = e;
    esum += e;
  }
  atomicMax((unsigned long long*)&cmax, (unsigned long long)__double_as_longlong(m));
  atomicAdd(&total, esum);
  __syncthreads();
  const double avg = total / (double)d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const double e = energy[c];
    if (e > 4.0 * avg) {
      int rank = 0;
      for (int o = 0; o < d; ++o) rank += (energy[o] > e || (energy[o] == e && o < c)) ? 1 : 0;
      if (rank < max_split) { sc->split_cols[rank] = c; atomicAdd(&n_sel, 1); }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    sc->n_split = n_sel;                          // ranks 0..n_sel-1 are exactly the selected columns
    double bound = __longlong_as_double((long long)*maxabs_bits) + cmax;   // >= max |x - c|
    if (!(bound > 0.0)) bound = 1.0;
    // power of two S with S * bound in [4, 8): FP16 elements stay far from overflow and
    // ||z||^2 <= 64 d < 65504 for d <= 1023
    int e;
    frexp(bound, &e);                             // bound = f * 2^e, f in [0.5,1)
    scale[0] = ldexp(1.0, 3 - e);                 // S * bound = f * 8 in [4, 8)
  }
}

// One thread per staged row.
//   b_tiles : reference-side operand, tiles of `brows` rows:  -2*zh | n_hi n_mid n_lo | (-2*zh_c, -2*zl_c) per split column | 0
//   a_tiles : query-side operand, tiles of TC_BM rows:        zh  |  1    1     1   | (  zl_c ,   zh_c ) per split column | 0
// tile layout (FP16 elements): [kpad/8 chunks][R rows][8]  -> core matrices of 8 rows x 16 B
// The REFERENCE side is stored in a pseudo-random order: row i goes to position p with i = (p * mult + off) mod n
// (golden-ratio stride, a bijection since gcd(mult, n) = 1).  Any prefix of the staged references is then an
// unbiased sample of all of them, whatever the order the caller's rows came in: the first `sample_size`
// positions ARE the threshold sample, and the candidates of the first reference chunks refine the thresholds
// for the later ones (tc_refine_kernel).  orig[p] = i maps a candidate position back to its row.
template <typename T>
__global__ void __launch_bounds__(128)
tc_stage_kernel(const T* __restrict__ X, long long n, int d, long long ld, long long n_staged,
                long long perm_minv, long long perm_off,
                const double* __restrict__ centre, const double* __restrict__ scale, int kpad, int brows,
                __half* __restrict__ b_tiles, __half* __restrict__ a_tiles, int32_t* __restrict__ orig,
                double* __restrict__ nrm, float4* __restrict__ norms4,
                const TcScalars* __restrict__ sc, int a_by_pos, int32_t* __restrict__ rpos, float2* __restrict__ cu) {
  const long long row = (long long)blockIdx.x * 128 + threadIdx.x;
  if (row >= n_staged) return;
  const bool real = row < n;
  // position of this row on the reference side (padding rows keep their place behind the n real ones)
  long long pos = row;
  if (real) {
    long long t = row - perm_off; if (t < 0) t += n;
    pos = (long long)(((unsigned long long)t * (unsigned long long)perm_minv) % (unsigned long long)n);
  }
  orig[pos] = real ? (int32_t)row : 0x7fffffff;
  const long long bt = pos / brows, br = pos % brows;       // brows = reference rows per staged tile (256, or 128 for CTA pairs)
  __half* bdst = b_tiles + bt * (long long)brows * kpad + br * 8;
  // query-side tiles: in row order, or (symmetric self-join) in the same position order as the reference side
  const long long apos = a_by_pos ? pos : row;
  const long long at = apos / TC_BM, ar = apos % TC_BM;
  __half* adst = a_tiles + at * (long long)TC_BM * kpad + ar * 8;
  const long long src = row;
  const double S = scale[0];
  double n2 = 0.0, d2 = 0.0, zh2 = 0.0;
  const int nchunk = kpad / 8;
  for (int ch = 0; ch < nchunk; ++ch) {
    __align__(16) __half hb[8];
    __align__(16) __half ha[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = ch * 8 + e;
      float zf = 0.f;
      if (real && c < d) {
        const double z = S * ((double)X[src * ld + c] - centre[c]);
        const __half h = __double2half(z);
        zf = __half2float(h);
        const double dl = z - (double)zf;
        n2 += z * z; d2 += dl * dl; zh2 += (double)zf * (double)zf;
      }
      ha[e] = __float2half_rn(zf);
      hb[e] = __float2half_rn(-2.0f * zf);
    }
    // norm columns d, d+1, d+2 (filled after the loop for the chunks that hold them)
    *reinterpret_cast<uint4*>(bdst + (long long)ch * brows * 8) = *reinterpret_cast<uint4*>(hb);
    if (adst) *reinterpret_cast<uint4*>(adst + (long long)ch * TC_BM * 8) = *reinterpret_cast<uint4*>(ha);
  }
  // split columns: hi+lo representation, two K columns each
  double l2 = 0.0;
  const int n_split = sc->n_split;
  for (int t = 0; t < n_split; ++t) {
    const int c = sc->split_cols[t];
    float zhf = 0.f, zlf = 0.f;
    if (real) {
      const double z = S * ((double)X[src * ld + c] - centre[c]);
      zhf = __half2float(__double2half(z));
      const double r = z - (double)zhf;
      zlf = __half2float(__double2half(r));
      const double rr = r - (double)zlf;
      d2 += rr * rr - r * r;                                     // residual of this column is now rr
      zh2 += ((double)zhf + (double)zlf) * ((double)zhf + (double)zlf) - (double)zhf * (double)zhf;
      l2 += (double)zlf * (double)zlf;
    }
    const int k0 = d + 5 + 2 * t, k1c = k0 + 1;
    bdst[(long long)(k0 >> 3) * brows * 8 + (k0 & 7)] = __float2half_rn(-2.0f * zhf);     // x A: zl_i
    bdst[(long long)(k1c >> 3) * brows * 8 + (k1c & 7)] = __float2half_rn(-2.0f * zlf);   // x A: zh_i
    if (adst) {
      adst[(long long)(k0 >> 3) * TC_BM * 8 + (k0 & 7)] = __float2half_rn(zlf);
      adst[(long long)(k1c >> 3) * TC_BM * 8 + (k1c & 7)] = __float2half_rn(zhf);
    }
  }
  if (d2 < 0.0) d2 = 0.0;
  // per-row norms, rounded UP (they only enter error bounds); the 1e-4 inflation covers the
  // accumulation error of the bound columns themselves
  const double c_acc = tc_c_acc(kpad);
  const float nz = __double2float_ru(sqrt(n2) * 1.0001);
  const float nd = __double2float_ru(sqrt(d2) * 1.0001);
  const float na = __double2float_ru(sqrt(zh2) * 1.0001);
  const float l2f = __double2float_ru(l2 * 1.0001);
  // norm columns d..d+2: FP16 hi/mid/lo split of  t_j = ||z_j||^2 (1 - c_acc) - ||l_j||^2  with the last
  // piece rounded DOWN, so pieces <= t_j  (padding rows: a huge value so they never pass a threshold)
  double tt = real ? (n2 * (1.0 - c_acc) - (double)l2f - 1e-12 * n2) : 60000.0;
  const __half h0 = __double2half(tt);
  const double r0 = tt - (double)__half2float(h0);
  const __half h1 = __double2half(r0);
  const double r1 = r0 - (double)__half2float(h1);
  const __half h2 = __float2half_rd(__double2float_rd(r1));
  const __half hn[3] = {h0, h1, h2};
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    const int c = d + e;
    bdst[(long long)(c >> 3) * brows * 8 + (c & 7)] = hn[e];
    if (adst) adst[(long long)(c >> 3) * TC_BM * 8 + (c & 7)] = __float2half_rn(real ? 1.0f : 0.0f);
  }
  // bound columns d+3, d+4:  A (u_i, v_i) = (||delta_i||, ||a_i||),  B = -(2||z_j||, 2||delta_j|| + 2 c_acc ||a_j||),
  // every factor rounded up to FP16  ->  the MMA subtracts >= the bilinear part of E_ij
  {
    const float pj = real ? 2.0f * nz : 0.f;
    const float qj = real ? __fadd_ru(2.0f * nd, __double2float_ru(2.0 * c_acc * (double)na)) : 0.f;
    const int c0 = d + 3, c1 = d + 4;
    bdst[(long long)(c0 >> 3) * brows * 8 + (c0 & 7)] = __hneg(__float2half_ru(pj));
    bdst[(long long)(c1 >> 3) * brows * 8 + (c1 & 7)] = __hneg(__float2half_ru(qj));
    if (adst) {
      adst[(long long)(c0 >> 3) * TC_BM * 8 + (c0 & 7)] = __float2half_ru(real ? nd : 0.f);
      adst[(long long)(c1 >> 3) * TC_BM * 8 + (c1 & 7)] = __float2half_ru(real ? na : 0.f);
    }
  }
  if (real) {
    nrm[row] = n2;
    norms4[row] = make_float4(nz, nd, na, l2f);
    if (rpos) rpos[row] = (int32_t)pos;
  }
  // (c, u) = ||z||^2 - sigma rounded down / up: L_ij + c_i <= S^2 d2_ij, and a lower bound of S^2 d2 minus u_j is a
  // lower bound in row j's frame (e_ji + sigma_j).  Padding positions: c = +inf (never a column hit), u = 0.
  if (cu) cu[pos] = real ? make_float2(__double2float_rd((n2 - (double)l2f) - 1e-12 * n2), __double2float_ru((n2 - (double)l2f) + 1e-12 * n2))
                         : make_float2(INFINITY, 0.f);
}

// Threshold and certificate bound of one query row from tau, an upper bound (with the planned confidence) of the
// row's k+1 smallest LOWER bounds.  Their exact scores may be up to the per-pair slack higher: estimate that
// slack for references at the threshold distance (norms of such neighbours from the triangle inequality); a
// poor estimate only costs a fallback, the certificate is what guarantees exactness.
__device__ __forceinline__ void tc_threshold(float tau, const float4 ni, double ni2, int kpad, int d, double S,
                                             float* thr_out, double* bound_out) {
  const double c_acc = tc_c_acc(kpad);
  const double rho = sqrt(fmax(0.0, (double)tau + ni2));
  const double zj = (double)ni.x + rho;
  const double ratio = fmax(ni.x > 0.f ? (double)ni.y / (double)ni.x : 0.0, 1.220703125e-4);
  float4 nj; nj.x = (float)zj; nj.y = (float)(ratio * zj + 6e-8 * sqrt((double)d)); nj.z = (float)zj; nj.w = ni.w;
  const double margin = 1.05 * tc_pair_upper_slack(ni, nj, c_acc);
  const float thr_f = __double2float_ru((double)tau + margin);
  *thr_out = thr_f;
  // every non-emitted j has L_ij > thr  =>  S^2 d2_ij - n_i = e_ij >= L_ij - ||l_i||^2 > thr - ||l_i||^2
  *bound_out = ((double)thr_f - (double)ni.w + ni2) / (S * S) * (1.0 - 1e-13);
}

__device__ __forceinline__ uint32_t tc_f2ord(float f) {          // order-preserving float -> uint32
  const uint32_t b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float tc_ord2f(uint32_t u) {
  return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu));
}

// Threshold refinement between filter launches.  After the first launches the candidate lists hold EVERY
// reference with L <= thr among the first `fraction` of the (pseudo-randomly ordered) references: a sample
// several times the size of the threshold sample.  The r-th smallest lower bound seen so far, r from the
// binomial tail of "how many of the row's k nearest fall into the part seen", bounds the k+1 smallest lower
// bounds of the row with the planned confidence: the threshold of the remaining launches drops to it (plus the
// usual slack), which roughly halves the candidates they emit.  One warp per query row; rows with more than
// 256 candidates so far, an overflowed segment or fewer than r candidates keep their threshold.
static constexpr int TC_REFINE_MAX = 512;
__global__ void __launch_bounds__(256)
tc_refine_kernel(const uint2* __restrict__ cand, const int32_t* __restrict__ cand_count, int cand_cap,
                 long long row_lo, long long n_rows, int r_stat,
                 const float4* __restrict__ norms4, const double* __restrict__ nrm, const double* __restrict__ scale,
                 int kpad, int d, float* __restrict__ thr, double* __restrict__ bound, const int32_t* __restrict__ slot_row) {
  const long long slot = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (slot >= n_rows) return;
  const int seg_cap = cand_cap >> 2;
  int c[4], n = 0; bool bad = false;
#pragma unroll
  for (int sgm = 0; sgm < 4; ++sgm) { c[sgm] = cand_count[slot * 4 + sgm]; bad |= c[sgm] > seg_cap; n += c[sgm]; }
  if (bad || n < r_stat || n > TC_REFINE_MAX) return;
  uint32_t key[TC_REFINE_MAX / 32];
  const uint2* base = cand + slot * (long long)cand_cap;
#pragma unroll
  for (int u = 0; u < TC_REFINE_MAX / 32; ++u) {
    int e = lane + 32 * u;
    key[u] = 0xffffffffu;
    if (e < n) {
      int sgm = 0;
      while (e >= c[sgm]) { e -= c[sgm]; ++sgm; }           // n <= sum of the four counts
      key[u] = tc_f2ord(__uint_as_float(base[(long long)sgm * seg_cap + e].y));
    }
  }
  // r-th smallest key = the largest v with #(keys < v) < r, built bit by bit -- only over the bits in which the row's
  // keys differ at all (they share their leading bits: 32 -> ~20 dependent steps) and only over the key registers the
  // row fills (lists of 100-300 entries after the first launches, not 512)
  uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
  for (int u = 0; u < TC_REFINE_MAX / 32; ++u) { kmin = min(kmin, key[u]); if (lane + 32 * u < n) kmax = max(kmax, key[u]); }
  kmin = __reduce_min_sync(0xffffffffu, kmin); kmax = __reduce_max_sync(0xffffffffu, kmax);
  const int hb = 31 - __clz(kmin ^ kmax);                    // -1: all keys equal
  uint32_t res = hb >= 31 ? 0u : (kmin & ~((2u << hb) - 1u));
  if (hb < 0) res = kmin;
  auto select = [&](auto nk_tag) {
    constexpr int NK = decltype(nk_tag)::value;
    for (int bit = hb; bit >= 0; --bit) {
      const uint32_t t = res | (1u << bit);
      int cnt = 0;
#pragma unroll
      for (int u = 0; u < NK; ++u) cnt += key[u] < t ? 1 : 0;
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (cnt < r_stat) res = t;
    }
  };
  if (n <= 128) select(std::integral_constant<int, 4>());
  else if (n <= 256) select(std::integral_constant<int, 8>());
  else select(std::integral_constant<int, TC_REFINE_MAX / 32>());
  if (lane == 0) {
    const long long grow = slot_row ? (long long)slot_row[slot] : row_lo + slot;
    float thr_new; double bound_new;
    tc_threshold(tc_ord2f(res), norms4[grow], nrm[grow], kpad, d, scale[0], &thr_new, &bound_new);
    if (thr_new < thr[slot]) { thr[slot] = thr_new; bound[slot] = bound_new; }
  }
}

// ---- symmetric self-join filter: helpers ----
// Column thresholds t_p = thr_p + u_p (rounded up) >= thr_p - sigma_p + ||z_p||^2: t8[g] = the largest of group g of 8
// positions (padding positions: -inf), and tq[p] = (c_p, u_p, t_p, -) for the transpose kernel.  Rerun after every
// threshold refinement.
__global__ void __launch_bounds__(256)
tc_colthr_kernel(const float* __restrict__ thr, const float2* __restrict__ cu, long long n, float* __restrict__ t8,
                 float4* __restrict__ tq) {
  const long long pp = (long long)blockIdx.x * TC_BN + threadIdx.x;
  float m = -INFINITY;
  const float2 c = cu[pp];
  if (pp < n) m = __fadd_ru(thr[pp], c.y);
  tq[pp] = make_float4(c.x, c.y, m, 0.f);
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 7) == 0) t8[(long long)blockIdx.x * TC_TT_FLOATS + (threadIdx.x >> 3)] = m;
}

// Symmetric filter: the 256 positions of every tile are re-ordered by descending column threshold (known after the
// sample pass), so that the 8 columns of a group share nearly the same threshold and the group pre-test of the filter
// epilogue stays tight whatever the spread of neighbourhood sizes in the data.  Tile membership is unchanged (the
// sample is still the first tiles, windows of tiles are still unbiased); both operand tiles and every per-position
// array are permuted in place through shared memory.  One CTA of 256 threads per tile.
__global__ void __launch_bounds__(256)
tc_sort_tiles_kernel(long long n, int kpad, __half* __restrict__ a_tiles, __half* __restrict__ b_tiles,
                     float* __restrict__ thr, double* __restrict__ bound, float2* __restrict__ cu,
                     int32_t* __restrict__ orig, int32_t* __restrict__ rpos) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  __shared__ float s_key[TC_BN];
  __shared__ int s_src[TC_BN];                     // s_src[r] = old local index of the position that moves to rank r
  const int tid = threadIdx.x;
  const long long base = (long long)blockIdx.x * TC_BN;
  const long long pp = base + tid;
  const float my_thr = thr[pp];                    // thr / bound / cu / orig are allocated for the padded tiles
  const double my_bound = bound[pp];
  const float2 my_cu = cu[pp];
  const int32_t my_orig = orig[pp];
  const float key = pp < n ? __fadd_ru(my_thr, my_cu.y) : -INFINITY;
  s_key[tid] = key;
  __syncthreads();
  // rank by (key descending, index ascending): 256 comparisons per thread
  int rank = 0;
  for (int o = 0; o < TC_BN; ++o) { const float ko = s_key[o]; rank += (ko > key || (ko == key && o < tid)) ? 1 : 0; }
  // Groups of 8 consecutive ranks stay together, but the groups are dealt round-robin to the four 64-column quarters
  // of the tile: a row's candidates (rows of similar threshold) keep spreading evenly over its four list segments,
  // which belong to the quarters.  The last, partly padded tile keeps the plain order (its real positions must stay
  // in front of the padding).
  if (base + TC_BN <= n) { const int gi = rank >> 3; rank = (gi & 3) * 64 + (gi >> 2) * 8 + (rank & 7); }
  s_src[rank] = tid;
  thr[base + rank] = my_thr; bound[base + rank] = my_bound; cu[base + rank] = my_cu; orig[base + rank] = my_orig;
  if (my_orig != 0x7fffffff) rpos[my_orig] = (int32_t)(base + rank);
  __syncthreads();
  // operand tiles: two half tiles of 128 rows, layout [kpad/8 chunks][128 rows][8 halves]; 16-byte pieces
  const int nchunk = kpad / 8;
  uint4* buf = reinterpret_cast<uint4*>(sort_smem);             // [nchunk][256] pieces of the whole tile
  for (int side = 0; side < 2; ++side) {
    uint4* g = reinterpret_cast<uint4*>((side ? b_tiles : a_tiles) + base * kpad);
    for (int e = tid; e < nchunk * TC_BN; e += 256) {
      const int half = e / (nchunk * 128), rem = e % (nchunk * 128);          // global piece order: [half][chunk][row]
      const int ch = rem / 128, row = rem % 128;
      buf[ch * TC_BN + half * 128 + row] = g[e];
    }
    __syncthreads();
    for (int e = tid; e < nchunk * TC_BN; e += 256) {
      const int half = e / (nchunk * 128), rem = e % (nchunk * 128);
      const int ch = rem / 128, row = rem % 128;
      g[e] = buf[ch * TC_BN + s_src[half * 128 + row]];
    }
    __syncthreads();
  }
}

// Symmetric filter, after every launch: the GROUP records of the launch sit in pool chunks, one stream per epilogue
// warp.  A record (i, j0, v[8]) holds the accumulator values L_ij of row position i for the 8 column positions j0..j0+7,
// written because the group's minimum passed the pre-test.  Here every element is resolved for both of its rows, with
// s = L_ij + c_i <= S^2 d2_ij:
//   row i receives candidate (j, s - u_i) iff s <= t_i   (a lower bound in i's frame: e_ij + sigma_i);
//   row j receives candidate (i, s - u_j) iff s <= t_j and the pair's tile was column-tested (not the diagonal tile, nor
//   the opposite tile of an even tile count, which hold both orders of their pairs themselves).
// t = thr + u >= thr - sigma + ||z||^2, so whatever is not delivered to a row has d2 > its bound.  Candidates are
// appended to one of the row's four segments with an atomic counter; a count beyond the segment capacity is the
// overflow mark the rerank already understands.  One warp per chunk, one record per lane and turn (eight lanes per
// record with coalesced reads measured 15 % slower: the kernel is bound by the scattered 8-byte candidate stores).
template <bool MAIL>
__global__ void __launch_bounds__(256, MAIL ? 3 : 6)
tc_transpose_kernel(const uint4* __restrict__ pool, int* __restrict__ pool_count, int pool_chunks,
                    long long n, long long n_tiles, long long blk_tiles, const float4* __restrict__ tq,
                    uint2* __restrict__ cand, int32_t* __restrict__ cand_count, int cand_cap,
                    unsigned long long* __restrict__ dbg, const TcMail mail) {
  const int lane = threadIdx.x & 31;
  const int n_chunks = min(*pool_count, pool_chunks);
  const int seg_cap = cand_cap >> 2;
  // a fixed grid of warps takes the chunks handed out one by one (their number is only known on the device, and they
  // are unevenly filled: a static split left a tail)
  for (;;) {
  int ch = 0;
  if (lane == 0) ch = atomicAdd(pool_count + 1, 1);
  ch = __shfl_sync(0xffffffffu, ch, 0);
  if (ch >= n_chunks) break;
  const uint4* src = pool + (long long)ch * TC_POOL_STRIDE;
  const int cnt = min((int)src[0].x, TC_POOL_CH);
  if (dbg && lane == 0) atomicAdd(&dbg[2], (unsigned long long)cnt);
  // every turn takes 32 records, one per lane (all lanes stay in the loop: the mailbox part below is warp-collective)
  for (int e0 = 0; e0 < cnt; e0 += 32) {
    const int e = e0 + lane;
    const bool valid = e < cnt;
    long long i = 0, j0 = 0;
    bool remote = false;
    unsigned colm = 0;                                           // columns of a remote owner that passed their test
    float sv[8];
    if (valid) {
      const uint4 hd = src[1 + 3 * e], q0 = src[2 + 3 * e], q1 = src[3 + 3 * e];
      i = hd.x; j0 = hd.y;
      const float v[8] = {__uint_as_float(q0.x), __uint_as_float(q0.y), __uint_as_float(q0.z), __uint_as_float(q0.w),
                          __uint_as_float(q1.x), __uint_as_float(q1.y), __uint_as_float(q1.z), __uint_as_float(q1.w)};
      const float4 qi = tq[i];                                   // (c_i, u_i, t_i)
      // pairs of one tile block are multiplied circulantly inside the block: its diagonal tiles, and the opposite tiles of
      // an even block length, hold both orders of their pairs themselves and were row-tested only
      const long long ti = i / TC_BN, tj = j0 / TC_BN;
      bool col_on = true;
      if (ti / blk_tiles == tj / blk_tiles) {
        const long long len = min(blk_tiles, n_tiles - (ti / blk_tiles) * blk_tiles);
        long long dt = tj - ti; if (dt < 0) dt = -dt;
        col_on = !(dt == 0 || dt * 2 == len);
      }
      // the eight columns of a record are consecutive positions of one tile: one owner
      remote = MAIL && !(j0 >= mail.own_lo && j0 < mail.own_hi);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        sv[c] = __fadd_rd(v[c], qi.x);
        const long long j = j0 + c;
        const int ts = (e + c) & 3;
        if (sv[c] <= qi.z) {
          const int at = atomicAdd(&cand_count[i * 4 + ts], 1);
          if (at < seg_cap) cand[i * (long long)cand_cap + (long long)ts * seg_cap + at] = make_uint2((uint32_t)j, __float_as_uint(__fsub_rd(sv[c], qi.y)));
          else if (dbg && at == seg_cap) atomicAdd(&dbg[1], 1ull);
        }
        if (col_on) {
          const float4 qj = tq[j];                               // padding positions: t = -inf
          if (sv[c] <= qj.z) {
            if (MAIL && remote) colm |= 1u << c;
            else {
              const int at = atomicAdd(&cand_count[j * 4 + ts], 1);
              if (at < seg_cap) cand[j * (long long)cand_cap + (long long)ts * seg_cap + at] = make_uint2((uint32_t)i, __float_as_uint(__fsub_rd(sv[c], qj.y)));
              else if (dbg && at == seg_cap) atomicAdd(&dbg[1], 1ull);
            }
          }
        }
      }
    }
    if (MAIL) {
      // columns that live on another GPU: into the outbox of their owner.  The lanes that send to the same rank take their
      // slots with ONE atomic per turn (single increments of one address: 30 ms for the 54 M records of 1M on two GPUs)
      const int rcnt = remote ? __popc(colm) : 0;
      const int dst = (int)(j0 / mail.blk_pos);
      unsigned pending = __ballot_sync(0xffffffffu, rcnt > 0);
      while (pending) {
        const int lead = __ffs(pending) - 1;
        const int d0 = __shfl_sync(0xffffffffu, dst, lead);
        const bool mine = rcnt > 0 && dst == d0;
        const unsigned grp = __ballot_sync(0xffffffffu, mine);
        const int mv = mine ? rcnt : 0;
        int incl = mv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int base = 0;
        if (lane == lead) base = atomicAdd(&mail.out_cnt[d0], total);
        base = __shfl_sync(0xffffffffu, base, lead);
        if (mine) {
          int at = base + incl - mv;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (colm >> c & 1u) {
              if (at < mail.cap) mail.out[d0][at] = make_uint4((uint32_t)(j0 + c), (uint32_t)i, __float_as_uint(sv[c]), 0u);
              ++at;
            }
          }
        }
        pending &= ~grp;
      }
    }
  }
  }
}

// ---------------------------------------------------------------------------
// the tcgen05 kernel
// ---------------------------------------------------------------------------
struct TcArgs {
  const __half* a_tiles;      // query operand tiles (TC_BM rows each), all n rows
  const __half* b_tiles;      // reference operand tiles (TC_BN / CG rows each, CG per MMA tile): full set or sample
  long long n_btiles;         // reference tiles walked by this launch ...
  long long bt0;              // ... starting at this tile (FILTER: the references are walked in L2-sized chunks)
  long long tile_m0;          // first query tile of this launch (global tile index)
  long long row_lo, row_hi;   // valid global query rows [row_lo,row_hi); slot = row - row_lo
  int kpad, nstage;
  // SAMPLE
  int order_stat;             // r (1-based)
  const float4* norms4; const double* nrm; const TcScalars* sc; const double* scale; int d;
  float* thr;                 // [slots]
  double* bound;              // [slots]
  // FILTER
  uint2* cand; int32_t* cand_count; int cand_cap;   // (column, score) pairs: 4 segments of cand_cap/4 per row
  // DUMP
  float* dump; long long dump_ld; const int32_t* orig;
  // slots -> rows (NULL: slot s of the launch is row row_lo + s); set when the query side is staged in position order
  const int32_t* slot_row;
  // FILTER_SYM: bt0 / n_btiles are the first circulant step and the number of steps of this launch; the CTA pair of
  // query tile I of the tile block [blk_lo, blk_lo + blk_len) multiplies reference tiles blk_lo + (I - blk_lo + step) mod
  // blk_len (circulant inside the block; the whole matrix is the block [0, T)), or -- sym_linear -- every pair of the
  // launch walks the same tiles bt0, bt0 + 1, ... (a block of query tiles against a block of other reference tiles)
  long long blk_lo, blk_len; int sym_linear;
  const float* t8;            // [tiles * 32] group maxima of the column thresholds
  const float2* cu;           // [positions] (c, u)
  uint4* pool; int* pool_count; int pool_chunks; int* pool_fail;    // group-record pool (chunks of TC_POOL_STRIDE uint4)
};

template <int CG> __device__ __forceinline__ void tc_commit(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else   // arrives on the barrier at this offset in BOTH CTAs of the pair once all prior MMAs have retired
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <int CG> __device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                             uint32_t accumulate) {
  if (CG == 1)
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// CG = 1: one CTA per 128-query tile, M=128 N=256 MMAs, the CTA stages whole 256-row reference tiles.
// CG = 2: a CTA pair (cluster of 2) per 256 query rows, tcgen05.mma.cta_group::2 M=256 N=256; each CTA
//         stages its own A tile and its 128-row half of every reference tile, the leader issues the MMAs
//         and multicasts their completion, the peer's MMA warp relays "my operands have landed".
// NH = accumulator stages per reference tile: 1 -> two stages of 256 TMEM columns (one N=256 MMA group per tile),
// 2 -> four stages of 128 columns (two N=128 groups per tile).  The MMA issuer may only reuse a stage once all
// epilogue warps of the pair have drained it, and that round trip (commit -> epilogue wake-up -> tcgen05.ld ->
// remote arrive -> issuer wake-up) is longer than one tile's MMAs: with two stages the tile period becomes
// (T_mma + round trip) / 2, with four half-tile stages (T_mma / 2 + round trip) / 2, which is below T_mma.
template <int MODE, int CG, int NH>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_kernel(const TcArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int BROWS = TC_BN / CG;                                // reference rows this CTA stages per tile
  const int kpad = p.kpad;
  const uint32_t a_bytes = (uint32_t)TC_BM * kpad * 2;
  const uint32_t b_bytes = (uint32_t)BROWS * kpad * 2;
  unsigned char* sA = smem;
  unsigned char* sB = smem + a_bytes;
  unsigned char* tail = sB + tc_ring_bytes(kpad, p.nstage, CG);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);          // [nstage]
  uint64_t* empty_bar = full_bar + TC_MAX_STAGES;                  // [nstage]
  constexpr int NS = 2 * NH;                                       // TMEM accumulator stages
  constexpr int SN = TC_BN / NH;                                   // columns per stage = N of one MMA group
  constexpr int RH = BROWS / NH;                                   // reference rows this CTA feeds into one MMA group
  uint64_t* tfull_bar = empty_bar + TC_MAX_STAGES;                 // [4]
  uint64_t* tempty_bar = tfull_bar + 4;                            // [4]  (CG = 2: the leader's are used)
  uint64_t* a_bar = tempty_bar + 4;                                // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_bar + 1);
  // FILTER_SYM: ring of the reference tiles' column thresholds (256 t_j + 32 group maxima per tile), filled by the
  // producer with the tile itself.  A slot is rewritten TC_T8_SLOTS tiles later: the producer runs at most nstage <= 6
  // tiles ahead of the retired MMAs and an MMA is only issued once every epilogue warp has finished the tile three
  // before it, so by then nobody reads the slot any more.
  float* t8ring = reinterpret_cast<float*>(tail + 256);            // [TC_T8_SLOTS][TC_TT_FLOATS]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;      // 0 = leader (issues the MMAs), 1 = peer
  const bool leader = cta_rank == 0;
  const long long tile_m = p.tile_m0 + blockIdx.x;                 // 128-row query tile of this CTA
  const long long n_tiles = p.n_btiles;

  if (threadIdx.x == 0) {
    // CG = 2, leader: a stage is full when its own half has landed (expect_tx arrive) AND the peer's relay has arrived
    const uint32_t nfull = (CG == 2 && leader) ? 2 : 1;
    for (int s = 0; s < p.nstage; ++s) { mbar_init(&full_bar[s], nfull); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < NS; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], CG * TC_EPI_WARPS); }
    mbar_init(a_bar, nfull);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_WARP_TMEM) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();           // barriers (of both CTAs) are initialised
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Register budget by role (setmaxnreg, per warpgroup of 4 warps): the producer / MMA / TMEM warpgroup gives registers
  // back, the four epilogue warpgroups take them: the symmetric epilogue keeps 64 accumulator values per thread live
  // next to two candidate lists and spilled at the kernel-wide 96.
  // (each setmaxnreg sits at the head of its role's branch below, so the compiler sees which budget applies where)

  // Producer and MMA warps run their loops with all 32 lanes converged (waits included) and elect one lane
  // only for the asynchronous instructions: everything stays warp-uniform, so the descriptors live in
  // uniform registers and the issue sequence per tile is a few dozen instructions.  (With the loop under
  // `if (lane == 0)` the compiler wrapped every tcgen05.mma in an election loop: ~170 dependent
  // instructions per tile, and the MMA warp spent 85 % of its time issuing instead of waiting.)
  if (warp >= TC_EPI_WARPS) {
  if (MODE == TC_FILTER_SYM) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
  if (warp == TC_WARP_TMA) {
    // ===== producer: own A tile once, then the ring of reference (half) tiles =====
    if (elect_one()) {
      mbar_expect_tx(a_bar, a_bytes);
      bulk_g2s(sA, p.a_tiles + tile_m * (long long)TC_BM * kpad, a_bytes, a_bar);
    }
    __syncwarp();
    const long long src_step = (long long)CG * BROWS * kpad;
    // FILTER_SYM: the pair of query tile I starts at reference tile (I + first step) mod T and wraps around
    long long jt = (MODE == TC_FILTER_SYM && !p.sym_linear) ? p.blk_lo + ((tile_m >> 1) - p.blk_lo + p.bt0) % p.blk_len : p.bt0;
    const __half* src = p.b_tiles + (CG * jt + cta_rank) * (long long)BROWS * kpad;
    int stage = 0; uint32_t phase = 0;
    int tslot = 0;
    for (long long t = 0; t < n_tiles; ++t) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], b_bytes + (MODE == TC_FILTER_SYM ? (uint32_t)(TC_TT_FLOATS * 4) : 0u));
        bulk_g2s(sB + (size_t)stage * b_bytes, src, b_bytes, &full_bar[stage]);
        if (MODE == TC_FILTER_SYM) bulk_g2s(t8ring + tslot * TC_TT_FLOATS, p.t8 + jt * TC_TT_FLOATS, (uint32_t)(TC_TT_FLOATS * 4), &full_bar[stage]);   // 128 B
      }
      __syncwarp();
      src += src_step;
      if (MODE == TC_FILTER_SYM) {
        if (++jt == p.blk_lo + p.blk_len && !p.sym_linear) { jt = p.blk_lo; src = p.b_tiles + (CG * jt + cta_rank) * (long long)BROWS * kpad; }
        if (++tslot == TC_T8_SLOTS) tslot = 0;
      }
      if (++stage == p.nstage) { stage = 0; phase ^= 1; }
    }
  } else if (warp == TC_WARP_MMA) {
    if (leader) {
      // ===== MMA issuer =====
      mbar_wait(a_bar, 0);
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(SN >> 3) << 17) | ((uint32_t)((CG * TC_BM) >> 4) << 24);
      // one k-step = 16 FP16 = two 16-byte chunks; chunk stride = staged rows * 16 B.  The start-address field
      // of a descriptor counts 16-byte units, so stepping k (or the ring stage, or the row half) is an add on the low word.
      const uint64_t adesc0 = tc_smem_desc(smem_u32(sA), TC_BM * 16, 128);
      const uint64_t bdesc0 = tc_smem_desc(smem_u32(sB), BROWS * 16, 128);
      constexpr uint32_t a_kstep = (2 * TC_BM * 16) >> 4, b_kstep = (2 * BROWS * 16) >> 4;
      const uint32_t b_stage_step = b_bytes >> 4;
      const int ksteps = kpad / 16;
      int stage = 0; uint32_t phase = 0;
      uint32_t sidx = 0;                                          // running accumulator-stage counter
      for (long long t = 0; t < n_tiles; ++t) {
#pragma unroll
        for (int h = 0; h < NH; ++h, ++sidx) {
          const int as = (int)(sidx % NS);
          mbar_wait(&tempty_bar[as], (uint32_t)(((sidx / NS) & 1) ^ 1));
          if (h == 0) mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d_addr = tmem_base + (uint32_t)as * SN;
            const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)stage * b_stage_step + (uint32_t)(h * RH));   // rows h*RH.. : 16 B each
#pragma unroll
            for (int ks = 0; ks < TC_MAX_KPAD / 16; ++ks)
              if (ks < ksteps) tc_mma_f16<CG>(d_addr, adesc0 + (uint64_t)(ks * a_kstep), bdesc + (uint64_t)(ks * b_kstep), idesc, ks > 0 ? 1u : 0u);
            if (h == NH - 1) tc_commit<CG>(&empty_bar[stage]);   // smem stage reusable once these MMAs retire
            tc_commit<CG>(&tfull_bar[as]);                       // accumulator stage ready for the epilogue
          }
          __syncwarp();
        }
        if (++stage == p.nstage) { stage = 0; phase ^= 1; }
      }
    } else {
      // ===== relay (CG = 2, peer CTA): tell the leader that this CTA's operands have landed =====
      mbar_wait(a_bar, 0);
      if (lane == 0) mbar_arrive_remote(a_bar, 0);
      int stage = 0; uint32_t phase = 0;
      for (long long t = 0; t < n_tiles; ++t) {
        mbar_wait(&full_bar[stage], phase);
        if (lane == 0) mbar_arrive_remote(&full_bar[stage], 0);
        if (++stage == p.nstage) { stage = 0; phase ^= 1; }
      }
    }
  }
  } else {
    // ===== epilogue =====
    if (MODE == TC_FILTER_SYM) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int q = warp & 3;                      // TMEM lane quadrant this warp may access
    const int w4 = warp >> 2;                    // which 64-column quarter of the 256-wide tile
    const int r_in_tile = q * 32 + lane;
    const long long grow = tile_m * TC_BM + r_in_tile;
    const bool row_ok = grow >= p.row_lo && grow < p.row_hi;
    const long long slot = grow - p.row_lo;
    float thr = -INFINITY;
    float gmin[32];
    if (MODE == TC_SAMPLE) {
#pragma unroll
      for (int c = 0; c < 32; ++c) gmin[c] = INFINITY;
    }
    // FILTER: this thread's candidate segment (row = thread, column quarter = warp): (column, score) pairs
    const int seg_cap = p.cand_cap >> 2;
    uint2* seg = nullptr;
    int cnt = 0;                                                // candidates in the segment so far
    const int cnt_lim = seg_cap - 32;                           // a 32-column chunk is appended only if it fits entirely
    bool overflow = false;
    if (MODE == TC_FILTER_SYM && row_ok) thr = p.thr[slot];
    if ((MODE == TC_FILTER || MODE == TC_PROBE_NOAPPEND) && row_ok) {
      thr = p.thr[slot];
      seg = p.cand + slot * (long long)p.cand_cap + (long long)w4 * seg_cap;
      if (p.bt0 > 0) cnt = p.cand_count[slot * 4 + w4];         // continue the segment of the earlier reference chunks
      overflow = cnt > seg_cap;
      cnt = min(cnt, seg_cap);
    }
    // FILTER_SYM: the epilogue only PRE-tests groups of 8 columns (minimum of the group against max(t_i, T8(group))) and
    // dumps the 8 accumulator values of a passing (row, group) as one record into the warp's stream of pool chunks (a new
    // chunk costs one atomic per TC_POOL_CH records).  The element tests, for both rows of every pair, happen in
    // tc_transpose_kernel after the launch: per tile the epilogue then costs what the one-sided filter's does, nothing per
    // row is kept in the kernel, and a row that is a candidate of thousands of others costs nothing special.
    float cadd = INFINITY;                                      // c_i: L_ij + c_i <= S^2 d2_ij
    float trow = -INFINITY;                                     // t_i = thr_i + u_i >= thr_i - sigma_i + ||z_i||^2
    int jt_sym = 0, tslot = 0;
    uint4* chunk = nullptr;                                     // warp-uniform: the warp's current pool chunk ...
    int wcnt = 0;                                               // ... and the records in it
    auto close_chunk = [&]() {
      if (lane == 0) chunk[0] = make_uint4((uint32_t)wcnt, 0u, 0u, 0u);
    };
    auto new_chunk = [&]() {
      int c = 0;
      if (lane == 0) c = atomicAdd(p.pool_count, 1);
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c >= p.pool_chunks) { c = p.pool_chunks - 1; if (lane == 0) *p.pool_fail = 1; }   // pool exhausted: the host redoes the build one-sided
      chunk = p.pool + (long long)c * TC_POOL_STRIDE;
      wcnt = 0;
    };
    if (MODE == TC_FILTER_SYM) {
      if (row_ok) {
        const float2 cui = p.cu[slot];
        cadd = cui.x; trow = __fadd_ru(thr, cui.y);
      }
      jt_sym = p.sym_linear ? (int)p.bt0 : (int)(p.blk_lo + ((tile_m >> 1) - p.blk_lo + p.bt0) % p.blk_len);
      new_chunk();
    }
    int probe_hits = 0;
    constexpr int WC = SN / 4;                   // columns of one accumulator stage this warp owns: 64 (NH = 1) or 32
    constexpr int NCHK = WC / 32;                // 32-column register chunks per stage
    // reference position of the first column of register chunk hh of half h:  accumulator column cc = w4 * WC + hh * 32
    // was fed by CTA cc / RH of the pair from row h * RH + cc % RH of its staged BROWS rows
    auto chunk_off = [&](int h, int hh) -> uint32_t {
      const int cc = w4 * WC + hh * 32;
      return (uint32_t)((cc / RH) * BROWS + h * RH + (cc % RH));
    };
    uint32_t jtile = (uint32_t)(p.bt0 * TC_BN) - (uint32_t)TC_BN;    // first position of the current tile; loop-carried
    uint32_t sidx = 0;                                                // running accumulator-stage counter

    for (int t = 0, nt = (int)n_tiles; t < nt; ++t) {      // 32-bit loop state: fewer uniform-datapath instructions per tile
      jtile += (uint32_t)TC_BN;
#pragma unroll
      for (int h = 0; h < NH; ++h, ++sidx) {
      const int as = (int)(sidx % NS);
      // the symmetric epilogue polls (no suspend): its warps wait for the MMA most of the time and the wake-up of a
      // suspended try_wait sits on the accumulator release path (121.8 -> 118.5 ms at 1M; the busier one-sided epilogue
      // lost 8 ms with polling, and polling in the MMA issuer changed nothing)
      if (MODE == TC_FILTER_SYM) mbar_spin(&tfull_bar[as], (uint32_t)((sidx / NS) & 1));
      else mbar_wait(&tfull_bar[as], (uint32_t)((sidx / NS) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * SN + w4 * WC);
      if (MODE == TC_PROBE_NOLD) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (leader) mbar_arrive(&tempty_bar[as]); else mbar_arrive_remote(&tempty_bar[as], 0); }
        continue;
      }
      uint32_t va[32], vb[NCHK == 2 ? 32 : 1];
      if constexpr (MODE == TC_FILTER_SYM && NH == 1 && CG == 2) {
        // ---- symmetric self-join: tile (I, J) serves the rows of I (row test, as FILTER) AND the rows of J: column j
        // gets row i as a candidate iff  L_ij + c_i <= t_j  (both sides are lower / upper bounds of S^2 d2_ij and of
        // thr_j - sigma_j + ||z_j||^2, so whatever is not emitted has d2 > bound_j).  The group minimum is tested first
        // against the largest t_j of the group of 8 columns: one add and one compare per group; the positions of a
        // tile are sorted by threshold (tc_sort_tiles_kernel), so that maximum is close to every t_j of its group.
        const int step = (int)p.bt0 + (int)t;
        const int jt = jt_sym;
        if (++jt_sym == (int)(p.blk_lo + p.blk_len) && !p.sym_linear) jt_sym = (int)p.blk_lo;
        // the diagonal tile holds both orders of every pair already; for an even tile count the opposite tile is
        // multiplied by both of its pairs: row test only in either case
        const bool col_on = p.sym_linear || (step > 0 && step * 2 != (int)p.blk_len);
        tmem_ld32(taddr, va);
        tmem_ld32(taddr + 32, vb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&tempty_bar[as], 0);       // the leader's barrier, from either CTA of the pair
        const float* tt = t8ring + tslot * TC_TT_FLOATS;          // the tile's 32 group maxima of the column thresholds
        if (++tslot == TC_T8_SLOTS) tslot = 0;
        // per group of 8 columns: minimum (three-input min instructions) against max(t_i, largest t_j of the group); the
        // warp's vote decides (uniformly) whether anything is dumped, its bit pattern where each lane's record goes
        const float4 tq0 = *reinterpret_cast<const float4*>(tt + w4 * 8);         // warp-uniform (broadcast) reads
        const float4 tq1 = *reinterpret_cast<const float4*>(tt + w4 * 8 + 4);
        const float tg[8] = {tq0.x, tq0.y, tq0.z, tq0.w, tq1.x, tq1.y, tq1.z, tq1.w};
        const uint32_t jbase = (uint32_t)jt * (uint32_t)TC_BN + (uint32_t)(w4 * 64);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint32_t* v = g < 4 ? va + 8 * g : vb + 8 * (g - 4);
          const float a0 = fminf(fminf(__uint_as_float(v[0]), __uint_as_float(v[1])), __uint_as_float(v[2]));
          const float a1 = fminf(fminf(__uint_as_float(v[3]), __uint_as_float(v[4])), __uint_as_float(v[5]));
          const float m = fminf(fminf(a0, a1), fminf(__uint_as_float(v[6]), __uint_as_float(v[7])));
          const float tm = col_on ? fmaxf(trow, tg[g]) : trow;
          const bool mine = __fadd_rd(m, cadd) <= tm;
          if (__any_sync(0xffffffffu, mine)) {                        // warp-uniform, about one group in six
            const uint32_t bal = __ballot_sync(0xffffffffu, mine);
            if (wcnt > TC_POOL_CH - 32) { close_chunk(); new_chunk(); }
            if (mine) {
              uint4* r = chunk + 1 + 3 * (wcnt + __popc(bal & ((1u << lane) - 1u)));
              r[0] = make_uint4((uint32_t)grow, jbase + 8u * g, 0u, 0u);
              r[1] = make_uint4(v[0], v[1], v[2], v[3]);
              r[2] = make_uint4(v[4], v[5], v[6], v[7]);
            }
            wcnt += __popc(bal);
          }
        }
        continue;
      }
      if (MODE == TC_FILTER || MODE == TC_PROBE_NOAPPEND) {
        // The warp's columns go to registers and the accumulator stage is released BEFORE any of it is
        // reduced: the MMA issuer waits for the slowest of the pair's 32 epilogue warps, so whatever a warp
        // does between "accumulator ready" and its release is on the tensor pipe's critical path.
        tmem_ld32(taddr, va);
        if constexpr (NCHK == 2) tmem_ld32(taddr + 32, vb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (leader) mbar_arrive(&tempty_bar[as]); else mbar_arrive_remote(&tempty_bar[as], 0); }
#pragma unroll
        for (int hh = 0; hh < NCHK; ++hh) {
          const uint32_t* v = hh ? vb : va;
          // minima of the four groups of 8 columns, then of the chunk: three-input min instructions, no branches
          float m8[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float a0 = fminf(fminf(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1])), __uint_as_float(v[8 * g + 2]));
            const float a1 = fminf(fminf(__uint_as_float(v[8 * g + 3]), __uint_as_float(v[8 * g + 4])), __uint_as_float(v[8 * g + 5]));
            m8[g] = fminf(fminf(a0, a1), fminf(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
          }
          const float mn = fminf(fminf(m8[0], m8[1]), fminf(m8[2], m8[3]));
          const bool hit = mn <= thr;            // about one lane in every other warp-chunk
          // The append path is WARP-UNIFORM: votes decide which 8-column groups are walked and every lane walks
          // them with its stores predicated.  (As per-lane branches the hitting lanes diverged by group and the
          // path cost ~145 issue slots per entry with one active thread: 40 % of the kernel's instructions.)
          if (__any_sync(0xffffffffu, hit)) {
            if (MODE == TC_PROBE_NOAPPEND) { probe_hits += hit ? 1 : 0; continue; }
            const bool full = cnt > cnt_lim;     // segment (nearly) full: the row goes to the exact filter
            overflow |= hit && full;
            const bool can = hit && !full;
            const uint32_t j0 = jtile + chunk_off(h, hh);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const bool hg = can && (m8[g] <= thr);
              if (__any_sync(0xffffffffu, hg)) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  // one 8-byte store and a counter bump per passing column; no capacity test here
                  if (hg && __uint_as_float(v[8 * g + c]) <= thr) { seg[cnt] = make_uint2(j0 + 8 * g + c, v[8 * g + c]); ++cnt; }
                }
              }
            }
          }
        }
        continue;
      }
      if (MODE == TC_SAMPLE) {
        // one 32-column chunk at a time: with two chunks AND the 32 running minima live the kernel exceeds its
        // 96 registers and spills the minima every tile (50 local-memory accesses per warp and tile)
        tmem_ld32(taddr, va);
        tmem_ld_wait();
        if constexpr (NCHK == 2) {
#pragma unroll
          for (int c = 0; c < 32; ++c) gmin[c] = fminf(gmin[c], __uint_as_float(va[c]));
          tmem_ld32(taddr + 32, va);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (leader) mbar_arrive(&tempty_bar[as]); else mbar_arrive_remote(&tempty_bar[as], 0); }
#pragma unroll
        for (int c = 0; c < 32; ++c) gmin[c] = fminf(gmin[c], __uint_as_float(va[c]));
        continue;
      }
      tmem_ld32(taddr, va);
      if constexpr (NCHK == 2) tmem_ld32(taddr + 32, vb);
      tmem_ld_wait();
      // the accumulator stage is free as soon as its values sit in registers
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (leader) mbar_arrive(&tempty_bar[as]); else mbar_arrive_remote(&tempty_bar[as], 0); }
      if (MODE == TC_PROBE_LDONLY) {
        if (va[0] == 0x7fc12345u && vb[0] == 0x7fc12345u) ++probe_hits;   // keeps the loads alive
      } else {   // TC_DUMP: positions -> original columns
        if (row_ok) {
#pragma unroll
          for (int hh = 0; hh < NCHK; ++hh) {
            const uint32_t* v = hh ? vb : va;
            const long long j0 = (long long)(p.bt0 + t) * TC_BN + chunk_off(h, hh);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int32_t o0 = p.orig[j0 + c];
              if (o0 != 0x7fffffff) p.dump[slot * p.dump_ld + o0] = __uint_as_float(v[c]);
            }
          }
        }
      }
      }
    }

    // ---- per-row results ----
    if (MODE == TC_FILTER || MODE == TC_FILTER_SYM) {
      // a count above the segment capacity marks the segment as overflowed (the rerank sends the row to the exact filter)
      if (MODE == TC_FILTER && row_ok) p.cand_count[slot * 4 + w4] = overflow ? seg_cap + 1 : cnt;
      if (MODE == TC_FILTER_SYM) close_chunk();
    }
    if ((MODE == TC_PROBE_NOAPPEND || MODE == TC_PROBE_LDONLY) && row_ok && probe_hits == 0x7fffffff) p.cand_count[slot * 4 + w4] = 1;
    if (MODE == TC_SAMPLE) {
      asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");
      // all MMAs have retired (last tfull observed by every epilogue warp) -> the B ring
      // is free: use it as scratch
      float* gm = reinterpret_cast<float*>(sB);                    // [TC_BM rows][TC_NGROUP + 1]: conflict-free both ways
#pragma unroll
      for (int c = 0; c < 32; ++c) gm[r_in_tile * (TC_NGROUP + 1) + w4 * 32 + c] = gmin[c];
      asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory");
      // r-th smallest of each row's TC_NGROUP group minima: one warp per row (8 rows per warp), four ordered keys
      // per lane, the answer built bit by bit from warp-wide counts (32 short steps instead of r serial scans
      // of all groups by one thread, which cost more than the sample's MMAs)
      for (int rr = 0; rr < TC_BM / TC_EPI_WARPS; ++rr) {
        const int r_sel = warp * (TC_BM / TC_EPI_WARPS) + rr;
        const long long grow_sel = tile_m * TC_BM + r_sel;
        if (grow_sel < p.row_lo || grow_sel >= p.row_hi) continue;  // warp-uniform
        uint32_t key[TC_NGROUP / 32];
#pragma unroll
        for (int u = 0; u < TC_NGROUP / 32; ++u) key[u] = tc_f2ord(gm[r_sel * (TC_NGROUP + 1) + lane + 32 * u]);
        uint32_t res = 0;                                            // largest v with #(keys < v) < r
        for (int bit = 31; bit >= 0; --bit) {
          const uint32_t tv = res | (1u << bit);
          int cnt = 0;
#pragma unroll
          for (int u = 0; u < TC_NGROUP / 32; ++u) cnt += key[u] < tv ? 1 : 0;
          cnt = __reduce_add_sync(0xffffffffu, cnt);
          if (cnt < p.order_stat) res = tv;
        }
        if (lane == 0) {
          float thr_f; double bound_f;
          const long long rid = p.slot_row ? (long long)p.slot_row[grow_sel] : grow_sel;
          tc_threshold(tc_ord2f(res), p.norms4[rid], p.nrm[rid], kpad, p.d, p.scale[0], &thr_f, &bound_f);
          p.thr[grow_sel - p.row_lo] = thr_f;
          p.bound[grow_sel - p.row_lo] = bound_f;
        }
      }
    }
  }

  tc_fence_before();
  // CG = 2: neither CTA leaves (or frees TMEM) while the other may still signal it or read its operands
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == TC_WARP_TMEM) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static double binom_tail_ge(int n, double p, int r) {   // P(Binomial(n,p) >= r)
  if (r <= 0) return 1.0;
  if (p <= 0) return 0.0;
  double q = 1.0 - p, term = pow(q, n), cdf = 0.0;      // k = 0
  for (int k = 0; k < r; ++k) {
    cdf += term;
    term *= (double)(n - k) / (double)(k + 1) * p / q;
  }
  return std::max(0.0, 1.0 - cdf);
}

TcPlan tc_make_plan(long long n_ref, int d, int k) {
  TcPlan p;
  // split ("precision") columns are chosen from the data by tc_prepare; size the staging
  // buffers for the most it may choose
  p.max_split = std::max(0, std::min(TC_MAX_SPLIT, std::min(d, (TC_MAX_KPAD - d - 5) / 2)));
  p.kpad_alloc = (d + 5 + 2 * p.max_split + 15) / 16 * 16;
  p.kpad = p.kpad_alloc;          // replaced by the actual width in tc_prepare
  p.ksteps = p.kpad / 16;
  p.n_tiles = (n_ref + TC_BN - 1) / TC_BN;
  p.n_groups = TC_NGROUP;
  p.stage_bytes = (size_t)TC_BN * p.kpad * 2;
  const int k1 = k + 1;
  // choose (sample size M, order statistic r): P(threshold admits < k+1 refs) <= 1e-4,
  // minimise   expected candidates + (cost of the sample pass expressed in candidates)
  double best_cost = 1e300; int best_m = 0, best_r = 0; double best_cnt = 0, best_sd = 0;
  for (long long m = 2048; m <= 65536; m <<= 1) {
    if (m > n_ref && m > 2048) break;
    const long long mm = std::min<long long>(m, (n_ref / TC_BN) * TC_BN > 0 ? (n_ref / TC_BN) * TC_BN : TC_BN);
    const double per_group = (double)mm / TC_NGROUP;
    const double pg = 1.0 - exp(-(double)k1 * per_group / (double)n_ref);
    int r = 1;
    while (r <= TC_NGROUP / 2 && binom_tail_ge(TC_NGROUP, pg, r) > 1e-4) ++r;
    if (r > TC_NGROUP / 2) continue;
    double h = 0, h2 = 0;
    for (int i = 0; i < r; ++i) { h += 1.0 / (TC_NGROUP - i); h2 += 1.0 / ((double)(TC_NGROUP - i) * (TC_NGROUP - i)); }
    const double scale = (double)n_ref / per_group;
    const double cnt = scale * h, sd = scale * sqrt(h2);
    const double cost = cnt + 0.004 * (double)mm;     // one sampled reference costs ~0.004 reranked candidates per row
    if (cost < best_cost) { best_cost = cost; best_m = (int)mm; best_r = r; best_cnt = cnt; best_sd = sd; }
  }
  if (best_m == 0) {            // tiny n: everything is a candidate
    best_m = (int)std::min<long long>(2048, p.n_tiles * TC_BN); best_r = TC_NGROUP / 2;
    best_cnt = (double)n_ref; best_sd = 0;
  }
  p.sample_size = (best_m + TC_BN - 1) / TC_BN * TC_BN;
  p.order_stat = best_r;
  // candidates are written to 4 per-row segments (one per epilogue warp of a TMEM lane quadrant,
  // each covering a fixed quarter of every reference tile): size a segment for total/4 + 6 sigma
  double tot = (best_cnt + 6.0 * best_sd) * 1.5 + 64;
  tot = std::max(tot, (double)(2 * k1));
  tot = std::min(tot, (double)n_ref + 4.0);
  double seg = tot / 4.0 + 6.0 * sqrt(tot / 4.0) + 16.0;
  seg = std::min(seg, 2048.0);
  p.cand_cap = 4 * ((int)((seg + 63) / 64) * 64);
  p.cand_cap_planned = p.cand_cap;
  p.pool_limit = 0;
  return p;
}

long long tc_chunk_rows(const TcPlan& p, long long n_ref) {
  // keep the candidate lists of one chunk under ~12 GB
  long long rows = (long long)(40e9 / ((double)p.cand_cap * 8.0));
  rows = std::max<long long>(rows / TC_BM * TC_BM, TC_BM * 148);
  return std::min<long long>(rows, 1 << 20);
}

// chunks of the group-record pool: one partly filled chunk per epilogue warp of a launch (2T CTAs x 16 warps) plus the
// records of the busiest launch (sized for an eighth of the planned list capacity of every row: the first launches run on
// the loosest thresholds, 3-4x the measured volume at 200k and 1M rows)
static int tc_pool_chunks(const TcPlan& p, long long n_ref) {
  const double recs = (double)n_ref * (double)p.cand_cap_planned / 8.0;
  const double c = 32.0 * (double)p.n_tiles + recs / TC_POOL_CH + 1024.0;
  return (int)std::min(c, 3.0e6);
}
static constexpr double TC_SYM_MAX_BYTES = 64e9;    // candidate lists + outboxes of all rows must be resident at once
bool tc_sym_supported(const TcPlan& p, long long n_ref) {
  return TC_DEFAULT_CG == 2 && TC_DEFAULT_NH == 1 && p.n_tiles >= 8 &&
         (double)n_ref * (double)p.cand_cap * 8.0 <= TC_SYM_MAX_BYTES;
}

static void tc_carve(Arena& a, const TcPlan& p, long long n_ref, long long chunk, TcState* st, bool sym) {
  const long long n_a = (n_ref + TC_BM - 1) / TC_BM * TC_BM;
  const long long n_b = p.n_tiles * TC_BN;
  (void)n_a;   // A tiles are staged by the same kernel that walks the n_b padded rows
  st->a_tiles = a.take<__half>((size_t)n_b * p.kpad_alloc);
  st->b_tiles = a.take<__half>((size_t)n_b * p.kpad_alloc);
  st->centre = a.take<double>(4096);
  st->colsum = a.take<double>(2 * 4096);     // column sums, then column sums of squares
  st->scale = a.take<double>(8);
  st->nrm = a.take<double>((size_t)n_b);
  st->norms4 = a.take<float4>((size_t)n_b);
  st->sc = a.take<TcScalars>(2);
  st->orig = a.take<int32_t>((size_t)n_b);
  st->thr = a.take<float>((size_t)chunk + TC_BN);        // + padding positions of the last tile (symmetric filter)
  st->maxabs = a.take<unsigned long long>(8);
  st->sym = sym ? 1 : 0;
  st->rpos = nullptr; st->cu = nullptr; st->t8 = nullptr; st->pool = nullptr; st->pool_count = nullptr; st->pool_chunks = 0;
  if (sym) {
    st->rpos = a.take<int32_t>((size_t)n_b);
    st->cu = a.take<float2>((size_t)n_b);
    st->tq = a.take<float4>((size_t)n_b);
    st->t8 = a.take<float>((size_t)p.n_tiles * TC_TT_FLOATS);
    st->pool_chunks = tc_pool_chunks(p, n_ref);
    if (p.pool_limit > 0) st->pool_chunks = std::min(st->pool_chunks, p.pool_limit);
    st->pool = a.take<uint4>((size_t)st->pool_chunks * TC_POOL_STRIDE);
    st->pool_count = a.take<int>(64);                    // [0] chunks handed out, [1] chunks taken by the transpose, [2] pool exhausted
  }
}

size_t tc_workspace_bytes(const TcPlan& p, long long n_ref, long long n_query, bool sym) {
  Arena a(nullptr, 0);
  TcState st;
  tc_carve(a, p, n_ref, n_query, &st, sym);
  return a.off + 1024;
}

static size_t tc_smem_bytes(int kpad, int nstage, int cg) {
  return (size_t)TC_BM * kpad * 2 + tc_ring_bytes(kpad, nstage, cg) + 256 /* barriers, TMEM slot */ +
         TC_T8_SLOTS * TC_TT_FLOATS * 4 /* column-threshold ring of the symmetric filter */ + 64;
}

static int tc_pick_stages(int kpad, int cg) {
  for (int s = TC_MAX_STAGES; s >= 2; --s)
    if (tc_smem_bytes(kpad, s, cg) <= 227 * 1024) return s;
  return 0;
}

// CTA pairs (cta_group::2) or single CTAs.  The environment is consulted by the probe build only
// (-DSBK_ENABLE_PROBES): SBK_TC_CG=1|2, SBK_TC_L2_MB, SBK_TC_PROBE, SBK_TC_REFINE.
#ifdef SBK_ENABLE_PROBES
static int tc_env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
#else
static int tc_env_int(const char*, int dflt) { return dflt; }
#endif
static int tc_cta_group() {
  static const int cg = tc_env_int("SBK_TC_CG", TC_DEFAULT_CG);
  return cg == 2 ? 2 : 1;
}

static long long coprime_multiplier(long long n) {
  long long m = (long long)(0.6180339887498949 * (double)n);
  if (m < 1) m = 1;
  auto gcd = [](long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; };
  while (gcd(m, n) != 1) ++m;
  return m;
}

// a^-1 mod n for gcd(a, n) = 1 (extended Euclid)
static long long mod_inverse(long long a, long long n) {
  long long t = 0, nt = 1, r = n, nr = a % n;
  while (nr != 0) {
    const long long q = r / nr;
    long long tmp = t - q * nt; t = nt; nt = tmp;
    tmp = r - q * nr; r = nr; nr = tmp;
  }
  if (t < 0) t += n;
  return t;
}

int tc_prepare(const void* X, int dtype, long long n_ref, int d, long long ld, TcPlan& p,
               long long chunk, void* ws, size_t ws_bytes, TcState* st, cudaStream_t stream, bool sym) {
  if ((d + 5 + 15) / 16 * 16 > TC_MAX_KPAD || d > 4096) {
    set_error("tensor kNN path supports feature width <= %d (got %d); use SBK_KNN_EXACT", TC_MAX_KPAD - 5, d);
    return SBK_ERR_INVALID;
  }
  Arena a(ws, ws_bytes);
  tc_carve(a, p, n_ref, chunk, st, sym);
  if (!a.ok()) { set_error("tensor kNN: staging workspace %zu < %zu", ws_bytes, a.off); return SBK_ERR_WORKSPACE; }
  st->n_ref = n_ref; st->d = d;
  SBK_CUDA_CHECK(cudaMemsetAsync(st->colsum, 0, 2 * 4096 * sizeof(double), stream));
  SBK_CUDA_CHECK(cudaMemsetAsync(st->maxabs, 0, 8 * sizeof(unsigned long long), stream));
  SBK_CUDA_CHECK(cudaMemsetAsync(st->sc, 0, 2 * sizeof(TcScalars), stream));
  const unsigned cb = (unsigned)((n_ref + TC_COLSTAT_ROWS - 1) / TC_COLSTAT_ROWS);
  const long long n_b = p.n_tiles * TC_BN;
  const long long n_a = (n_ref + TC_BM - 1) / TC_BM * TC_BM;
  // zero the A padding rows beyond n_b (n_a <= n_b always since TC_BN is a multiple of TC_BM)
  (void)n_a;
  if (dtype == SBK_F32) tc_colstat_kernel<float><<<cb, 256, 0, stream>>>((const float*)X, n_ref, d, ld, st->colsum, st->colsum + 4096, st->maxabs);
  else tc_colstat_kernel<double><<<cb, 256, 0, stream>>>((const double*)X, n_ref, d, ld, st->colsum, st->colsum + 4096, st->maxabs);
  SBK_LAUNCH_CHECK();
  tc_centre_kernel<<<1, 256, 0, stream>>>(st->colsum, st->colsum + 4096, st->maxabs, n_ref, d, p.max_split, st->centre, st->scale, st->sc);
  SBK_LAUNCH_CHECK();
  // the operand width depends on how many precision columns the data asked for
  int n_split = 0;
  SBK_CUDA_CHECK(cudaMemcpyAsync(&n_split, &st->sc->n_split, sizeof(int), cudaMemcpyDeviceToHost, stream));
  SBK_CUDA_CHECK(cudaStreamSynchronize(stream));
  p.n_split = n_split;
  p.kpad = (d + 5 + 2 * n_split + 15) / 16 * 16;
  p.ksteps = p.kpad / 16;
  p.stage_bytes = (size_t)TC_BN * p.kpad * 2;
  st->cg = tc_cta_group();
  st->nstage = tc_pick_stages(p.kpad, st->cg);
  if (p.kpad > p.kpad_alloc || st->nstage == 0) { set_error("tensor kNN: operand width %d exceeds the staged capacity", p.kpad); return SBK_ERR_INTERNAL; }
  // pseudo-random (golden-ratio stride) order of the reference side: position p holds row (p * mult + off) mod n
  st->perm_mult = coprime_multiplier(n_ref);
  st->perm_off = n_ref / 3;
  st->perm_minv = mod_inverse(st->perm_mult, n_ref);
  const unsigned sb = (unsigned)((n_b + 127) / 128);
  if (dtype == SBK_F32)
    tc_stage_kernel<float><<<sb, 128, 0, stream>>>((const float*)X, n_ref, d, ld, n_b, st->perm_minv, st->perm_off, st->centre, st->scale,
                                                   p.kpad, TC_BN / st->cg, st->b_tiles, st->a_tiles, st->orig, st->nrm, st->norms4, st->sc,
                                                   st->sym, st->rpos, st->cu);
  else
    tc_stage_kernel<double><<<sb, 128, 0, stream>>>((const double*)X, n_ref, d, ld, n_b, st->perm_minv, st->perm_off, st->centre, st->scale,
                                                    p.kpad, TC_BN / st->cg, st->b_tiles, st->a_tiles, st->orig, st->nrm, st->norms4, st->sc,
                                                    st->sym, st->rpos, st->cu);
  SBK_LAUNCH_CHECK();
  if (st->sym && st->cg != 2) { set_error("tensor kNN: the symmetric filter needs CTA pairs"); return SBK_ERR_INTERNAL; }
  return SBK_OK;
}

template <int MODE, int CG, int NH>
static int tc_launch_cg(const TcArgs& args, long long n_mtiles, cudaStream_t stream) {
  const size_t smem = tc_smem_bytes(args.kpad, args.nstage, CG);
  SBK_CUDA_CHECK(cudaFuncSetAttribute(tc_kernel<MODE, CG, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // one CTA per 128-row query tile; CG = 2 launches them as clusters of two (the grid must then be even)
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)((n_mtiles + CG - 1) / CG * CG));
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  SBK_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc_kernel<MODE, CG, NH>, args));
  return SBK_OK;
}
template <int MODE>
static int tc_launch(const TcArgs& args, long long n_mtiles, int cg, cudaStream_t stream) {
  // accumulator stages per tile: four half-tile stages by default (see tc_kernel); SBK_TC_NH=1 (probe build) = two full-tile stages
  static const int nh = tc_env_int("SBK_TC_NH", TC_DEFAULT_NH);
  if (nh == 1) return cg == 2 ? tc_launch_cg<MODE, 2, 1>(args, n_mtiles, stream) : tc_launch_cg<MODE, 1, 1>(args, n_mtiles, stream);
  return cg == 2 ? tc_launch_cg<MODE, 2, 2>(args, n_mtiles, stream) : tc_launch_cg<MODE, 1, 2>(args, n_mtiles, stream);
}

static void tc_chunk_args(TcState* st, const TcPlan& p, long long row0, long long n_rows, TcArgs& a,
                          long long& n_mtiles) {
  memset(&a, 0, sizeof(a));
  a.a_tiles = st->a_tiles;
  a.kpad = p.kpad; a.nstage = st->nstage;
  a.tile_m0 = row0 / (2 * TC_BM) * 2;      // even first tile: the odd partner of a CTA pair never runs past the staged tiles
  n_mtiles = (row0 + n_rows + TC_BM - 1) / TC_BM - a.tile_m0;
  a.row_lo = row0; a.row_hi = row0 + n_rows;
  a.order_stat = p.order_stat;
  a.norms4 = st->norms4; a.nrm = st->nrm; a.sc = st->sc; a.scale = st->scale; a.d = st->d;
  a.thr = st->thr; a.cand_cap = p.cand_cap;
  a.slot_row = st->sym ? st->orig : nullptr;      // symmetric build: slots are staged positions
}

int tc_sample_chunk(TcState* st, const TcPlan& p, long long row0, long long n_rows, double* bound,
                    cudaStream_t stream) {
  TcArgs a; long long n_mtiles;
  tc_chunk_args(st, p, row0, n_rows, a, n_mtiles);
  a.bound = bound;
  a.b_tiles = st->b_tiles; a.n_btiles = p.sample_size / TC_BN;   // the first staged references are the sample
  return tc_launch<TC_SAMPLE>(a, n_mtiles, st->cg, stream);
}

// smallest r with P(Binomial(n, p) >= r) <= eps
static int binom_quantile(int n, double p, double eps) {
  int r = (int)(n * p);
  while (r <= n && binom_tail_ge(n, p, r) > eps) ++r;
  return r;
}

int tc_filter_chunk(TcState* st, const TcPlan& p, long long row0, long long n_rows, int k,
                    uint2* cand, int32_t* cand_count, double* bound, int* n_launches, int* n_refines, cudaStream_t stream) {
  TcArgs a; long long n_mtiles;
  tc_chunk_args(st, p, row0, n_rows, a, n_mtiles);
  a.cand = cand; a.cand_count = cand_count;
  a.b_tiles = st->b_tiles;
  // Walk the references in chunks that stay resident in L2 while every query tile of the launch
  // streams them: one launch per chunk, the per-segment counters carry over in cand_count.
  // (One launch over all references lets the CTAs drift apart until the working set is the
  // whole operand: 553 GB of DRAM reads per 1M x 1M pass instead of ~2 GB.)
  static const long long l2_bytes = (long long)tc_env_int("SBK_TC_L2_MB", (int)(TC_L2_CHUNK_BYTES >> 20)) << 20;
  const long long per = std::max<long long>(1, l2_bytes / (long long)p.stage_bytes);
  const long long n_launch = (p.n_tiles + per - 1) / per;
  const long long even = (p.n_tiles + n_launch - 1) / n_launch;
  static const int probe = tc_env_int("SBK_TC_PROBE", 0);
  static const int refine = tc_env_int("SBK_TC_REFINE", 1);
  int li = 0;
  for (long long b0 = 0; b0 < p.n_tiles; b0 += even, ++li) {
    a.bt0 = b0; a.n_btiles = std::min(even, p.n_tiles - b0);
    int rc = probe == 3 ? tc_launch<TC_PROBE_NOLD>(a, n_mtiles, st->cg, stream)
           : probe == 4 ? tc_launch<TC_PROBE_LDONLY>(a, n_mtiles, st->cg, stream)
           : probe == 5 ? tc_launch<TC_PROBE_NOAPPEND>(a, n_mtiles, st->cg, stream)
                        : tc_launch<TC_FILTER>(a, n_mtiles, st->cg, stream);
    if (rc) return rc;
    if (n_launches) ++*n_launches;
    // threshold refinement from the candidates seen so far (after launches 1, 2 and 4; later ones gain little)
    const long long seen = std::min<long long>(st->n_ref, (b0 + a.n_btiles) * TC_BN);
    if (refine && probe == 0 && b0 + even < p.n_tiles && (li == 0 || li == 1 || li == 3)) {
      const double f = (double)seen / (double)st->n_ref;
      const int r = binom_quantile(k, f, 2e-6) + 1;       // +1: the row itself is always among the seen or unseen k+1
      if (r <= TC_REFINE_MAX / 2) {
        tc_refine_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, stream>>>(cand, cand_count, p.cand_cap, row0, n_rows, r,
                                                                          st->norms4, st->nrm, st->scale, p.kpad, st->d, st->thr, bound, nullptr);
        SBK_LAUNCH_CHECK();
        if (n_refines) ++*n_refines;
      }
    }
  }
  return SBK_OK;
}

// Symmetric self-join filter.  With both operands in the same (pseudo-random) position order, the CTA pair of query
// tile I multiplies reference tiles I, I+1, ..., I+floor(T/2) (mod T): every unordered pair of tiles exactly once
// (the opposite tile twice when T is even; it is then row-tested only, like the diagonal).  Each product serves the
// rows of I directly and the rows of J through the column test; the column hits of a launch sit in per-row outboxes
// and tc_transpose_kernel appends them to the candidate lists of their rows before anything reads those lists.
// Half the tensor work of the one-sided filter for the same candidate lists.
// The steps are cut into launches (a) at 1/14, 1/7 and 2/7 of T, after which every row has seen a window of 1/7, 2/7
// and 4/7 of the (pseudo-randomly ordered) references around its own tile -- the same threshold refinement as the
// one-sided filter -- and (b) so that the reference tiles a launch walks stay L2-resident.
// ---- pieces shared by the two schedules of the symmetric filter ----
static void tc_sym_args(TcState* st, const TcPlan& p, uint2* cand, int32_t* cand_count, TcArgs& a) {
  long long n_mtiles;
  tc_chunk_args(st, p, 0, st->n_ref, a, n_mtiles);
  a.cand = cand; a.cand_count = cand_count; a.b_tiles = st->b_tiles;
  a.blk_lo = 0; a.blk_len = p.n_tiles; a.sym_linear = 0; a.t8 = st->t8; a.cu = st->cu;
  a.pool = st->pool; a.pool_count = st->pool_count; a.pool_chunks = st->pool_chunks; a.pool_fail = st->pool_count + 2;
}
static long long tc_sym_max_steps(const TcPlan& p) {
  static const long long l2_bytes = (long long)tc_env_int("SBK_TC_L2_MB", (int)(TC_L2_CHUNK_BYTES >> 20)) << 20;
  return std::max<long long>(16, l2_bytes / (long long)p.stage_bytes);
}
// SBK_SHARD_TRACE=1: device and host time between named points of the sharded build, printed per rank (development aid)
struct TcTraceMark { const char* what; cudaEvent_t ev; double host_ms; };
static std::vector<TcTraceMark> g_trace_marks;
void tc_trace(const char* what, cudaStream_t stream) {
  static const int on = tc_env_int("SBK_SHARD_TRACE", 0);
  if (!on) return;
  TcTraceMark m; m.what = what;
  cudaEventCreate(&m.ev); cudaEventRecord(m.ev, stream);
  timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  m.host_ms = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  g_trace_marks.push_back(m);
}
void tc_trace_dump(int rank) {
  if (g_trace_marks.empty()) return;
  cudaDeviceSynchronize();
  std::string line = "shard trace rank " + std::to_string(rank) + ":";
  for (size_t i = 1; i < g_trace_marks.size(); ++i) {
    float ms = 0.f; cudaEventElapsedTime(&ms, g_trace_marks[i - 1].ev, g_trace_marks[i].ev);
    char buf[128]; snprintf(buf, sizeof(buf), " %s %.2f(h%.2f)", g_trace_marks[i].what, ms, g_trace_marks[i].host_ms - g_trace_marks[i - 1].host_ms);
    line += buf;
  }
  fprintf(stderr, "%s\n", line.c_str());
  for (auto& m : g_trace_marks) cudaEventDestroy(m.ev);
  g_trace_marks.clear();
}

// one filter launch over query tiles [q_lo, q_lo + q_len) and the transposes of its records
static int tc_sym_launch(TcState* st, const TcPlan& p, TcArgs& a, long long q_lo, long long q_len, long long blk_tiles,
                         uint2* cand, int32_t* cand_count, int* n_launches, int* n_aux_launches, cudaStream_t stream,
                         const TcMail* mail = nullptr) {
  TcMail mv; memset(&mv, 0, sizeof(mv));
  if (mail) mv = *mail;
  a.tile_m0 = 2 * q_lo;
  int rc = tc_launch_cg<TC_FILTER_SYM, 2, 1>(a, 2 * q_len, stream);
  if (rc) return rc;
  tc_trace(a.sym_linear ? "lin" : "circ", stream);
  if (n_launches) ++*n_launches;
  unsigned long long* dbg = nullptr;
#ifdef SBK_ENABLE_PROBES
  static const int sym_debug = tc_env_int("SBK_SYM_DEBUG", 0);
  if (sym_debug) { dbg = reinterpret_cast<unsigned long long*>(st->maxabs) + 4; SBK_CUDA_CHECK(cudaMemsetAsync(dbg, 0, 32, stream)); }
#endif
  const unsigned tr_grid = (unsigned)std::min((st->pool_chunks + 7) / 8, 148 * 8);
  if (mail) tc_transpose_kernel<true><<<tr_grid, 256, 0, stream>>>(st->pool, st->pool_count, st->pool_chunks, st->n_ref, p.n_tiles, blk_tiles, st->tq, cand,
                                                                   cand_count, p.cand_cap, dbg, mv);
  else tc_transpose_kernel<false><<<tr_grid, 256, 0, stream>>>(st->pool, st->pool_count, st->pool_chunks, st->n_ref, p.n_tiles, blk_tiles, st->tq, cand,
                                                               cand_count, p.cand_cap, dbg, mv);
  SBK_LAUNCH_CHECK();
  SBK_CUDA_CHECK(cudaMemsetAsync(st->pool_count, 0, 2 * sizeof(int), stream));  // chunks handed out, chunks transposed; [2], the exhaustion flag, stays
  tc_trace("tr", stream);
#ifdef SBK_ENABLE_PROBES
  if (sym_debug) {
    unsigned long long h[4];
    SBK_CUDA_CHECK(cudaMemcpyAsync(h, dbg, 32, cudaMemcpyDeviceToHost, stream));
    SBK_CUDA_CHECK(cudaStreamSynchronize(stream));
    fprintf(stderr, "sym launch: query tiles [%lld,%lld) %s first %lld count %lld: list-segment overflows %llu, group records %llu, pool of %d chunks\n",
            q_lo, q_lo + q_len, a.sym_linear ? "x tiles" : "circulant steps", a.bt0, a.n_btiles, h[1], h[2], st->pool_chunks);
  }
#endif
  if (n_aux_launches) ++*n_aux_launches;
  return SBK_OK;
}
// threshold refinement of the positions [slot_lo, slot_hi) that have seen the fraction f of the references, then the
// column thresholds of everything
static int tc_sym_refine(TcState* st, const TcPlan& p, int k, double f, long long slot_lo, long long slot_hi,
                         uint2* cand, int32_t* cand_count, double* bound, int* n_aux_launches, cudaStream_t stream) {
  static const int refine = tc_env_int("SBK_TC_REFINE", 1);
  const int r = binom_quantile(k, f, 2e-6) + 1;
  if (!refine || !(f < 1.0) || r > TC_REFINE_MAX / 2 || slot_hi <= slot_lo) return SBK_OK;
  const long long ns = slot_hi - slot_lo;
  tc_refine_kernel<<<(unsigned)((ns + 7) / 8), 256, 0, stream>>>(cand + slot_lo * p.cand_cap, cand_count + slot_lo * 4, p.cand_cap, 0, ns, r,
                                                                st->norms4, st->nrm, st->scale, p.kpad, st->d, st->thr + slot_lo,
                                                                bound + slot_lo, st->orig + slot_lo);
  SBK_LAUNCH_CHECK();
  tc_colthr_kernel<<<(unsigned)p.n_tiles, 256, 0, stream>>>(st->thr, st->cu, st->n_ref, st->t8, st->tq);
  SBK_LAUNCH_CHECK();
  if (n_aux_launches) *n_aux_launches += 2;
  tc_trace("refine", stream);
  return SBK_OK;
}

int tc_sym_begin(TcState* st, const TcPlan& p, uint2* cand, int32_t* cand_count, double* bound, int* n_aux_launches,
                 cudaStream_t stream) {
  if (!st->sym || st->cg != 2) { set_error("tensor kNN: state not prepared for the symmetric filter"); return SBK_ERR_INTERNAL; }
  const long long n = st->n_ref, T = p.n_tiles;
  SBK_CUDA_CHECK(cudaMemsetAsync(cand_count, 0, (size_t)n * 4 * sizeof(int32_t), stream));
  SBK_CUDA_CHECK(cudaMemsetAsync(st->pool_count, 0, 64 * sizeof(int), stream));
  // positions of every tile in descending-threshold order, then the column thresholds of that order
  const size_t sm = (size_t)TC_BN * p.kpad * 2;
  SBK_CUDA_CHECK(cudaFuncSetAttribute(tc_sort_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  tc_sort_tiles_kernel<<<(unsigned)T, 256, sm, stream>>>(n, p.kpad, st->a_tiles, st->b_tiles, st->thr, bound, st->cu, st->orig, st->rpos);
  SBK_LAUNCH_CHECK();
  tc_colthr_kernel<<<(unsigned)T, 256, 0, stream>>>(st->thr, st->cu, n, st->t8, st->tq);
  SBK_LAUNCH_CHECK();
  if (n_aux_launches) *n_aux_launches += 2;
  return SBK_OK;
}

int tc_refresh_colthr(TcState* st, const TcPlan& p, cudaStream_t stream) {
  tc_colthr_kernel<<<(unsigned)p.n_tiles, 256, 0, stream>>>(st->thr, st->cu, st->n_ref, st->t8, st->tq);
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}

int tc_sym_end(TcState* st, cudaStream_t stream) {
  // a launch that ran out of pool chunks lost group records: the caller redoes the build with the one-sided filter
  int h_fail = 0;
  SBK_CUDA_CHECK(cudaMemcpyAsync(&h_fail, st->pool_count + 2, sizeof(int), cudaMemcpyDeviceToHost, stream));
  SBK_CUDA_CHECK(cudaStreamSynchronize(stream));
  return h_fail ? TC_SYM_POOL_EXHAUSTED : SBK_OK;
}

int tc_filter_sym(TcState* st, const TcPlan& p, int k, uint2* cand, int32_t* cand_count, double* bound,
                  int* n_launches, int* n_aux_launches, cudaStream_t stream) {
  int rc = tc_sym_begin(st, p, cand, cand_count, bound, n_aux_launches, stream);
  if (rc) return rc;
  const long long n = st->n_ref, T = p.n_tiles;
  TcArgs a;
  tc_sym_args(st, p, cand, cand_count, a);
  const long long n_steps = T / 2 + 1;                         // steps 0 .. floor(T/2)
  const long long max_steps = tc_sym_max_steps(p);
  // launch boundaries: the refinement points, then an even split of whatever is still longer than max_steps
  std::vector<long long> cuts;
  std::vector<char> refine_at;
  {
    const long long marks[3] = {(T + 13) / 14, (T + 6) / 7, (2 * T + 6) / 7};
    long long prev = 0;
    auto add_range = [&](long long end, bool is_mark) {
      if (end <= prev) return;
      const long long len = end - prev, parts = (len + max_steps - 1) / max_steps;
      for (long long q = 1; q <= parts; ++q) {
        cuts.push_back(prev + len * q / parts);
        refine_at.push_back(is_mark && q == parts);
      }
      prev = end;
    };
    for (int m = 0; m < 3; ++m) if (marks[m] < n_steps) add_range(marks[m], true);
    add_range(n_steps, false);
  }
  long long s0 = 0;
  for (size_t li = 0; li < cuts.size(); ++li) {
    a.bt0 = s0; a.n_btiles = cuts[li] - s0;
    rc = tc_sym_launch(st, p, a, 0, T, T, cand, cand_count, n_launches, n_aux_launches, stream);
    if (rc) return rc;
    s0 = cuts[li];
    if (refine_at[li] && s0 < n_steps) {
      // rows of tile I have now seen tiles I-s0+1 .. I+s0-1
      rc = tc_sym_refine(st, p, k, std::min(1.0, (double)(2 * s0 - 1) * TC_BN / (double)n), 0, n, cand, cand_count, bound,
                         n_aux_launches, stream);
      if (rc) return rc;
    }
  }
  return tc_sym_end(st, stream);
}

// Block-row schedule of the symmetric filter (host-buffer builds): the T tiles are cut into n_blocks blocks; block row r
// multiplies the pairs (block r, block c >= r) -- its own block circulantly, the later blocks tile by tile -- after
// which every pair that involves a row of block r has been multiplied: the rows of block r are FINISHED and can be
// reranked and downloaded while the later block rows are multiplied.  (In the whole-matrix circulant no row is finished
// before the last launch, and a 3.2 GB download would start when the GPU has nothing left to hide it behind.)
// Thresholds: block 0 refines after its own block (it has seen 1/n_blocks of the references); at the start of block row
// r >= 1 all unfinished rows have seen the blocks before r completely and are refined with f = r / n_blocks.
// (a block is a whole number of waves of CTA pairs, 74 on a B200, so that only the last launch of a block row has a ragged tail)
long long tc_sym_block_tiles(const TcPlan& p, int n_blocks) {
  const long long tb = (p.n_tiles + n_blocks - 1) / n_blocks;
  return tb >= 74 ? (tb + 73) / 74 * 74 : tb;
}

int tc_sym_block(TcState* st, const TcPlan& p, int k, uint2* cand, int32_t* cand_count, double* bound, int n_blocks, int r,
                 int* n_launches, int* n_aux_launches, cudaStream_t stream) {
  const long long n = st->n_ref, T = p.n_tiles, tb = tc_sym_block_tiles(p, n_blocks);
  const long long lo = (long long)r * tb, len = std::min(tb, T - lo);
  if (len <= 0) return SBK_OK;
  TcArgs a;
  tc_sym_args(st, p, cand, cand_count, a);
  const long long max_steps = tc_sym_max_steps(p);
  int rc;
  if (r > 0) {
    rc = tc_sym_refine(st, p, k, (double)lo / (double)T, lo * TC_BN, n, cand, cand_count, bound, n_aux_launches, stream);
    if (rc) return rc;
  }
  // own block, circulant inside it
  a.blk_lo = lo; a.blk_len = len; a.sym_linear = 0;
  const long long n_steps = len / 2 + 1;
  for (long long s0 = 0; s0 < n_steps; s0 += max_steps) {
    a.bt0 = s0; a.n_btiles = std::min(max_steps, n_steps - s0);
    rc = tc_sym_launch(st, p, a, lo, len, tb, cand, cand_count, n_launches, n_aux_launches, stream);
    if (rc) return rc;
  }
  if (r == 0 && lo + len < T) {
    rc = tc_sym_refine(st, p, k, (double)len / (double)T, 0, std::min(n, len * TC_BN), cand, cand_count, bound, n_aux_launches, stream);
    if (rc) return rc;
  }
  // the later blocks, tile by tile (every pair of the launch walks the same tiles)
  a.sym_linear = 1;
  for (long long j0 = lo + len; j0 < T; j0 += max_steps) {
    a.bt0 = j0; a.n_btiles = std::min(max_steps, T - j0);
    rc = tc_sym_launch(st, p, a, lo, len, tb, cand, cand_count, n_launches, n_aux_launches, stream);
    if (rc) return rc;
  }
  return SBK_OK;
}

// ---------------------------------------------------------------------------
// sharded symmetric build
// ---------------------------------------------------------------------------
long long tc_shard_block_tiles(const TcPlan& p, int world) { return (p.n_tiles + world - 1) / world; }

__global__ void tc_bcast_thr_kernel(const float* __restrict__ thr, long long lo, long long hi, int world, TcMail ptrs) {
  // ptrs.out[r] reused as the float exchange array of rank r
  const long long s = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= hi) return;
  const float v = thr[s];
  for (int r = 0; r < world; ++r) reinterpret_cast<float*>(ptrs.out[r])[s] = v;
}

int tc_shard_bcast_thr(TcState* st, const TcPlan& p, int rank, int world, float* const* xthr, cudaStream_t stream) {
  const long long tb = tc_shard_block_tiles(p, world);
  const long long lo = std::min<long long>(st->n_ref, (long long)rank * tb * TC_BN), hi = std::min<long long>(st->n_ref, ((long long)rank + 1) * tb * TC_BN);
  if (hi <= lo) return SBK_OK;
  TcMail ptrs; memset(&ptrs, 0, sizeof(ptrs));
  for (int r = 0; r < world; ++r) ptrs.out[r] = reinterpret_cast<uint4*>(xthr[r]);
  tc_bcast_thr_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, stream>>>(st->thr, lo, hi, world, ptrs);
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}

__global__ void tc_mail_counts_kernel(const int* __restrict__ out_cnt, int rank, int world, TcMail ptrs) {
  // ptrs.out[r] reused as the int count_in array of rank r: count_in[source] = records this source sent
  const int r = threadIdx.x;
  if (r < world) reinterpret_cast<int*>(ptrs.out[r])[rank] = out_cnt[r];
}

// The tile products of one rank of the sharded symmetric build, as tile ranges.  kind 0: the own block [q_lo, q_lo + q_len),
// circulant inside itself (steps 0 .. q_len / 2; j_lo / j_len repeat the block); kind 1: query tiles [q_lo, q_lo + q_len)
// against every tile of [j_lo, j_lo + j_len).  Over all ranks every ordered pair of tiles is tested exactly once: a
// product (I, J) tests the rows of I against J and -- except on the diagonal and the opposite tiles of an even-length
// circulant block, which hold both orders themselves -- the rows of J against I (tests/test_host_logic.py walks this
// list for every world size through sbk_debug_sym_shard_schedule; tc_sym_shard_filter launches exactly this list).
std::vector<TcShardStep> tc_shard_schedule(long long T, int world, int rank) {
  std::vector<TcShardStep> steps;
  const long long tb = (T + world - 1) / world;
  const long long lo = (long long)rank * tb, len = std::min(tb, T - lo);
  if (len <= 0) return steps;
  steps.push_back({0, lo, len, lo, len});
  const int h_full = (world - 1) / 2;
  for (int s = 1; s <= h_full; ++s) {
    const long long c = (rank + s) % world, clo = c * tb, clen = std::min(tb, T - clo);
    if (clen > 0) steps.push_back({1, lo, len, clo, clen});
  }
  if ((world & 1) == 0) {
    // the opposite block pair is shared: the lower rank takes its rows x the first half of the other block's tiles, the
    // upper rank the second half of ITS tiles (as rows) x all tiles of the lower block
    const long long c = (rank + world / 2) % world, clo = c * tb, clen = std::max<long long>(0, std::min(tb, T - clo));
    if (rank < world / 2) { if (clen / 2 > 0) steps.push_back({1, lo, len, clo, clen / 2}); }
    else { const long long h = len / 2; if (len - h > 0 && clen > 0) steps.push_back({1, lo + h, len - h, clo, clen}); }
  }
  return steps;
}

int tc_sym_shard_filter(TcState* st, const TcPlan& p, int k, int rank, int world, uint2* cand, int32_t* cand_count, double* bound,
                        const TcMail& mail, int* const* peer_count_in, int* n_launches, int* n_aux_launches, cudaStream_t stream,
                        int (*mid_exchange)(void*, int), void* mid_arg) {
  if (!st->sym || st->cg != 2) { set_error("tensor kNN: state not prepared for the symmetric filter"); return SBK_ERR_INTERNAL; }
  const long long n = st->n_ref, T = p.n_tiles, tb = tc_shard_block_tiles(p, world);
  const long long lo = (long long)rank * tb, len = std::min(tb, T - lo);
  SBK_CUDA_CHECK(cudaMemsetAsync(mail.out_cnt, 0, 8 * sizeof(int), stream));
  int rc = SBK_OK;
  // (an error in the own block is carried into the mid-pass exchange, whose barrier every rank has to reach)
  auto own_block = [&]() -> int {
    TcArgs a;
    tc_sym_args(st, p, cand, cand_count, a);
    const long long max_steps = tc_sym_max_steps(p);
    // own block, circulant inside it, with the refinement points of the single-GPU schedule: after 1/14, 1/7 and 2/7 of
    // the block the own rows have seen a window of 1/7, 2/7, 4/7 of their block = (2s - 1) / T of all references (what
    // other ranks find for them waits in their outboxes and does not disturb the count)
    a.blk_lo = lo; a.blk_len = len; a.sym_linear = 0;
    const long long n_steps = len / 2 + 1;
    std::vector<long long> cuts; std::vector<char> refine_at;
    {
      const long long marks[3] = {(len + 13) / 14, (len + 6) / 7, (2 * len + 6) / 7};
      long long prev = 0;
      auto add_range = [&](long long end, bool is_mark) {
        if (end <= prev) return;
        const long long l2 = end - prev, parts = (l2 + max_steps - 1) / max_steps;
        for (long long q = 1; q <= parts; ++q) { cuts.push_back(prev + l2 * q / parts); refine_at.push_back(is_mark && q == parts); }
        prev = end;
      };
      for (int m = 0; m < 3; ++m) if (marks[m] < n_steps) add_range(marks[m], true);
      add_range(n_steps, false);
    }
    long long s0 = 0;
    for (size_t li = 0; li < cuts.size(); ++li) {
      a.bt0 = s0; a.n_btiles = cuts[li] - s0;
      int rc = tc_sym_launch(st, p, a, lo, len, tb, cand, cand_count, n_launches, n_aux_launches, stream, &mail);
      if (rc) return rc;
      s0 = cuts[li];
      if (refine_at[li] && s0 < n_steps) {
        rc = tc_sym_refine(st, p, k, std::min(1.0, (double)(2 * s0 - 1) / (double)T), lo * TC_BN, std::min(n, (lo + len) * TC_BN),
                           cand, cand_count, bound, n_aux_launches, stream);
        if (rc) return rc;
      }
    }
    // the own rows have now seen their whole block
    return tc_sym_refine(st, p, k, (double)len / (double)T, lo * TC_BN, std::min(n, (lo + len) * TC_BN), cand, cand_count, bound,
                         n_aux_launches, stream);
  };
  if (len > 0) rc = own_block();
  // refined thresholds of every rank before the blocks of other ranks are walked (every rank calls this, rows or not)
  if (mid_exchange) rc = mid_exchange(mid_arg, rc);
  if (rc) return rc;
  if (len > 0) {
    TcArgs a;
    tc_sym_args(st, p, cand, cand_count, a);
    const long long max_steps = tc_sym_max_steps(p);
    a.blk_lo = lo; a.blk_len = len;
    a.sym_linear = 1;
    auto linear = [&](long long q_lo, long long q_len, long long j_lo, long long j_len) -> int {
      for (long long j0 = j_lo; j0 < j_lo + j_len; j0 += max_steps) {
        a.bt0 = j0; a.n_btiles = std::min(max_steps, j_lo + j_len - j0);
        int r2 = tc_sym_launch(st, p, a, q_lo, q_len, tb, cand, cand_count, n_launches, n_aux_launches, stream, &mail);
        if (r2) return r2;
      }
      return SBK_OK;
    };
    for (const TcShardStep& sp : tc_shard_schedule(T, world, rank)) {
      if (sp.kind != 1) continue;                    // (kind 0, the own block, ran before the exchange)
      rc = linear(sp.q_lo, sp.q_len, sp.j_lo, sp.j_len);
      if (rc) return rc;
    }
  }
  // how many records went to every rank (into their count_in[this rank])
  TcMail ptrs; memset(&ptrs, 0, sizeof(ptrs));
  for (int r = 0; r < world; ++r) ptrs.out[r] = reinterpret_cast<uint4*>(peer_count_in[r]);
  tc_mail_counts_kernel<<<1, 32, 0, stream>>>(mail.out_cnt, rank, world, ptrs);
  SBK_LAUNCH_CHECK();
  return tc_sym_end(st, stream);
}

__global__ void __launch_bounds__(256)
tc_mail_merge_kernel(const TcMail boxes, const int* __restrict__ count_in, int world, int cap,
                     const float4* __restrict__ tq, uint2* __restrict__ cand, int32_t* __restrict__ cand_count, int cand_cap,
                     int* __restrict__ overflow_flag) {
  const int seg_cap = cand_cap >> 2;
  for (int src = 0; src < world; ++src) {
    const int cnt = count_in[src];
    if (cnt > cap) { if (blockIdx.x == 0 && threadIdx.x == 0) *overflow_flag = 1; }
    const uint4* box = boxes.out[src];                           // in the SOURCE rank's memory: coalesced peer loads
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < min(cnt, cap); e += (long long)gridDim.x * blockDim.x) {
      const uint4 rec = box[e];
      const long long j = rec.x;
      const float sv = __uint_as_float(rec.z);
      const float4 qj = tq[j];
      if (sv <= qj.z) {                                          // this rank's (refined) threshold
        const int ts = (int)(e & 3);
        const int at = atomicAdd(&cand_count[j * 4 + ts], 1);
        if (at < seg_cap) cand[j * (long long)cand_cap + (long long)ts * seg_cap + at] = make_uint2(rec.y, __float_as_uint(__fsub_rd(sv, qj.y)));
      }
    }
  }
}

int tc_mail_merge(TcState* st, const TcPlan& p, uint4* const* boxes, const int* count_in, int world, int cap,
                  uint2* cand, int32_t* cand_count, int* overflow_flag, cudaStream_t stream) {
  TcMail b; memset(&b, 0, sizeof(b));
  for (int r = 0; r < world; ++r) b.out[r] = boxes[r];
  tc_mail_merge_kernel<<<148 * 8, 256, 0, stream>>>(b, count_in, world, cap, st->tq, cand, cand_count, p.cand_cap, overflow_flag);
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}

int tc_dump_scores(TcState* st, const TcPlan& p, long long n_q, float* out, long long out_ld, cudaStream_t stream) {
  TcArgs a; memset(&a, 0, sizeof(a));
  a.a_tiles = st->a_tiles; a.b_tiles = st->b_tiles; a.n_btiles = p.n_tiles;
  a.kpad = p.kpad; a.nstage = st->nstage;
  a.tile_m0 = 0; a.row_lo = 0; a.row_hi = n_q;
  a.dump = out; a.dump_ld = out_ld; a.orig = st->orig;
  return tc_launch<TC_DUMP>(a, (n_q + TC_BM - 1) / TC_BM, st->cg, stream);
}

__global__ void tc_debug_norms_kernel(const float4* norms4, const double* nrm, const double* scale,
                                      const TcScalars* sc, int kpad, long long n, float* norms4_out,
                                      double* nrm_out, int32_t* info) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = norms4[i];
  norms4_out[4 * i] = v.x; norms4_out[4 * i + 1] = v.y; norms4_out[4 * i + 2] = v.z; norms4_out[4 * i + 3] = v.w;
  nrm_out[i] = nrm[i];
  if (i == 0) { nrm_out[n] = scale[0]; info[0] = kpad; info[1] = sc->n_split; }
}

int tc_debug_norms(TcState* st, const TcPlan& p, long long n, float* norms4_out, double* nrm_out, int32_t* info,
                   cudaStream_t stream) {
  tc_debug_norms_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(st->norms4, st->nrm, st->scale, st->sc, p.kpad, n,
                                                                        norms4_out, nrm_out, info);
  SBK_LAUNCH_CHECK();
  return SBK_OK;
}

}  // namespace sbk
