// api.cu -- extern "C" entry points of libsbk.so (see include/sbk.h) and the
// host-side orchestration of the kNN pipeline.
#include "common.cuh"
#include "knn.cuh"
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace sbk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------
// kNN orchestration
// ---------------------------------------------------------------------------
static constexpr int KNN_MAX_K = 1023;
static constexpr long long EXACT_CHUNK = 1 << 18;       // query rows per chunk (exact path)
static constexpr long long TC_MIN_N = 16384;            // AUTO switches to the tensor path above this
static constexpr double SYM_ROOMY_LIST_BYTES = 56e9;     // symmetric build: candidate lists may take this much when doubled
static constexpr int SYM_SINK_BLOCKS = 8;               // block rows of the symmetric build when the result streams to host memory

// Profiling switches exist only in the probe build (-DSBK_ENABLE_PROBES, libsbk_probe.so): the product library
// never reads the environment, so a stray variable cannot change or corrupt a result.
#ifdef SBK_ENABLE_PROBES
static int probe_mode() { static const int v = getenv("SBK_TC_PROBE") ? atoi(getenv("SBK_TC_PROBE")) : 0; return v; }
static bool probe_block_rerank() { static const int v = getenv("SBK_RERANK_BLOCK") ? atoi(getenv("SBK_RERANK_BLOCK")) : 0; return v != 0; }
static bool probe_no_row_order() { static const int v = getenv("SBK_NO_ROW_ORDER") ? atoi(getenv("SBK_NO_ROW_ORDER")) : 0; return v != 0; }
#else
static constexpr bool probe_block_rerank() { return false; }
static constexpr bool probe_no_row_order() { return false; }
#endif

struct KnnWs {
  uint32_t* cand_idx; uint2* cand_pair; int32_t* cand_count; double* bound;   // exact path: columns; tensor path: (column, score)
  int32_t* fail_rows; int* n_fail; unsigned long long* cand_total;
  uint32_t* fb_idx; int32_t* fb_count;                  // fallback (exact filter) candidate lists
  int32_t* retry;                                       // tensor path: rows handed back by the warp rerank
  void* order_ws; size_t order_bytes;                   // tensor path: locality order of the chunk's rows for the warp rerank
  void* tc;                                             // tensor-path staging area
  size_t tc_bytes;
};

struct KnnShape {
  int algo; long long chunk; long long cand_stride; int n_split; long long fb_rows; int fb_split;
  int tcap; long long retry_cap;                        // warp rerank: sort capacity per row (0 = block rerank only), retry list length
  bool sym;                                             // symmetric self-join filter (all rows in one call, lists of every row resident)
  TcPlan plan;
};

static int exact_split(long long n_q, long long n_ref) {
  long long ctas = (n_q + 7) / 8;
  int s = 1;
  if (ctas < 296) s = (int)std::min<long long>(8, (296 + ctas - 1) / ctas);
  long long max_by_ref = std::max<long long>(1, n_ref / 2048);
  return (int)std::min<long long>(s, max_by_ref);
}

static KnnShape knn_shape(int dtype, long long n_ref, int d, long long n_q, int k, int algo) {
  KnnShape s; memset(&s, 0, sizeof(s));
  const int k1 = k + 1;
  if (algo == SBK_KNN_AUTO) algo = (n_ref >= TC_MIN_N && (d + 5 + 15) / 16 * 16 <= 176) ? SBK_KNN_TENSOR : SBK_KNN_EXACT;
  const bool starved = algo == SBK_KNN_TENSOR_SYMMETRIC_STARVED;
  const bool onesided = algo == SBK_KNN_TENSOR_ONESIDED, force_sym = algo == SBK_KNN_TENSOR_SYMMETRIC || starved;
  if (onesided || force_sym) algo = SBK_KNN_TENSOR;
  s.algo = algo;
  if (algo == SBK_KNN_EXACT) {
    s.chunk = std::min<long long>(std::max<long long>(n_q, 1), EXACT_CHUNK);
    s.n_split = exact_split(s.chunk, n_ref);
    s.cand_stride = (long long)s.n_split * k1;
    s.fb_rows = 0; s.fb_split = 1;
  } else {
    s.plan = tc_make_plan(n_ref, d, k);
    s.chunk = std::min<long long>(std::max<long long>(n_q, 1), tc_chunk_rows(s.plan, n_ref));
    s.cand_stride = s.plan.cand_cap;
    s.fb_rows = std::min<long long>(s.chunk, 8192);   // fallback rows handled per round
    s.fb_split = 32;                                  // few rows x all references: parallelise over references
    if (n_ref / 2048 < s.fb_split) s.fb_split = (int)std::max<long long>(1, n_ref / 2048);
    while (s.fb_split > 1 && (size_t)s.fb_split * k1 * 12 > 150 * 1024) s.fb_split /= 2;
    s.tcap = rerank_warp_supported(dtype, d, k) ? rerank_warp_cap(k) : 0;
    // a build of all rows: every tile product is used for both of its row sets (half the tensor work); the candidate
    // lists of all rows are then live at once, so the whole call is one filter "chunk"
    // (below ~400k rows its extra launches, record transposes and refinements cost what the halved tensor work saves)
    s.sym = !onesided && n_q == n_ref && s.tcap > 0 && tc_sym_supported(s.plan, n_ref) && (force_sym || s.plan.n_tiles >= TC_SYM_MIN_TILES);
    if (s.sym) {
      s.chunk = n_q;
      if (starved) s.plan.pool_limit = 48;          // test hook: the record pool runs dry in the first launch
      // Lists of all rows are resident anyway: give every row twice the planned room when that still fits.  Rows of
      // genomes with thousands of contigs (every cluster mate passes the threshold) then go through the block rerank
      // instead of overflowing into the FP64 fallback (2 011 rows = 70 ms in the large-genome mix at 1M).
      const int cap2 = 2 * s.plan.cand_cap;
      if ((double)n_ref * (double)cap2 * 8.0 <= SYM_ROOMY_LIST_BYTES && cap2 <= 8192) s.plan.cand_cap = cap2;
      s.cand_stride = s.plan.cand_cap;
    }
    s.retry_cap = s.tcap ? std::max<long long>(8192, s.chunk / 4) : 0;
  }
  return s;
}

static void carve(Arena& a, const KnnShape& s, long long n_ref, long long n_q, int k, KnnWs& w) {
  const int k1 = k + 1;
  w.cand_idx = s.algo == SBK_KNN_TENSOR ? nullptr : a.take<uint32_t>((size_t)s.chunk * s.cand_stride);
  w.cand_pair = s.algo == SBK_KNN_TENSOR ? a.take<uint2>((size_t)s.chunk * s.cand_stride) : nullptr;
  w.cand_count = a.take<int32_t>((size_t)s.chunk * 4);
  w.bound = a.take<double>((size_t)s.chunk + TC_BN);     // + padding positions of the last tile (symmetric filter)
  w.fail_rows = a.take<int32_t>((size_t)n_q + (size_t)s.fb_rows + 1);     // all failing rows of the call, then the fallback's own sink
  w.n_fail = a.take<int>(64);
  w.cand_total = a.take<unsigned long long>(16);  // [0] candidates emitted, [1] reranked in FP64, [2..5] failure classes; [8..] same, retry launches
  w.fb_idx = a.take<uint32_t>((size_t)s.fb_rows * s.fb_split * k1);
  w.fb_count = a.take<int32_t>((size_t)s.fb_rows + 1);
  w.retry = a.take<int32_t>((size_t)s.retry_cap + 1);
  w.order_bytes = s.tcap ? row_order_workspace_bytes(s.chunk) : 0;
  w.order_ws = a.take<char>(w.order_bytes);
  w.tc = nullptr; w.tc_bytes = 0;
  if (s.algo == SBK_KNN_TENSOR) {
    w.tc_bytes = tc_workspace_bytes(s.plan, n_ref, s.chunk, s.sym);
    w.tc = a.take<char>(w.tc_bytes);
  }
}

// CUDA-event stage timer (only armed when the caller asked for stats).
struct StageTimer {
  struct Span { cudaEvent_t a, b; float* acc; };
  std::vector<Span> spans;
  cudaStream_t stream; bool on;
  StageTimer(cudaStream_t s, bool enable) : stream(s), on(enable) {}
  int begin(float* acc) {
    if (!on) return -1;
    Span sp; sp.acc = acc;
    if (cudaEventCreate(&sp.a) != cudaSuccess || cudaEventCreate(&sp.b) != cudaSuccess) { on = false; return -1; }
    cudaEventRecord(sp.a, stream);
    spans.push_back(sp);
    return (int)spans.size() - 1;
  }
  void end(int h) { if (h >= 0) cudaEventRecord(spans[h].b, stream); }
  void collect() {          // call after the stream is synchronised
    for (auto& sp : spans) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) *sp.acc += ms;
      cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);
    }
    spans.clear();
  }
  ~StageTimer() { for (auto& sp : spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); } }
};

}  // namespace sbk

using namespace sbk;

#ifndef SBK_SOURCE_HASH_STR
#define SBK_SOURCE_HASH_STR "SBK_SOURCE_HASH=unknown"
#endif
// "SBK_SOURCE_HASH=<sha256 prefix of the sources this binary was built from>" (semibin_b200/_build.py)
static const char g_source_hash[] = SBK_SOURCE_HASH_STR;

extern "C" int sbk_version(void) { return SBK_VERSION; }
extern "C" const char* sbk_source_hash(void) { return g_source_hash + 16; }
extern "C" int sbk_probes_enabled(void) {
#ifdef SBK_ENABLE_PROBES
  return 1;
#else
  return 0;
#endif
}
extern "C" const char* sbk_last_error(void) { return g_err; }

extern "C" int sbk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static int knn_check(int dtype, int64_t n_ref, int32_t d, int64_t q_begin, int64_t q_end, int32_t k) {
  if (dtype != SBK_F32 && dtype != SBK_F64) { set_error("knn: dtype %d not supported", dtype); return SBK_ERR_INVALID; }
  if (n_ref < 2 || n_ref >= (1LL << 31) - 1) { set_error("knn: n_ref=%lld outside [2, 2^31)", (long long)n_ref); return SBK_ERR_INVALID; }
  if (d < 1 || d > 4096) { set_error("knn: d=%d outside [1,4096]", d); return SBK_ERR_INVALID; }
  if (k < 1 || k >= n_ref) {
    // sklearn/neighbors/_base.py:840-851
    set_error("Expected n_neighbors < n_samples_fit, but n_neighbors = %d, n_samples_fit = %lld", k, (long long)n_ref);
    return SBK_ERR_INVALID;
  }
  if (k > KNN_MAX_K) { set_error("knn: n_neighbors=%d > %d not supported", k, KNN_MAX_K); return SBK_ERR_INVALID; }
  if (q_begin < 0 || q_end < q_begin || q_end > n_ref) { set_error("knn: query block [%lld,%lld) outside [0,%lld)", (long long)q_begin, (long long)q_end, (long long)n_ref); return SBK_ERR_INVALID; }
  return SBK_OK;
}

extern "C" size_t sbk_knn_workspace_bytes(int dtype, int64_t n_ref, int32_t d, int64_t n_query, int32_t k,
                                          int algo) {
  if (knn_check(dtype, n_ref, d, 0, std::min<int64_t>(n_query, n_ref), k) != SBK_OK) return 0;
  KnnShape s = knn_shape(dtype, n_ref, d, n_query, k, algo);
  Arena a(nullptr, 0);
  KnnWs w;
  carve(a, s, n_ref, std::min<int64_t>(n_query, n_ref), k, w);
  size_t need = a.off;
  if (s.sym) {                                   // the symmetric build may fall back to the one-sided one in the same workspace
    KnnShape s1 = knn_shape(dtype, n_ref, d, n_query, k, SBK_KNN_TENSOR_ONESIDED);
    Arena a1(nullptr, 0);
    carve(a1, s1, n_ref, std::min<int64_t>(n_query, n_ref), k, w);
    need = std::max(need, a1.off);
  }
  return need + 1024;
}

// Host sink of sbk_knn_stream / sbk_knn_graph: finished query chunks are copied to host memory on `copy_stream`
// while the next chunk is computed.  wide = 1 delivers scikit-learn's CSR arrays (float64 data, int64 indices,
// neighbors/_base.py:1031-1041): a widening kernel on the copy stream fills a device staging buffer ahead of each
// download, so the host never converts anything.
struct HostSink {
  void* dist; void* idx;                 // host destinations (dist may be NULL: mode='connectivity')
  long long chunk_rows; cudaStream_t copy_stream;
  int wide; double* st_dist; long long* st_idx; long long st_rows;
};

template <typename T>
__global__ void __launch_bounds__(256)
widen_kernel(const T* __restrict__ d, const int32_t* __restrict__ i, long long n, double* __restrict__ od,
             long long* __restrict__ oi) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    if (od) od[t] = (double)d[t];
    oi[t] = (long long)i[t];
  }
}

// Finished rows of the symmetric block-row build, straight into (mapped, pinned) host memory: one warp per row at a
// time, the row's k entries as consecutive 4/8-byte stores (zero-copy over PCIe; WIDE: scikit-learn's float64 / int64
// CSR dtypes).  `rows` lists row ids in any order; the device result (d, i) and the host arrays are in row order.
// The kernel runs NEXT TO the filter of the following block rows and is bound by PCIe for milliseconds.  A filter CTA
// owns its SM (218 KB of shared memory, 61 440 registers), so every scatter block takes an SM -- and with it the CTA
// pair of its TPC -- away from the filter for as long as it lives.  Measured at 1M (end to end / filter stage, ms):
// 148 blocks of 128 threads 227 / 173, 64: 210 / 156, 16: 201 / 144, and with 512-thread blocks 12: 198 / 143,
// 8: 197 / 141, 4: 193 / 137, 2: 204 / 135 (two SMs no longer keep up with the block rows).  Four blocks it is.
static constexpr int SCATTER_BLOCKS = 4;
static constexpr int SCATTER_THREADS = 512;
template <typename T, bool WIDE>
__global__ void __launch_bounds__(SCATTER_THREADS)
scatter_rows_kernel(const int32_t* __restrict__ rows, long long n_rows, int k, const T* __restrict__ d,
                    const int32_t* __restrict__ i, void* __restrict__ host_d, void* __restrict__ host_i) {
  const int wpb = SCATTER_THREADS / 32;
  for (long long w = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); w < n_rows; w += (long long)gridDim.x * wpb) {
    const long long o = (long long)rows[w] * k;
    for (int c = threadIdx.x & 31; c < k; c += 32) {
      if (WIDE) {
        if (host_d) reinterpret_cast<double*>(host_d)[o + c] = (double)d[o + c];
        reinterpret_cast<long long*>(host_i)[o + c] = (long long)i[o + c];
      } else {
        if (host_d) reinterpret_cast<T*>(host_d)[o + c] = d[o + c];
        reinterpret_cast<int32_t*>(host_i)[o + c] = i[o + c];
      }
    }
  }
}

// rows [r0, r0 + n_rows) of the device result (od / oi point at row r0) -> host rows [r0, ...); enqueued on the copy stream
static int sink_rows(const HostSink* sk, int dtype, int k, const void* od, const int32_t* oi, long long r0, long long n_rows) {
  const size_t esz = dtype == SBK_F32 ? 4 : 8;
  if (!sk->wide) {
    if (sk->dist) SBK_CUDA_CHECK(cudaMemcpyAsync((char*)sk->dist + (size_t)r0 * k * esz, od, (size_t)n_rows * k * esz, cudaMemcpyDeviceToHost, sk->copy_stream));
    SBK_CUDA_CHECK(cudaMemcpyAsync((int32_t*)sk->idx + (size_t)r0 * k, oi, (size_t)n_rows * k * sizeof(int32_t), cudaMemcpyDeviceToHost, sk->copy_stream));
    return SBK_OK;
  }
  for (long long p0 = 0; p0 < n_rows; p0 += sk->st_rows) {
    const long long pn = std::min(sk->st_rows, n_rows - p0);
    const long long cells = pn * k;
    const unsigned blocks = (unsigned)std::min<long long>((cells + 255) / 256, 148 * 8);
    const char* pd = (const char*)od + (size_t)p0 * k * esz;
    const int32_t* pi = oi + (size_t)p0 * k;
    const bool need_d = sk->dist != nullptr && dtype == SBK_F32;      // float64 distances are downloaded as they are
    if (dtype == SBK_F32) widen_kernel<float><<<blocks, 256, 0, sk->copy_stream>>>((const float*)pd, pi, cells, need_d ? sk->st_dist : nullptr, sk->st_idx);
    else widen_kernel<double><<<blocks, 256, 0, sk->copy_stream>>>((const double*)pd, pi, cells, nullptr, sk->st_idx);
    SBK_LAUNCH_CHECK();
    if (sk->dist) SBK_CUDA_CHECK(cudaMemcpyAsync((double*)sk->dist + (size_t)(r0 + p0) * k, need_d ? (const void*)sk->st_dist : (const void*)pd,
                                                 (size_t)cells * sizeof(double), cudaMemcpyDeviceToHost, sk->copy_stream));
    SBK_CUDA_CHECK(cudaMemcpyAsync((long long*)sk->idx + (size_t)(r0 + p0) * k, sk->st_idx, (size_t)cells * sizeof(long long), cudaMemcpyDeviceToHost, sk->copy_stream));
  }
  return SBK_OK;
}

// Sharded symmetric build (sbk_knn_sym_shard): this process is `rank` of `world` GPUs; xchg[r] = base of rank r's exchange
// buffer (peer memory; same layout everywhere: thresholds of all positions, count_in[8], one outbox per destination rank);
// barrier() returns once every rank has called it and the device work each rank enqueued before it is complete.
// It also carries an error vote: every rank passes "I failed" and gets "somebody failed", so that all ranks leave the
// build together (a rank that returned alone would leave the others waiting in the next barrier).
struct ShardCtx { int rank, world; void* const* xchg; int (*barrier)(void*, int); void* barrier_arg; };
static int shard_vote(const ShardCtx* sc, int rc) {
  const int any = sc->barrier(sc->barrier_arg, rc != SBK_OK);
  if (rc) return rc;
  if (any) { set_error("knn_sym_shard: another rank failed, the build was abandoned on all ranks"); return SBK_ERR_INTERNAL; }
  return SBK_OK;
}
static long long g_debug_mail_cap = 0;             // test hook (sbk_debug_shard_mail_cap): forces the overflow path
static long long shard_mail_cap(const TcPlan& p, long long n_ref, int world) {
  if (g_debug_mail_cap > 0) return g_debug_mail_cap;
  // one outbox holds what ONE rank finds for the rows of ONE other rank: at most one block pair, walked in one direction.
  // Sized as if every row of the build were at its planned list capacity (the positions are shuffled, so the blocks are
  // statistically alike): ~2.4x the synthetic load at 1M.  An overflow is detected by the receiver and fails the build
  // on all ranks (the host side then shards by row blocks).
  const long long c = n_ref * (long long)p.cand_cap_planned / ((long long)world * world);
  return std::max<long long>(c, 1 << 16);
}
static size_t shard_thr_bytes(const TcPlan& p) { return align_up((size_t)p.n_tiles * TC_BN * sizeof(float), 256); }

// every rank hands its (refined) thresholds to all ranks and takes theirs: broadcast by peer stores, barrier, local copy
struct ShardExchange { TcState* tcs; const TcPlan* plan; const ShardCtx* shard; float* const* xthr; long long n_q; cudaStream_t stream; };
static int shard_exchange_thr(void* arg, int rc) {
  ShardExchange* x = static_cast<ShardExchange*>(arg);
  if (!rc) rc = tc_shard_bcast_thr(x->tcs, *x->plan, x->shard->rank, x->shard->world, x->xthr, x->stream);
  if (!rc && cudaStreamSynchronize(x->stream) != cudaSuccess) { set_error("knn_sym_shard: stream synchronize failed"); rc = SBK_ERR_CUDA; }
  rc = shard_vote(x->shard, rc);
  if (rc) return rc;
  SBK_CUDA_CHECK(cudaMemcpyAsync(x->tcs->thr, x->xthr[x->shard->rank], (size_t)x->n_q * sizeof(float), cudaMemcpyDeviceToDevice, x->stream));
  return SBK_OK;
}

static int knn_impl(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld,
                    int64_t q_begin, int64_t q_end, int32_t k,
                    void* out_dist, int32_t* out_idx,
                    void* workspace, size_t ws_bytes, int algo,
                    sbk_knn_stats* stats, cudaStream_t stream, const HostSink* sink, const PeerOut* peers = nullptr,
                    const ShardCtx* shard = nullptr) {
  int rc = knn_check(dtype, n_ref, d, q_begin, q_end, k);
  if (rc) return rc;
  if (ld < d) { set_error("knn: ld=%lld < d=%d", (long long)ld, d); return SBK_ERR_INVALID; }
  if (sbk_device_count() < 1) { set_error("knn: no CUDA device (there is no CPU fallback)"); return SBK_ERR_CUDA; }
  const long long n_q = q_end - q_begin;
  const int k1 = k + 1;
  KnnShape s = knn_shape(dtype, n_ref, d, n_q, k, algo);
  Arena a(workspace, ws_bytes);
  KnnWs w;
  carve(a, s, n_ref, n_q, k, w);
  if (sink && sink->chunk_rows > 0)             // smaller chunks only: the workspace was sized for the default
    s.chunk = std::min<long long>(s.chunk, std::max<long long>(256, sink->chunk_rows / 256 * 256));
  if (!a.ok() || workspace == nullptr) { set_error("knn: workspace %zu B < required %zu B", ws_bytes, a.off); return SBK_ERR_WORKSPACE; }
  sbk_knn_stats st; memset(&st, 0, sizeof(st));
  st.algo_used = s.algo;
  const size_t esz = dtype == SBK_F32 ? 4 : 8;
  SBK_CUDA_CHECK(cudaMemsetAsync(w.cand_total, 0, 16 * sizeof(unsigned long long), stream));
  SBK_CUDA_CHECK(cudaMemsetAsync(w.n_fail, 0, 64 * sizeof(int), stream));      // [0] failing rows of the whole call

  StageTimer tm(stream, stats != nullptr);
  TcState tcs; memset(&tcs, 0, sizeof(tcs));
  bool blocks_done = false;                      // symmetric block-row build: every row already reranked and written out
  if (s.algo == SBK_KNN_TENSOR) {
    int h = tm.begin(&st.ms_prepare);
    rc = tc_prepare(X, dtype, n_ref, d, ld, s.plan, s.sym ? n_q : s.chunk, w.tc, w.tc_bytes, &tcs, stream, s.sym);
    tm.end(h);
    st.n_launches += 3;
    if (rc) return rc;
    tm.on = (stats != nullptr);
    st.sample_size = s.plan.sample_size; st.cand_capacity = s.plan.cand_cap;
    st.order_stat = s.plan.order_stat; st.kpad = s.plan.kpad; st.n_split_cols = s.plan.n_split;
    st.symmetric = s.sym ? 1 : 0;
    // host-buffer build: the block-row schedule finishes the rows block by block, each block is reranked and written to
    // the (mapped) host arrays while the later block rows are multiplied
    void* zc_dist = nullptr; void* zc_idx = nullptr;
    if (s.sym && sink) {
      const bool ok_i = cudaHostGetDevicePointer(&zc_idx, sink->idx, 0) == cudaSuccess;
      const bool ok_d = !sink->dist || cudaHostGetDevicePointer(&zc_dist, sink->dist, 0) == cudaSuccess;
      if (!ok_i || !ok_d) { cudaGetLastError(); zc_idx = nullptr; zc_dist = nullptr; }    // not mapped: chunked copies at the end
    }
    if (shard && !s.sym) { set_error("knn_sym_shard: the symmetric filter does not take this build (size / k / width)"); return SBK_ERR_INVALID; }
    if (s.sym && shard) {
      // ---- one process per GPU, every tile pair multiplied once across all of them ----
      const int rank = shard->rank, world = shard->world;
      const long long tb = tc_shard_block_tiles(s.plan, world);
      const long long slot_lo = std::min<long long>(n_q, (long long)rank * tb * TC_BN), slot_hi = std::min<long long>(n_q, ((long long)rank + 1) * tb * TC_BN);
      const long long cn = slot_hi - slot_lo;
      const size_t thr_bytes = shard_thr_bytes(s.plan);
      const long long mcap = shard_mail_cap(s.plan, n_ref, world);
      float* xthr[8]; int* xcount[8]; uint4* xmail[8];
      for (int r = 0; r < world; ++r) {
        xthr[r] = reinterpret_cast<float*>(shard->xchg[r]);
        xcount[r] = reinterpret_cast<int*>((char*)shard->xchg[r] + thr_bytes);
        xmail[r] = reinterpret_cast<uint4*>((char*)shard->xchg[r] + thr_bytes + 256);
      }
      // thresholds of the own rows, handed to every rank
      tc_trace("start", stream);
      h = tm.begin(&st.ms_sample);
      if (cn > 0) {
        TcState own = tcs; own.thr = tcs.thr + slot_lo;
        rc = tc_sample_chunk(&own, s.plan, slot_lo, cn, w.bound + slot_lo, stream);
      }
      ShardExchange xc; xc.tcs = &tcs; xc.plan = &s.plan; xc.shard = shard; xc.xthr = xthr; xc.n_q = n_q; xc.stream = stream;
      tc_trace("sample", stream);
      rc = shard_exchange_thr(&xc, rc);
      tm.end(h);
      if (rc) return rc;
      tc_trace("xchg1", stream);
      int nfl = 0, naux = 0;
      h = tm.begin(&st.ms_filter);
      rc = tc_sym_begin(&tcs, s.plan, w.cand_pair, w.cand_count, w.bound, &naux, stream);
      tc_trace("begin", stream);
      TcMail mail; memset(&mail, 0, sizeof(mail));
      for (int r = 0; r < world; ++r) mail.out[r] = xmail[rank] + (long long)r * mcap;      // own outbox for rank r (local memory)
      mail.out_cnt = w.n_fail + 8; mail.cap = (int)mcap; mail.blk_pos = tb * TC_BN; mail.own_lo = slot_lo; mail.own_hi = slot_hi;
      // (after its own block every rank refines its thresholds; they are exchanged once more so that the blocks of other
      // ranks are walked against refined column thresholds, and T8 / t are recomputed from them)
      struct Mid { ShardExchange* xc; TcState* tcs; const TcPlan* plan; cudaStream_t stream; bool left; } mid = {&xc, &tcs, &s.plan, stream, false};
      auto mid_fn = [](void* a, int err) -> int {
        Mid* m = static_cast<Mid*>(a);
        tc_trace("own_done", m->stream);
        int r2 = shard_exchange_thr(m->xc, err);
        tc_trace("xchg2", m->stream);
        m->left = (r2 != SBK_OK);          // the vote failed: every rank leaves here
        return r2 ? r2 : tc_refresh_colthr(m->tcs, *m->plan, m->stream);
      };
      if (rc) rc = mid_fn(&mid, rc);          // keeps the barrier count of the other ranks
      else rc = tc_sym_shard_filter(&tcs, s.plan, k, rank, world, w.cand_pair, w.cand_count, w.bound, mail, xcount, &nfl, &naux, stream,
                                        mid_fn, &mid);
      tm.end(h);
      if (rc == TC_SYM_POOL_EXHAUSTED) { set_error("knn_sym_shard: group-record pool exhausted"); rc = SBK_ERR_INTERNAL; }
      if (mid.left) return rc;                         // agreed at the mid-pass vote: every rank has left already
      tc_trace("filter_end", stream);
      rc = shard_vote(shard, rc);                      // every mailbox and count is complete (the filter ended with a stream sync)
      if (rc) return rc;
      h = tm.begin(&st.ms_filter);
      uint4* inbox[8];
      for (int r = 0; r < world; ++r) inbox[r] = xmail[r] + (long long)rank * mcap;          // rank r's outbox for this rank (peer memory)
      tc_trace("vote3", stream);
      rc = tc_mail_merge(&tcs, s.plan, inbox, xcount[rank], world, (int)mcap, w.cand_pair, w.cand_count, w.n_fail + 16, stream);
      tm.end(h);
      tc_trace("merge", stream);
      if (rc) return rc;
      st.n_filter_launches += nfl; st.n_launches += 4 + nfl + naux;
      if (cn > 0) {
        h = tm.begin(&st.ms_rerank);
        SBK_CUDA_CHECK(cudaMemsetAsync(w.n_fail + 2, 0, sizeof(int), stream));
        int32_t* perm = nullptr;
        rc = launch_row_order(X, dtype, d, ld, q_begin, cn, tcs.centre, w.order_ws, w.order_bytes, &perm, stream, tcs.orig + slot_lo);
        if (!rc) rc = launch_rerank_warp(X, dtype, d, ld, q_begin, cn, perm, tcs.rpos, w.cand_pair, w.cand_count, (int)s.cand_stride, tcs.thr, w.bound,
                                         tcs.nrm, tcs.norms4, tcs.scale, tcs.orig, k, out_dist, out_idx, q_begin, w.fail_rows, w.n_fail, w.retry,
                                         w.n_fail + 2, (int)s.retry_cap, w.cand_total, stream, peers);
        if (!rc) rc = launch_rerank(X, dtype, n_ref, d, ld, tcs.orig, q_begin, cn, nullptr, w.cand_pair, s.cand_stride, 4, w.cand_count,
                                    w.bound, tcs.nrm, tcs.norms4, tcs.scale, tcs.orig, k, out_dist, out_idx, q_begin, w.fail_rows, w.n_fail,
                                    w.cand_total + 8, stream, w.retry, w.n_fail + 2, (int)s.retry_cap, peers);
        tm.end(h);
        st.n_launches += 5;
        if (rc) return rc;
      }
      int h_over = 0;
      SBK_CUDA_CHECK(cudaMemcpyAsync(&h_over, w.n_fail + 16, sizeof(int), cudaMemcpyDeviceToHost, stream));
      SBK_CUDA_CHECK(cudaStreamSynchronize(stream));
      tc_trace("rerank", stream);
      tc_trace_dump(rank);
      if (h_over) { set_error("knn_sym_shard: a mailbox overflowed (more cross-GPU candidates than planned)"); return SBK_ERR_INTERNAL; }
      st.n_chunks++;
      blocks_done = true;
    } else if (s.sym && zc_idx) {
      h = tm.begin(&st.ms_sample);
      rc = tc_sample_chunk(&tcs, s.plan, 0, n_q, w.bound, stream);
      tm.end(h);
      if (rc) return rc;
      int nfl = 0, naux = 0;
      h = tm.begin(&st.ms_filter);
      rc = tc_sym_begin(&tcs, s.plan, w.cand_pair, w.cand_count, w.bound, &naux, stream);
      tm.end(h);
      if (rc) return rc;
      const int n_blocks = SYM_SINK_BLOCKS;
      const long long tb = tc_sym_block_tiles(s.plan, n_blocks);
      std::vector<cudaEvent_t> bev;
      auto drop_bev = [&]() { for (cudaEvent_t e : bev) cudaEventDestroy(e); bev.clear(); };
      for (int r = 0; r < n_blocks && (long long)r * tb < s.plan.n_tiles; ++r) {
        h = tm.begin(&st.ms_filter);
        rc = tc_sym_block(&tcs, s.plan, k, w.cand_pair, w.cand_count, w.bound, n_blocks, r, &nfl, &naux, stream);
        tm.end(h);
        if (rc) { drop_bev(); return rc; }
        const long long slot_lo = (long long)r * tb * TC_BN, slot_hi = std::min<long long>(n_q, ((long long)r + 1) * tb * TC_BN);
        const long long cn = slot_hi - slot_lo;
        if (cn <= 0) continue;
        h = tm.begin(&st.ms_rerank);
        SBK_CUDA_CHECK(cudaMemsetAsync(w.n_fail + 2, 0, sizeof(int), stream));
        int32_t* perm = nullptr;                             // the block's rows (orig[positions]) in locality order
        rc = launch_row_order(X, dtype, d, ld, q_begin, cn, tcs.centre, w.order_ws, w.order_bytes, &perm, stream, tcs.orig + slot_lo);
        if (!rc) rc = launch_rerank_warp(X, dtype, d, ld, q_begin, cn, perm, tcs.rpos, w.cand_pair, w.cand_count, (int)s.cand_stride, tcs.thr, w.bound,
                                         tcs.nrm, tcs.norms4, tcs.scale, tcs.orig, k, out_dist, out_idx, q_begin, w.fail_rows, w.n_fail, w.retry,
                                         w.n_fail + 2, (int)s.retry_cap, w.cand_total, stream, peers);
        if (!rc) rc = launch_rerank(X, dtype, n_ref, d, ld, tcs.orig, q_begin, cn, nullptr, w.cand_pair, s.cand_stride, 4, w.cand_count,
                                    w.bound, tcs.nrm, tcs.norms4, tcs.scale, tcs.orig, k, out_dist, out_idx, q_begin, w.fail_rows, w.n_fail,
                                    w.cand_total + 8, stream, w.retry, w.n_fail + 2, (int)s.retry_cap, peers);
        tm.end(h);
        st.n_launches += 5;
        if (rc) { drop_bev(); return rc; }
        cudaEvent_t ev;
        SBK_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        bev.push_back(ev);
        SBK_CUDA_CHECK(cudaEventRecord(ev, stream));
        SBK_CUDA_CHECK(cudaStreamWaitEvent(sink->copy_stream, ev, 0));
        unsigned sg = (unsigned)std::min<long long>((cn + 15) / 16, SCATTER_BLOCKS);
#ifdef SBK_ENABLE_PROBES
        { static const int pg = getenv("SBK_SCATTER_GRID") ? atoi(getenv("SBK_SCATTER_GRID")) : -1; if (pg == 0) continue; if (pg > 0) sg = (unsigned)pg; }
#endif
        const int32_t* rl = tcs.orig + slot_lo;               // row ids (q_begin = 0 in a build of all rows)
        if (sink->wide) {
          if (dtype == SBK_F32) scatter_rows_kernel<float, true><<<sg, SCATTER_THREADS, 0, sink->copy_stream>>>(rl, cn, k, (const float*)out_dist, out_idx, zc_dist, zc_idx);
          else scatter_rows_kernel<double, true><<<sg, SCATTER_THREADS, 0, sink->copy_stream>>>(rl, cn, k, (const double*)out_dist, out_idx, zc_dist, zc_idx);
        } else {
          if (dtype == SBK_F32) scatter_rows_kernel<float, false><<<sg, SCATTER_THREADS, 0, sink->copy_stream>>>(rl, cn, k, (const float*)out_dist, out_idx, zc_dist, zc_idx);
          else scatter_rows_kernel<double, false><<<sg, SCATTER_THREADS, 0, sink->copy_stream>>>(rl, cn, k, (const double*)out_dist, out_idx, zc_dist, zc_idx);
        }
        SBK_LAUNCH_CHECK();
        st.n_launches += 1;
        st.n_chunks++;
      }
      rc = tc_sym_end(&tcs, stream);
      drop_bev();
      st.n_filter_launches += nfl; st.n_launches += 1 + nfl + naux;
      if (rc == TC_SYM_POOL_EXHAUSTED) {
        SBK_CUDA_CHECK(cudaStreamSynchronize(sink->copy_stream));
        return knn_impl(X, dtype, n_ref, d, ld, q_begin, q_end, k, out_dist, out_idx, workspace, ws_bytes, SBK_KNN_TENSOR_ONESIDED,
                        stats, stream, sink, peers);
      }
      if (rc) return rc;
      blocks_done = true;
    } else if (s.sym) {
      // thresholds of every row, then the symmetric filter over all tile pairs: the chunk loop below only reranks
      h = tm.begin(&st.ms_sample);
      rc = tc_sample_chunk(&tcs, s.plan, 0, n_q, w.bound, stream);
      tm.end(h);
      if (rc) return rc;
      h = tm.begin(&st.ms_filter);
      int nfl = 0, naux = 0;
      rc = tc_filter_sym(&tcs, s.plan, k, w.cand_pair, w.cand_count, w.bound, &nfl, &naux, stream);
      tm.end(h);
      if (rc == TC_SYM_POOL_EXHAUSTED)             // more pair records than the pool holds (pathological hub structure): one-sided build
        return knn_impl(X, dtype, n_ref, d, ld, q_begin, q_end, k, out_dist, out_idx, workspace, ws_bytes, SBK_KNN_TENSOR_ONESIDED,
                        stats, stream, sink, peers);
      if (rc) return rc;
      st.n_filter_launches += nfl; st.n_launches += 1 + nfl + naux;
    }
  }

  // The chunk loop never waits for the GPU: rows whose certificate fails are collected over the whole call
  // (their output rows stay untouched) and recomputed once at the end.
  std::vector<cudaEvent_t> events;
  auto drop_events = [&]() { for (cudaEvent_t e : events) cudaEventDestroy(e); events.clear(); };
  for (long long c0 = 0; c0 < n_q && !blocks_done; c0 += s.chunk) {
    const long long cn = std::min(s.chunk, n_q - c0);
    const long long row0 = q_begin + c0;
    void* od = (char*)out_dist + (size_t)c0 * k * esz;
    int32_t* oi = out_idx + (size_t)c0 * k;
    if (s.algo == SBK_KNN_EXACT) {
      int h = tm.begin(&st.ms_exact_filter);
      rc = launch_exact_filter(X, dtype, n_ref, d, ld, nullptr, row0, cn, s.n_split, k1,
                               w.cand_idx, s.cand_stride, w.cand_count, stream);
      tm.end(h);
      if (rc) { drop_events(); return rc; }
      h = tm.begin(&st.ms_rerank);
      rc = launch_rerank(X, dtype, n_ref, d, ld, nullptr, row0, cn, w.cand_idx, nullptr, s.cand_stride, 1, w.cand_count,
                         nullptr, nullptr, nullptr, nullptr, nullptr, k, od, oi, row0, w.fail_rows, w.n_fail, w.cand_total, stream, nullptr, nullptr, 0, peers);
      tm.end(h);
      st.n_launches += 2;
      if (rc) { drop_events(); return rc; }
    } else {
      int h = -1;
      if (!s.sym) {
      h = tm.begin(&st.ms_sample);
      rc = tc_sample_chunk(&tcs, s.plan, row0, cn, w.bound, stream);
      tm.end(h);
      if (rc) { drop_events(); return rc; }
      h = tm.begin(&st.ms_filter);
      int nfl = 0, nrf = 0;
      rc = tc_filter_chunk(&tcs, s.plan, row0, cn, k, w.cand_pair, w.cand_count, w.bound, &nfl, &nrf, stream);
      st.n_filter_launches += nfl; st.n_launches += nfl + nrf - 1;
#ifdef SBK_ENABLE_PROBES
      if (probe_mode() > 0) {                    // profiling build only: time the filter variants, results are meaningless
        tm.end(h); SBK_CUDA_CHECK(cudaStreamSynchronize(stream)); st.n_chunks++; continue;
      }
#endif
      tm.end(h);
      if (rc) { drop_events(); return rc; }
      }
      // symmetric build: lists / thresholds / bounds are indexed by staged position (rpos: row -> position, orig: back)
      const int32_t* row_slot = s.sym ? tcs.rpos : nullptr;
      const int32_t* slot_rows = s.sym ? tcs.orig : nullptr;
      h = tm.begin(&st.ms_rerank);
      if (s.tcap > 0 && (s.sym || !probe_block_rerank())) {
        // warp-per-row rerank straight from the candidate lists, then the block-per-row kernel for the few rows it
        // hands back (device-side list, no host sync)
        SBK_CUDA_CHECK(cudaMemsetAsync(w.n_fail + 2, 0, sizeof(int), stream));
        int32_t* perm = nullptr;
        if (!probe_no_row_order()) rc = launch_row_order(X, dtype, d, ld, row0, cn, tcs.centre, w.order_ws, w.order_bytes, &perm, stream);
        if (!rc) rc = launch_rerank_warp(X, dtype, d, ld, row0, cn, perm, row_slot, w.cand_pair, w.cand_count, (int)s.cand_stride, tcs.thr, w.bound, tcs.nrm, tcs.norms4,
                                tcs.scale, tcs.orig, k, od, oi, row0, w.fail_rows, w.n_fail, w.retry, w.n_fail + 2, (int)s.retry_cap,
                                w.cand_total, stream, peers);
        if (!rc) rc = launch_rerank(X, dtype, n_ref, d, ld, slot_rows, row0, cn, nullptr, w.cand_pair, s.cand_stride, 4, w.cand_count,
                                    w.bound, tcs.nrm, tcs.norms4, tcs.scale, tcs.orig, k, od, oi, row0, w.fail_rows, w.n_fail, w.cand_total + 8, stream,
                                    w.retry, w.n_fail + 2, (int)s.retry_cap, peers);
        st.n_launches += 5;
      } else {
        rc = launch_rerank(X, dtype, n_ref, d, ld, nullptr, row0, cn, nullptr, w.cand_pair, s.cand_stride, 4, w.cand_count,
                           w.bound, tcs.nrm, tcs.norms4, tcs.scale, tcs.orig, k, od, oi, row0, w.fail_rows, w.n_fail, w.cand_total, stream,
                           nullptr, nullptr, 0, peers);
        st.n_launches += 3;
      }
      tm.end(h);
      if (rc) { drop_events(); return rc; }
    }
    if (sink) {                                  // download this chunk while the next one is computed
      cudaEvent_t ev;
      SBK_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      events.push_back(ev);
      SBK_CUDA_CHECK(cudaEventRecord(ev, stream));
      SBK_CUDA_CHECK(cudaStreamWaitEvent(sink->copy_stream, ev, 0));
      rc = sink_rows(sink, dtype, k, od, oi, c0, cn);
      if (rc) { drop_events(); return rc; }
    }
    st.n_chunks++;
  }
  int h_fail = 0;
  SBK_CUDA_CHECK(cudaMemcpyAsync(&h_fail, w.n_fail, sizeof(int), cudaMemcpyDeviceToHost, stream));
  SBK_CUDA_CHECK(cudaStreamSynchronize(stream));
  std::vector<int32_t> h_rows;
  if (h_fail > 0) {
    if (s.algo == SBK_KNN_EXACT) {
      drop_events();
      set_error("knn: exact filter produced an incomplete candidate list for %d rows", h_fail);
      return SBK_ERR_INTERNAL;
    }
    st.n_fallback_rows += h_fail;
    // rows whose certificate failed: exact FP64 filter over all references
    int hfb = tm.begin(&st.ms_fallback);
    for (int f0 = 0; f0 < h_fail; f0 += (int)s.fb_rows) {
      st.n_launches += 2;
      const int fn = std::min<int>((int)s.fb_rows, h_fail - f0);
      rc = launch_exact_filter(X, dtype, n_ref, d, ld, w.fail_rows + f0, 0, fn, s.fb_split, k1,
                               w.fb_idx, (long long)s.fb_split * k1, w.fb_count, stream);
      if (rc) { drop_events(); return rc; }
      // n_fail[1] counts rows the fallback itself could not resolve (none are expected)
      rc = launch_rerank(X, dtype, n_ref, d, ld, w.fail_rows + f0, 0, fn, w.fb_idx, nullptr, (long long)s.fb_split * k1, 1,
                         w.fb_count, nullptr, nullptr, nullptr, nullptr, nullptr, k, out_dist, out_idx, q_begin,
                         w.fail_rows + n_q, w.n_fail + 1, nullptr, stream, nullptr, nullptr, 0, peers);
      if (rc) { drop_events(); return rc; }
    }
    tm.end(hfb);
    if (sink) {
      h_rows.resize((size_t)h_fail);
      SBK_CUDA_CHECK(cudaMemcpyAsync(h_rows.data(), w.fail_rows, (size_t)h_fail * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    }
  }
  unsigned long long h_total[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  int h_fb_fail = 0;
  SBK_CUDA_CHECK(cudaMemcpyAsync(h_total, w.cand_total, sizeof(h_total), cudaMemcpyDeviceToHost, stream));
  SBK_CUDA_CHECK(cudaMemcpyAsync(&h_fb_fail, w.n_fail + 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
  SBK_CUDA_CHECK(cudaStreamSynchronize(stream));
  if (sink) {
    // the recomputed rows replace what the chunk downloads carried for them (copies on one stream stay ordered)
    if (h_fail > (sink->wide ? 256 : 4096)) {
      rc = sink_rows(sink, dtype, k, out_dist, out_idx, 0, n_q);
      if (rc) { drop_events(); return rc; }
    } else {
      for (int f = 0; f < h_fail; ++f) {
        const size_t r = (size_t)(h_rows[f] - q_begin);
        rc = sink_rows(sink, dtype, k, (char*)out_dist + r * k * esz, out_idx + r * k, (long long)r, 1);
        if (rc) { drop_events(); return rc; }
      }
    }
    SBK_CUDA_CHECK(cudaStreamSynchronize(sink->copy_stream));
  }
  drop_events();
  if (h_fb_fail > 0) { set_error("knn: exact fallback left %d rows unresolved", h_fb_fail); return SBK_ERR_INTERNAL; }
  st.n_candidates = (long long)h_total[0];
  st.n_reranked = (long long)(h_total[1] + h_total[9]);
  st.n_fail_overflow = (long long)(h_total[2] + h_total[10]); st.n_fail_few = (long long)(h_total[3] + h_total[11]);
  st.n_fail_survivors = (long long)(h_total[4] + h_total[12]); st.n_fail_certificate = (long long)(h_total[5] + h_total[13]);
  st.n_retry_rows = (long long)h_total[6];
  tm.collect();
  if (stats) *stats = st;
  return SBK_OK;
}

extern "C" int sbk_knn(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld,
                       int64_t q_begin, int64_t q_end, int32_t k,
                       void* out_dist, int32_t* out_idx,
                       void* workspace, size_t ws_bytes, int algo,
                       sbk_knn_stats* stats, void* cuda_stream) {
  return knn_impl(X, dtype, n_ref, d, ld, q_begin, q_end, k, out_dist, out_idx, workspace, ws_bytes, algo, stats,
                  (cudaStream_t)cuda_stream, nullptr);
}

extern "C" int sbk_knn_peers(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld,
                             int64_t q_begin, int64_t q_end, int32_t k,
                             void* out_dist, int32_t* out_idx,
                             int32_t n_peers, void* const* peer_dist, int32_t* const* peer_idx,
                             void* workspace, size_t ws_bytes, int algo,
                             sbk_knn_stats* stats, void* cuda_stream) {
  if (n_peers < 0 || n_peers > 8 || (n_peers > 0 && (!peer_dist || !peer_idx))) { set_error("knn_peers: 0..8 peer buffers"); return SBK_ERR_INVALID; }
  PeerOut po; po.n = n_peers;
  for (int t = 0; t < n_peers; ++t) {
    if (!peer_dist[t] || !peer_idx[t]) { set_error("knn_peers: null peer buffer %d", t); return SBK_ERR_INVALID; }
    po.dist[t] = peer_dist[t]; po.idx[t] = peer_idx[t];
  }
  return knn_impl(X, dtype, n_ref, d, ld, q_begin, q_end, k, out_dist, out_idx, workspace, ws_bytes, algo, stats,
                  (cudaStream_t)cuda_stream, nullptr, &po);
}

extern "C" void sbk_debug_shard_mail_cap(int64_t records) { g_debug_mail_cap = records; }

// host only (no device needed): the tile products sbk_knn_sym_shard launches on `rank`, five int64 per step
extern "C" int sbk_debug_sym_shard_schedule(int64_t n_tiles, int32_t world, int32_t rank, int64_t* out, int32_t max_steps) {
  if (n_tiles < 1 || world < 1 || world > 8 || rank < 0 || rank >= world) { set_error("sym_shard_schedule: bad arguments"); return SBK_ERR_INVALID; }
  const std::vector<TcShardStep> steps = tc_shard_schedule(n_tiles, world, rank);
  for (size_t i = 0; i < steps.size() && (int)i < max_steps && out; ++i) {
    out[5 * i + 0] = steps[i].kind; out[5 * i + 1] = steps[i].q_lo; out[5 * i + 2] = steps[i].q_len;
    out[5 * i + 3] = steps[i].j_lo; out[5 * i + 4] = steps[i].j_len;
  }
  return (int)steps.size();
}

extern "C" size_t sbk_knn_sym_shard_xchg_bytes(int dtype, int64_t n_ref, int32_t d, int32_t k, int32_t world) {
  if (knn_check(dtype, n_ref, d, 0, n_ref, k) != SBK_OK || world < 2 || world > 8) return 0;
  KnnShape s = knn_shape(dtype, n_ref, d, n_ref, k, SBK_KNN_TENSOR_SYMMETRIC);
  if (!s.sym) return 0;
  return shard_thr_bytes(s.plan) + 256 + (size_t)world * (size_t)shard_mail_cap(s.plan, n_ref, world) * sizeof(uint4);
}

extern "C" int sbk_knn_sym_shard(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld, int32_t k,
                                 int32_t rank, int32_t world, void* const* xchg, int (*barrier)(void*, int), void* barrier_arg,
                                 void* out_dist, int32_t* out_idx, int32_t n_peers, void* const* peer_dist, int32_t* const* peer_idx,
                                 void* workspace, size_t ws_bytes, sbk_knn_stats* stats, void* cuda_stream) {
  if (world < 2 || world > 8 || rank < 0 || rank >= world || !xchg || !barrier) { set_error("knn_sym_shard: bad rank / world / exchange table"); return SBK_ERR_INVALID; }
  if (n_peers < 0 || n_peers > 8 || (n_peers > 0 && (!peer_dist || !peer_idx))) { set_error("knn_sym_shard: 0..8 peer buffers"); return SBK_ERR_INVALID; }
  PeerOut po; po.n = n_peers;
  for (int t = 0; t < n_peers; ++t) { po.dist[t] = peer_dist[t]; po.idx[t] = peer_idx[t]; }
  ShardCtx sc; sc.rank = rank; sc.world = world; sc.xchg = xchg; sc.barrier = barrier; sc.barrier_arg = barrier_arg;
  return knn_impl(X, dtype, n_ref, d, ld, 0, n_ref, k, out_dist, out_idx, workspace, ws_bytes, SBK_KNN_TENSOR_SYMMETRIC, stats,
                  (cudaStream_t)cuda_stream, nullptr, &po, &sc);
}

extern "C" int sbk_knn_stream(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld,
                              int64_t q_begin, int64_t q_end, int32_t k,
                              void* out_dist, int32_t* out_idx,
                              void* host_dist, int32_t* host_idx, int64_t chunk_rows,
                              void* workspace, size_t ws_bytes, int algo,
                              sbk_knn_stats* stats, void* cuda_stream, void* copy_stream) {
  if (!host_dist || !host_idx) { set_error("knn_stream: host output buffers required"); return SBK_ERR_INVALID; }
  if (copy_stream == cuda_stream) { set_error("knn_stream: the copy stream must differ from the compute stream"); return SBK_ERR_INVALID; }
  HostSink sink; memset(&sink, 0, sizeof(sink));
  sink.dist = host_dist; sink.idx = host_idx; sink.chunk_rows = chunk_rows; sink.copy_stream = (cudaStream_t)copy_stream;
  return knn_impl(X, dtype, n_ref, d, ld, q_begin, q_end, k, out_dist, out_idx, workspace, ws_bytes, algo, stats,
                  (cudaStream_t)cuda_stream, &sink);
}

// ---- sbk_knn_graph: the arrays of sklearn.neighbors.kneighbors_graph straight into host memory ----
static constexpr long long GRAPH_STAGE_ROWS = 1 << 17;
struct GraphWs { void* dev_dist; int32_t* dev_idx; double* st_dist; long long* st_idx; long long st_rows; void* knn_ws; size_t knn_bytes; };
static void graph_carve(Arena& a, int dtype, int64_t n_ref, int32_t d, int64_t n_q, int32_t k, int algo, GraphWs& g) {
  const size_t esz = dtype == SBK_F32 ? 4 : 8;
  g.dev_dist = a.take<char>((size_t)n_q * k * esz);
  g.dev_idx = a.take<int32_t>((size_t)n_q * k);
  g.st_rows = std::min<long long>(std::max<long long>(n_q, 1), GRAPH_STAGE_ROWS);
  g.st_dist = a.take<double>(dtype == SBK_F32 ? (size_t)g.st_rows * k : 0);
  g.st_idx = a.take<long long>((size_t)g.st_rows * k);
  g.knn_bytes = sbk_knn_workspace_bytes(dtype, n_ref, d, n_q, k, algo);
  g.knn_ws = a.take<char>(g.knn_bytes);
}

extern "C" size_t sbk_knn_graph_workspace_bytes(int dtype, int64_t n_ref, int32_t d, int64_t n_query, int32_t k, int algo) {
  if (knn_check(dtype, n_ref, d, 0, std::min<int64_t>(n_query, n_ref), k) != SBK_OK) return 0;
  Arena a(nullptr, 0);
  GraphWs g;
  graph_carve(a, dtype, n_ref, d, n_query, k, algo, g);
  return a.off + 1024;
}

extern "C" int sbk_knn_graph(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld,
                             int64_t q_begin, int64_t q_end, int32_t k,
                             double* host_data, int64_t* host_indices, int64_t chunk_rows,
                             void* workspace, size_t ws_bytes, int algo,
                             sbk_knn_stats* stats, void* cuda_stream, void* copy_stream) {
  int rc = knn_check(dtype, n_ref, d, q_begin, q_end, k);
  if (rc) return rc;
  if (!host_indices) { set_error("knn_graph: host indices buffer required"); return SBK_ERR_INVALID; }
  if (copy_stream == cuda_stream) { set_error("knn_graph: the copy stream must differ from the compute stream"); return SBK_ERR_INVALID; }
  Arena a(workspace, ws_bytes);
  GraphWs g;
  graph_carve(a, dtype, n_ref, d, q_end - q_begin, k, algo, g);
  if (!a.ok() || workspace == nullptr) { set_error("knn_graph: workspace %zu B < required %zu B", ws_bytes, a.off); return SBK_ERR_WORKSPACE; }
  HostSink sink; memset(&sink, 0, sizeof(sink));
  sink.dist = host_data; sink.idx = host_indices; sink.chunk_rows = chunk_rows; sink.copy_stream = (cudaStream_t)copy_stream;
  sink.wide = 1; sink.st_dist = g.st_dist; sink.st_idx = g.st_idx; sink.st_rows = g.st_rows;
  return knn_impl(X, dtype, n_ref, d, ld, q_begin, q_end, k, g.dev_dist, g.dev_idx, g.knn_ws, g.knn_bytes, algo, stats,
                  (cudaStream_t)cuda_stream, &sink);
}

extern "C" size_t sbk_debug_tc_workspace_bytes(int64_t n_ref, int32_t d) {
  TcPlan p = tc_make_plan(n_ref, d, 1);
  return tc_workspace_bytes(p, n_ref, TC_BM) + 1024;
}

extern "C" int sbk_debug_tc_scores(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld,
                                   int64_t n_q, float* out, float* norms4, double* nrm, int32_t* info,
                                   void* workspace, size_t ws_bytes, void* cuda_stream) {
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  if (sbk_device_count() < 1) { set_error("no CUDA device"); return SBK_ERR_CUDA; }
  if (n_q > n_ref || n_ref > 65536 || n_ref < 2) { set_error("debug scores: n_q <= n_ref <= 65536 required"); return SBK_ERR_INVALID; }
  TcPlan p = tc_make_plan(n_ref, d, 1);
  TcState st; memset(&st, 0, sizeof(st));
  int rc = tc_prepare(X, dtype, n_ref, d, ld, p, TC_BM, workspace, ws_bytes, &st, stream);
  if (rc) return rc;
  rc = tc_dump_scores(&st, p, n_q, out, p.n_tiles * TC_BN, stream);
  if (rc) return rc;
  rc = tc_debug_norms(&st, p, n_ref, norms4, nrm, info, stream);
  if (rc) return rc;
  SBK_CUDA_CHECK(cudaStreamSynchronize(stream));
  return SBK_OK;
}
