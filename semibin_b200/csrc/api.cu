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

// Profiling switches exist only in the probe build (-DSBK_ENABLE_PROBES, libsbk_probe.so): the product library
// never reads the environment, so a stray variable cannot change or corrupt a result.
#ifdef SBK_ENABLE_PROBES
static int probe_mode() { static const int v = getenv("SBK_TC_PROBE") ? atoi(getenv("SBK_TC_PROBE")) : 0; return v; }
static bool probe_block_rerank() { static const int v = getenv("SBK_RERANK_BLOCK") ? atoi(getenv("SBK_RERANK_BLOCK")) : 0; return v != 0; }
#else
static constexpr bool probe_block_rerank() { return false; }
#endif

struct KnnWs {
  uint32_t* cand_idx; uint2* cand_pair; int32_t* cand_count; double* bound;   // exact path: columns; tensor path: (column, score)
  int32_t* fail_rows; int* n_fail; unsigned long long* cand_total;
  uint32_t* fb_idx; int32_t* fb_count;                  // fallback (exact filter) candidate lists
  uint32_t* tlist; int32_t* tcount; int32_t* retry;     // tensor path: survivor lists of tc_select, rows handed back by the warp rerank
  void* tc;                                             // tensor-path staging area
  size_t tc_bytes;
};

struct KnnShape {
  int algo; long long chunk; long long cand_stride; int n_split; long long fb_rows; int fb_split;
  int tcap; long long retry_cap;                        // warp rerank: survivor capacity per row (0 = block rerank), retry list length
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
    s.retry_cap = s.tcap ? std::max<long long>(8192, s.chunk / 8) : 0;
  }
  return s;
}

static void carve(Arena& a, const KnnShape& s, long long n_ref, long long n_q, int k, KnnWs& w) {
  const int k1 = k + 1;
  w.cand_idx = s.algo == SBK_KNN_TENSOR ? nullptr : a.take<uint32_t>((size_t)s.chunk * s.cand_stride);
  w.cand_pair = s.algo == SBK_KNN_TENSOR ? a.take<uint2>((size_t)s.chunk * s.cand_stride) : nullptr;
  w.cand_count = a.take<int32_t>((size_t)s.chunk * 4);
  w.bound = a.take<double>((size_t)s.chunk);
  w.fail_rows = a.take<int32_t>((size_t)n_q + (size_t)s.fb_rows + 1);     // all failing rows of the call, then the fallback's own sink
  w.n_fail = a.take<int>(64);
  w.cand_total = a.take<unsigned long long>(16);  // [0] candidates emitted, [1] reranked in FP64, [2..5] failure classes; [8..] same, retry launches
  w.fb_idx = a.take<uint32_t>((size_t)s.fb_rows * s.fb_split * k1);
  w.fb_count = a.take<int32_t>((size_t)s.fb_rows + 1);
  w.tlist = a.take<uint32_t>((size_t)s.chunk * s.tcap);
  w.tcount = a.take<int32_t>((size_t)s.chunk + 1);
  w.retry = a.take<int32_t>((size_t)s.retry_cap + 1);
  w.tc = nullptr; w.tc_bytes = 0;
  if (s.algo == SBK_KNN_TENSOR) {
    w.tc_bytes = tc_workspace_bytes(s.plan, n_ref, s.chunk);
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
  return a.off + 1024;
}

// Host sink of sbk_knn_stream: finished query chunks are copied to host memory on `copy_stream` while the
// next chunk is computed.
struct HostSink { void* dist; int32_t* idx; long long chunk_rows; cudaStream_t copy_stream; };

static int knn_impl(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld,
                    int64_t q_begin, int64_t q_end, int32_t k,
                    void* out_dist, int32_t* out_idx,
                    void* workspace, size_t ws_bytes, int algo,
                    sbk_knn_stats* stats, cudaStream_t stream, const HostSink* sink) {
  int rc = knn_check(dtype, n_ref, d, q_begin, q_end, k);
  if (rc) return rc;
  if (ld < d) { set_error("knn: ld=%lld < d=%d", (long long)ld, d); return SBK_ERR_INVALID; }
  if (sbk_device_count() < 1) { set_error("knn: no CUDA device (there is no CPU fallback)"); return SBK_ERR_CUDA; }
  const long long n_q = q_end - q_begin;
  const int k1 = k + 1;
  KnnShape s = knn_shape(dtype, n_ref, d, n_q, k, algo);
  if (sink && sink->chunk_rows > 0)             // smaller chunks only: the workspace was sized for the default
    s.chunk = std::min<long long>(s.chunk, std::max<long long>(256, sink->chunk_rows / 256 * 256));
  Arena a(workspace, ws_bytes);
  KnnWs w;
  carve(a, s, n_ref, n_q, k, w);
  if (!a.ok() || workspace == nullptr) { set_error("knn: workspace %zu B < required %zu B", ws_bytes, a.off); return SBK_ERR_WORKSPACE; }
  sbk_knn_stats st; memset(&st, 0, sizeof(st));
  st.algo_used = s.algo;
  const size_t esz = dtype == SBK_F32 ? 4 : 8;
  SBK_CUDA_CHECK(cudaMemsetAsync(w.cand_total, 0, 16 * sizeof(unsigned long long), stream));
  SBK_CUDA_CHECK(cudaMemsetAsync(w.n_fail, 0, 64 * sizeof(int), stream));      // [0] failing rows of the whole call

  StageTimer tm(stream, stats != nullptr);
  TcState tcs; memset(&tcs, 0, sizeof(tcs));
  if (s.algo == SBK_KNN_TENSOR) {
    int h = tm.begin(&st.ms_prepare);
    rc = tc_prepare(X, dtype, n_ref, d, ld, s.plan, s.chunk, w.tc, w.tc_bytes, &tcs, stream);
    tm.end(h);
    st.n_launches += 3;
    if (rc) return rc;
    tm.on = (stats != nullptr);
    st.sample_size = s.plan.sample_size; st.cand_capacity = s.plan.cand_cap;
    st.order_stat = s.plan.order_stat; st.kpad = s.plan.kpad; st.n_split_cols = s.plan.n_split;
  }

  // The chunk loop never waits for the GPU: rows whose certificate fails are collected over the whole call
  // (their output rows stay untouched) and recomputed once at the end.
  std::vector<cudaEvent_t> events;
  auto drop_events = [&]() { for (cudaEvent_t e : events) cudaEventDestroy(e); events.clear(); };
  for (long long c0 = 0; c0 < n_q; c0 += s.chunk) {
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
                         nullptr, nullptr, nullptr, nullptr, nullptr, k, od, oi, row0, w.fail_rows, w.n_fail, w.cand_total, stream);
      tm.end(h);
      st.n_launches += 2;
      if (rc) { drop_events(); return rc; }
    } else {
      int h = tm.begin(&st.ms_sample);
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
      h = tm.begin(&st.ms_rerank);
      if (s.tcap > 0 && !probe_block_rerank()) {
        // final selection (thr / bound tightened from the complete lists) -> warp-per-row rerank of the survivors ->
        // block-per-row rerank of the few rows with more survivors than a warp sorts (device-side list, no host sync)
        SBK_CUDA_CHECK(cudaMemsetAsync(w.n_fail + 2, 0, sizeof(int), stream));
        rc = tc_select_chunk(&tcs, s.plan, row0, cn, k, w.cand_pair, w.cand_count, w.bound, w.tlist, w.tcount, s.tcap, w.cand_total, stream);
        if (!rc) rc = launch_rerank_warp(X, dtype, d, ld, row0, cn, w.tlist, w.tcount, s.tcap, w.bound, k, od, oi, row0,
                                         w.fail_rows, w.n_fail, w.retry, w.n_fail + 2, (int)s.retry_cap, w.cand_total, stream);
        if (!rc) rc = launch_rerank(X, dtype, n_ref, d, ld, nullptr, row0, cn, nullptr, w.cand_pair, s.cand_stride, 4, w.cand_count,
                                    w.bound, tcs.nrm, tcs.norms4, tcs.scale, tcs.orig, k, od, oi, row0, w.fail_rows, w.n_fail, w.cand_total + 8, stream,
                                    w.retry, w.n_fail + 2, (int)s.retry_cap);
        st.n_launches += 4;
      } else {
        rc = launch_rerank(X, dtype, n_ref, d, ld, nullptr, row0, cn, nullptr, w.cand_pair, s.cand_stride, 4, w.cand_count,
                           w.bound, tcs.nrm, tcs.norms4, tcs.scale, tcs.orig, k, od, oi, row0, w.fail_rows, w.n_fail, w.cand_total, stream);
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
      SBK_CUDA_CHECK(cudaMemcpyAsync((char*)sink->dist + (size_t)c0 * k * esz, od, (size_t)cn * k * esz, cudaMemcpyDeviceToHost, sink->copy_stream));
      SBK_CUDA_CHECK(cudaMemcpyAsync(sink->idx + (size_t)c0 * k, oi, (size_t)cn * k * sizeof(int32_t), cudaMemcpyDeviceToHost, sink->copy_stream));
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
                         w.fail_rows + n_q, w.n_fail + 1, nullptr, stream);
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
    if (h_fail > 4096) {
      SBK_CUDA_CHECK(cudaMemcpyAsync(sink->dist, out_dist, (size_t)n_q * k * esz, cudaMemcpyDeviceToHost, sink->copy_stream));
      SBK_CUDA_CHECK(cudaMemcpyAsync(sink->idx, out_idx, (size_t)n_q * k * sizeof(int32_t), cudaMemcpyDeviceToHost, sink->copy_stream));
    } else {
      for (int f = 0; f < h_fail; ++f) {
        const size_t r = (size_t)(h_rows[f] - q_begin);
        SBK_CUDA_CHECK(cudaMemcpyAsync((char*)sink->dist + r * k * esz, (char*)out_dist + r * k * esz, (size_t)k * esz, cudaMemcpyDeviceToHost, sink->copy_stream));
        SBK_CUDA_CHECK(cudaMemcpyAsync(sink->idx + r * k, out_idx + r * k, (size_t)k * sizeof(int32_t), cudaMemcpyDeviceToHost, sink->copy_stream));
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

extern "C" int sbk_knn_stream(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld,
                              int64_t q_begin, int64_t q_end, int32_t k,
                              void* out_dist, int32_t* out_idx,
                              void* host_dist, int32_t* host_idx, int64_t chunk_rows,
                              void* workspace, size_t ws_bytes, int algo,
                              sbk_knn_stats* stats, void* cuda_stream, void* copy_stream) {
  if (!host_dist || !host_idx) { set_error("knn_stream: host output buffers required"); return SBK_ERR_INVALID; }
  if (copy_stream == cuda_stream) { set_error("knn_stream: the copy stream must differ from the compute stream"); return SBK_ERR_INVALID; }
  HostSink sink; sink.dist = host_dist; sink.idx = host_idx; sink.chunk_rows = chunk_rows; sink.copy_stream = (cudaStream_t)copy_stream;
  return knn_impl(X, dtype, n_ref, d, ld, q_begin, q_end, k, out_dist, out_idx, workspace, ws_bytes, algo, stats,
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
