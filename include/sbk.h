/* sbk.h -- C-ABI of libsbk.so: the B200 (sm_100a) kernels behind SemiBin's
 * clustering front end.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Every entry point replaces one call site of the reference (paths relative to
 * the SemiBin source tree; sklearn/ paths relative to site-packages of the
 * scikit-learn 1.9.0 wheel the reference pins, pixi.lock:4152-4153):
 *
 *   sbk_knn                 sklearn.neighbors.kneighbors_graph(X, k, mode='distance', p=2)
 *                             SemiBin/cluster.py:94-99 (float32 embedding), :100-105 (float64
 *                             features), SemiBin/long_read_cluster.py:108-113
 *                             -> sklearn/neighbors/_base.py:756-1041, ArgKmin32/64
 *   sbk_edges_stage1        SemiBin/cluster.py:109-122  (kNN(emb) ∩ kNN(features) as a bit mask,
 *                             row max of 1-min(d,1), per-threshold counts for the max_node search)
 *   sbk_edges_stage2        SemiBin/cluster.py:127-151 + cal_kl :13-60 evaluated only at
 *                             surviving edges (threshold prune, KL factor, <=1e-6 cut,
 *                             upper triangle, row-major compaction)
 *   sbk_radius_sweep        sklearn.cluster.dbscan(..., metric='precomputed') neighbourhood
 *                             + core test, SemiBin/long_read_cluster.py:115-120
 *                             -> sklearn/cluster/_dbscan.py:430-463
 *   sbk_dbscan_labels       sklearn/cluster/_dbscan_inner.pyx (label expansion)
 *
 * Ownership: the caller allocates every buffer (inputs, outputs, workspace) on
 * the current CUDA device and keeps it alive until the stream is synchronised.
 * The library allocates nothing persistent.
 * Errors: functions return 0 on success or a negative sbk_status; the message
 * is available from sbk_last_error() (thread local).  Nothing throws across the
 * ABI.  There is no CPU fallback: without a CUDA device every compute entry
 * point returns SBK_ERR_CUDA.
 * Threading: re-entrant per (device, stream); the only global state is the
 * thread-local error slot.
 */
#ifndef SBK_H_
#define SBK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBK_VERSION 100          /* 0.1.0 */

typedef enum sbk_status {
  SBK_OK = 0,
  SBK_ERR_INVALID = -1,          /* bad argument (k >= n, unsupported width, ...) */
  SBK_ERR_WORKSPACE = -2,        /* workspace too small */
  SBK_ERR_CUDA = -3,             /* CUDA runtime error / no device */
  SBK_ERR_INTERNAL = -4          /* exactness certificate could not be established */
} sbk_status;

typedef enum sbk_dtype { SBK_F32 = 0, SBK_F64 = 1 } sbk_dtype;

typedef enum sbk_knn_algo {
  SBK_KNN_AUTO = 0,              /* tensor-core filter for large n, exact filter otherwise */
  SBK_KNN_EXACT = 1,             /* FP64 CUDA-core filter (also the certificate fallback) */
  SBK_KNN_TENSOR = 2,            /* tcgen05 FP16 filter + FP64 rerank with exactness certificate; a build of ALL rows
                                    (q_begin = 0, q_end = n_ref) of at least ~400k rows uses the symmetric self-join
                                    filter: every tile product serves both of its row sets, half the tensor work */
  SBK_KNN_TENSOR_ONESIDED = 3,   /* the same, but never the symmetric filter (row blocks always run one-sided) */
  SBK_KNN_TENSOR_SYMMETRIC = 4,  /* the same, symmetric filter whenever the call covers all rows (any size >= 2048 rows) */
  SBK_KNN_TENSOR_SYMMETRIC_STARVED = 5  /* test hook: symmetric filter with a record pool of 48 chunks, so that the
                                    pool-exhaustion path (the build is redone one-sided) can be exercised */
} sbk_knn_algo;

/* Filled (when non-NULL) after the call's work is complete; passing a non-NULL
 * pointer makes sbk_knn synchronise the stream before returning. */
typedef struct sbk_knn_stats {
  int32_t algo_used;             /* sbk_knn_algo actually run */
  int32_t n_chunks;              /* query chunks processed */
  int64_t n_fallback_rows;       /* rows whose certificate failed -> exact filter */
  int64_t n_candidates;          /* total candidates emitted by the filter */
  int32_t sample_size;           /* reference sample used for thresholds (tensor path) */
  int32_t cand_capacity;         /* per-row candidate slots (tensor path) */
  int32_t order_stat;            /* which group-min order statistic set the threshold */
  int32_t kpad;                  /* padded K of the FP16 operands */
  int32_t n_launches;            /* kernels launched by this call */
  int32_t n_filter_launches;     /* launches of the tcgen05 filter kernel (one per L2-sized reference chunk and query chunk) */
  /* device time per stage, CUDA events on `cuda_stream`, summed over chunks (ms) */
  float ms_prepare;              /* centre/scale/FP16 staging */
  float ms_sample;               /* tcgen05 threshold pass over the reference sample */
  float ms_filter;               /* tcgen05 filter pass over all references (dominant kernel) */
  float ms_rerank;               /* FP64 rerank + certificate */
  float ms_fallback;             /* exact filter + rerank of uncertified rows */
  float ms_exact_filter;         /* exact path: FP64 filter */
  int64_t n_reranked;            /* candidates that survived the score cut and got an FP64 distance */
  /* why rows went to the exact fallback (tensor path) */
  int64_t n_fail_overflow;       /* a candidate segment overflowed */
  int64_t n_fail_few;            /* fewer than k+1 candidates passed the threshold */
  int64_t n_fail_survivors;      /* more survivors of the score cut than the sort buffer holds */
  int64_t n_fail_certificate;    /* (k+1)-th exact distance not below the exclusion bound */
  int32_t n_split_cols;          /* high-energy columns given an FP16 hi+lo representation */
  int32_t symmetric;             /* 1 when the symmetric self-join filter produced the candidate lists */
  int64_t n_retry_rows;          /* rows with more survivors than the warp rerank sorts -> block-per-row rerank */
} sbk_knn_stats;

int sbk_version(void);
/* sha256 prefix (32 hex digits) of the sources this binary was compiled from; the Python binding recomputes it
 * from the tree and refuses a stale binary. */
const char* sbk_source_hash(void);
/* 1 in the profiling build (libsbk_probe.so, -DSBK_ENABLE_PROBES: environment switches alter the pipeline and
 * results may be meaningless), 0 in the product library, which never reads the environment. */
int sbk_probes_enabled(void);
const char* sbk_last_error(void);
/* Number of CUDA devices visible (0 when none / no driver). */
int sbk_device_count(void);

/* Workspace (bytes) sbk_knn needs for a block of n_query rows. */
size_t sbk_knn_workspace_bytes(int dtype, int64_t n_ref, int32_t d, int64_t n_query, int32_t k,
                               int algo);

/* Exact Euclidean k nearest neighbours of rows [q_begin, q_end) of X among all
 * n_ref rows of X, self excluded (by index; if self is not among the k+1
 * nearest because of >k exact duplicates the first entry is dropped), ordered
 * by (distance, index).  X is row-major [n_ref, ld] on the device.
 *   out_dist : float[(q_end-q_begin)*k]  for SBK_F32 input, value
 *              f32(sqrt(f64(f32(d2))))  (sklearn _argkmin.pyx.tp:285-295)
 *              double[...]              for SBK_F64 input, value sqrt(d2)
 *   out_idx  : int32[(q_end-q_begin)*k]
 * d2 is the FP64 direct-difference squared distance. */
int sbk_knn(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld,
            int64_t q_begin, int64_t q_end, int32_t k,
            void* out_dist, int32_t* out_idx,
            void* workspace, size_t ws_bytes, int algo,
            sbk_knn_stats* stats, void* cuda_stream);
/* sbk_knn synchronises `cuda_stream` before it returns (it reads back the
 * certificate-failure count to size the fallback launch). */

/* sbk_knn for the row-block sharded multi-GPU build (north_star item 4): besides out_dist / out_idx (this GPU's row
 * block, rows [q_begin, q_end)), every finished row r is ALSO stored at row r of the n_peers FULL [n_ref, k] result
 * buffers peer_dist[t] / peer_idx[t] -- device memory of the other GPUs of the box mapped into this process (CUDA
 * IPC / symmetric memory / peer access).  The stores are plain P2P stores issued by the rerank kernels themselves, so
 * the gather of the row blocks rides on the compute kernel over NVLink instead of following it as a collective; the
 * caller only has to order a barrier after the call before it reads its own full buffer.  peer_dist / peer_idx are
 * HOST arrays of device pointers; n_peers <= 8; 0 peers == sbk_knn. */
int sbk_knn_peers(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld,
                  int64_t q_begin, int64_t q_end, int32_t k,
                  void* out_dist, int32_t* out_idx,
                  int32_t n_peers, void* const* peer_dist, int32_t* const* peer_idx,
                  void* workspace, size_t ws_bytes, int algo,
                  sbk_knn_stats* stats, void* cuda_stream);

/* sbk_knn with the result also delivered to HOST memory (the buffers kneighbors_graph hands to scipy,
 * cluster.py:94-105): the query rows are processed in chunks of `chunk_rows` (0 = the default chunking) and
 * every finished chunk is copied to host_dist / host_idx (pinned memory for an asynchronous copy) on
 * `copy_stream` while the next chunk is computed; rows recomputed by the exact fallback are copied again at
 * the end.  out_dist / out_idx are still the full device outputs.  Both streams are synchronised on return. */
int sbk_knn_stream(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld,
                   int64_t q_begin, int64_t q_end, int32_t k,
                   void* out_dist, int32_t* out_idx,
                   void* host_dist, int32_t* host_idx, int64_t chunk_rows,
                   void* workspace, size_t ws_bytes, int algo,
                   sbk_knn_stats* stats, void* cuda_stream, void* copy_stream);

/* The whole of sklearn.neighbors.kneighbors_graph(X, k, mode=..., p=2) for rows [q_begin, q_end) (cluster.py:94-105,
 * long_read_cluster.py:108-113): the CSR arrays scikit-learn hands back (neighbors/_base.py:1031-1041) are written
 * straight into HOST memory, widened on the device,
 *   host_data    : double[(q_end-q_begin)*k]   distances in ascending order per row (float32-rounded values for
 *                  SBK_F32 input); NULL for mode='connectivity' (the caller fills ones)
 *   host_indices : int64[(q_end-q_begin)*k]
 * (indptr is arange(0, n*k+1, k)).  Pinned host buffers make the per-chunk downloads overlap the next chunk's
 * compute; device outputs, widening staging and the kNN workspace all live in `workspace`.  Both streams are
 * synchronised on return. */
size_t sbk_knn_graph_workspace_bytes(int dtype, int64_t n_ref, int32_t d, int64_t n_query, int32_t k, int algo);
int sbk_knn_graph(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld,
                  int64_t q_begin, int64_t q_end, int32_t k,
                  double* host_data, int64_t* host_indices, int64_t chunk_rows,
                  void* workspace, size_t ws_bytes, int algo,
                  sbk_knn_stats* stats, void* cuda_stream, void* copy_stream);

/* Number of thresholds of the max_node search (cluster.py:118-125). */
#define SBK_MAX_THRESHOLDS 32

/* Stage 1 of the short-read edge list for rows [row_begin, row_begin+n_rows):
 *   keep_bits[r*W + c/32] bit c%32 = 1  iff  emb_idx[r*k+c] is also in kmer_idx row r with kmer_dist != 0 and
 *              emb_dist != 0  (the intersection of cluster.py:109-113; W = ceil(k/32) words per row)
 *   rowmax[r] = max over the kept entries of w = 1 - min(emb_dist, 1), 0 for an empty row   (cluster.py:115-119)
 *   counts[m] += #{r : rowmax[r] > thresholds[m]}  (int64, caller zeroes)   (:122)
 * kmer_dist may be NULL (then every feature-graph entry counts as non-zero).  The N x k float64 weight matrix of
 * the reference is never materialised: stage 2 recomputes w from emb_dist. */
int sbk_edges_stage1(const float* emb_dist, const int32_t* emb_idx,
                     const double* kmer_dist, const int32_t* kmer_idx,
                     int64_t n_rows, int32_t k,
                     const double* thresholds, int32_t n_thresholds,
                     uint32_t* keep_bits, double* rowmax, long long* counts, void* cuda_stream);

size_t sbk_edges_stage2_workspace_bytes(int64_t n_rows, int32_t k, int64_t n_total, int32_t n_sample);

/* Stage 2: w = 1 - min(emb_dist, 1) at the kept entries, drop w <= threshold, multiply by the KL factor (float32
 * op order of cluster.py:129-141 / cal_kl numpy branch; the float32 log of every contig's variance is evaluated
 * once per contig and sample, correctly rounded) unless is_combined, drop <= 1e-6, keep column > row, emit
 * row-major sorted.  depth is float[n_total, 2*n_sample] interleaved [mean_s, var_s] for ALL n_total contigs
 * (columns are global indices).
 *   out_x/out_y : int32[capacity], out_v : double[capacity]
 *   n_edges     : device int64, number of edges produced (if > capacity only
 *                 the first `capacity` are written). */
int sbk_edges_stage2(const uint32_t* keep_bits, const float* emb_dist, const int32_t* emb_idx,
                     int64_t row_begin, int64_t n_rows, int32_t k, double threshold,
                     const float* depth, int64_t n_total, int32_t n_sample, int is_combined,
                     int32_t* out_x, int32_t* out_y, double* out_v, int64_t capacity,
                     long long* n_edges, void* workspace, size_t ws_bytes, void* cuda_stream);

/* eps sweep over a distance-sorted kNN graph, rows [row_begin,row_begin+n_rows):
 *   prefix_len[e*n_rows+r] = #{c : dist[r*k+c] <= eps[e]}   (double compare, eps python floats)
 *   core[e*n_rows+r] = weight[row] + sum_{c<prefix} weight[idx] >= min_samples
 * weight may be NULL (all ones).  dist/idx hold only the n_rows local rows;
 * idx values and weight are global.  eps is a HOST array (n_eps <= 16); every
 * other pointer is device memory. */
int sbk_radius_sweep(const float* dist, const int32_t* idx, int64_t row_begin, int64_t n_rows,
                     int32_t k, const double* eps, int32_t n_eps,
                     const double* weight, double min_samples,
                     int32_t* prefix_len, uint8_t* core, void* cuda_stream);

size_t sbk_dbscan_labels_workspace_bytes(int64_t n, int32_t n_eps);

/* DBSCAN labels for n_eps neighbourhood graphs at once (all n rows resident):
 * labels[e*n+i] = rank of min{u core : u reaches i through core points} among
 * the distinct seeds, or -1 (noise) -- identical to sklearn's dbscan_inner. */
int sbk_dbscan_labels(const int32_t* idx, int64_t n, int32_t k,
                      const int32_t* prefix_len, const uint8_t* core, int32_t n_eps,
                      long long* labels, void* workspace, size_t ws_bytes, void* cuda_stream);

size_t sbk_integrate_workspace_bytes(int64_t n, int32_t n_results);

/* Long-read ensemble integration: the greedy bin extraction of
 * SemiBin/long_read_cluster.py:12-49 (get_best_bin) and :122-150 (the loop and the final labels).
 *   labels       : long long[n_results * n], DBSCAN labels of every result (-1 = noise), device
 *   lengths      : long long[n], contig lengths (len(contig_dict[name])), device
 *   marker_masks : uint64[n * mask_words], bit m set iff the contig carries marker m (each marker at
 *                  most once per contig, markers.py:63), device
 *   n_markers    : the recall denominator (107, long_read_cluster.py:35)
 *   out_labels   : long long[n], index of the extraction that took the contig, or -1, device
 *   n_extracted  : host int, number of bins extracted
 * Every bin of every result is scored as in :26-46 (length >= minfasta, at least one marker, contamination
 * step, F1 in the reference's floating-point operation order, ties to the smaller length and then to the
 * bin met LAST), the winner's contigs are retired from all results and the loop repeats until the
 * remaining length is below minfasta, one contig is left (it becomes a bin) or no bin qualifies.
 * Synchronises `cuda_stream`. */
int sbk_integrate(const long long* labels, int32_t n_results, int64_t n, const long long* lengths,
                  const unsigned long long* marker_masks, int32_t mask_words, int32_t n_markers,
                  int64_t minfasta, long long* out_labels, int32_t* n_extracted,
                  void* workspace, size_t ws_bytes, void* cuda_stream);

size_t sbk_kmeans_batch_workspace_bytes(int64_t n_total, int32_t d, int32_t n_problems, int64_t k_total);

/* Re-clustering step, SemiBin/cluster.py:229-246: for every bin with more than one seed the reference calls
 *   KMeans(n_clusters=num_bin, init=seeds_embedding, n_init=1, random_state=seed).fit(re_bin_features,
 *          sample_weight=length_weight)                         (scikit-learn 1.9.0, algorithm 'lloyd')
 * one bin after the other; this entry point solves ALL bins of a sample in one launch (one CTA per bin runs the whole
 * Lloyd loop: E-step, weighted M-step, empty-cluster relocation, the two convergence tests, the final E-step).
 *   X              : float[n_total, ld >= d], the rows of every problem CONTIGUOUS (problem p = rows
 *                    [point_offsets[p], point_offsets[p+1])), device
 *   weight         : float[n_total], sample weights (contig lengths), device
 *   point_offsets  : int64[n_problems + 1], HOST
 *   cluster_offsets: int32[n_problems + 1], HOST: problem p has cluster_offsets[p+1] - cluster_offsets[p] clusters
 *   centers_init   : float[cluster_offsets[n_problems], d], the seed rows (init=...), device
 *   max_iter, tol  : KMeans defaults 300 and 1e-4 (tol is relative to the mean column variance, _kmeans.py:_tolerance)
 *   labels         : int32[n_total] out, cluster index within the problem (KMeans.labels_), device
 *   n_iter         : int32[n_problems] out (KMeans.n_iter_), device
 *   centers_out    : double[cluster_offsets[n_problems], d] out (KMeans.cluster_centers_) or NULL, device
 * With an array init and n_init = 1 nothing is random.  Sums are FP64 in index order on the float32 data centred by
 * its float32 column mean (as fit() does); labels equal scikit-learn's except where a point's two nearest final
 * centres tie to ~1e-4 relative.  n_samples < n_clusters in a problem -> SBK_ERR_INVALID with sklearn's message. */
int sbk_kmeans_batch(const float* X, int64_t ld, int32_t d, const float* weight, int64_t n_total,
                     const int64_t* point_offsets, const int32_t* cluster_offsets, int32_t n_problems,
                     const float* centers_init, int32_t max_iter, double tol,
                     int32_t* labels, int32_t* n_iter, double* centers_out,
                     void* workspace, size_t ws_bytes, void* cuda_stream);

/* Symmetric build of the whole kNN graph across the GPUs of one box, one process per GPU (north_star item 4 with the
 * symmetric filter): GPU `rank` owns the staged positions of tile block `rank`, multiplies its own block, the following
 * (world-1)/2 blocks and -- for an even world -- half of the opposite block, so that every pair of tiles is multiplied
 * once across all GPUs.  Candidates it finds for rows of another GPU are stored straight into that GPU's mailbox (peer
 * memory over NVLink) and merged there; every GPU reranks its own rows and stores each finished row into its own and
 * every peer's full [n_ref, k] result (out_dist / out_idx, peer_dist / peer_idx: as sbk_knn_peers).
 *   xchg[r]  : base of rank r's exchange buffer (sbk_knn_sym_shard_xchg_bytes each, same size on every rank, mapped into
 *              this process: e.g. torch symmetric memory), r = 0 .. world-1, own buffer included
 *   barrier  : barrier(barrier_arg, failed) is called three times (after the sample thresholds, after every rank's own
 *              block, after the last tile pair); it must return once every rank has called it, with a non-zero value when
 *              any rank passed failed != 0 (the library synchronises its stream first).  The vote makes all ranks leave
 *              a failed build together.
 * Returns SBK_ERR_INVALID when the symmetric filter does not take the build, SBK_ERR_INTERNAL when a record pool or a
 * mailbox was too small on this or (voted) another rank; a mailbox overflow is only known to the receiving rank at the
 * end, so the caller votes once more after the call.  In all these cases the caller shards by row blocks instead. */
size_t sbk_knn_sym_shard_xchg_bytes(int dtype, int64_t n_ref, int32_t d, int32_t k, int32_t world);
int sbk_knn_sym_shard(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld, int32_t k,
                      int32_t rank, int32_t world, void* const* xchg, int (*barrier)(void*, int), void* barrier_arg,
                      void* out_dist, int32_t* out_idx, int32_t n_peers, void* const* peer_dist, int32_t* const* peer_idx,
                      void* workspace, size_t ws_bytes, sbk_knn_stats* stats, void* cuda_stream);

/* Debug / test hook: fixes the capacity (records) of every mailbox of the sharded symmetric build, 0 = planned size.
 * Set it identically on every rank BEFORE sbk_knn_sym_shard_xchg_bytes; a small value forces the overflow vote. */
void sbk_debug_shard_mail_cap(int64_t records);

/* Debug / test hook, host only: the tile products sbk_knn_sym_shard launches on `rank` of `world` for a build of n_tiles
 * 256-row tiles.  Writes five int64 per step into out (at most max_steps steps): kind, q_lo, q_len, j_lo, j_len --
 * kind 0: the block of query tiles [q_lo, q_lo + q_len) against itself, circulant (tile I x tile q_lo + (I - q_lo + s) mod
 * q_len for s = 0 .. q_len / 2); kind 1: query tiles [q_lo, q_lo + q_len) x reference tiles [j_lo, j_lo + j_len).
 * Returns the number of steps (< 0: error).  Over all ranks every ordered pair of tiles is tested exactly once. */
int sbk_debug_sym_shard_schedule(int64_t n_tiles, int32_t world, int32_t rank, int64_t* out, int32_t max_steps);

/* Debug / test hook: raw tcgen05 tile product.  Runs the same FP16 operand staging and
 * MMA pipeline as the tensor filter and returns the accumulator values
 *   L[i,j] <= S^2 d2_ij - nrm_i + norms4[i].w      (i < n_q, j < n_ref <= 65536)
 * in out[n_q, ceil(n_ref/256)*256] (float), the per-row norms of the error model
 * norms4[n_ref*4] = (||z||, ||delta||, ||a||, ||l||^2), nrm[n_ref+1] = exact ||z||^2 followed by the
 * scale S, and info[2] = {kpad, n_split_cols}, so tests can verify the lower/upper bound claims. */
int sbk_debug_tc_scores(const void* X, int dtype, int64_t n_ref, int32_t d, int64_t ld,
                        int64_t n_q, float* out, float* norms4, double* nrm, int32_t* info,
                        void* workspace, size_t ws_bytes, void* cuda_stream);
size_t sbk_debug_tc_workspace_bytes(int64_t n_ref, int32_t d);

#ifdef __cplusplus
}
#endif
#endif /* SBK_H_ */
