// knn.cuh -- internal launch interfaces of the kNN pipeline stages.
#pragma once
#include "common.cuh"
#include <vector>

namespace sbk {

// Mirror destinations of the rerank output: the full [n_ref, k] result buffers of the other GPUs of the box (peer
// memory mapped into this process).  The rerank kernels store every finished row to all of them with plain P2P
// stores over NVLink, so the row-block gather of the multi-GPU build rides on the compute kernel instead of
// following it as a collective.
struct PeerOut { int n; void* dist[8]; int32_t* idx[8]; };

// ---- exact (FP64 CUDA-core) filter + rerank: knn_exact.cu ----
int exact_filter_cap(int k1);
size_t exact_filter_smem(int d, int k1);
int launch_exact_filter(const void* X, int dtype, long long n_ref, int d, long long ld,
                        const int32_t* rows, long long q_begin, long long n_q, int n_split, int k1,
                        uint32_t* cand_idx, long long cand_stride, int32_t* cand_count,
                        cudaStream_t stream);
int launch_rerank(const void* X, int dtype, long long n_ref, int d, long long ld,
                  const int32_t* rows, long long q_begin, long long n_slots,
                  const uint32_t* cand_idx, const uint2* cand_pair, long long cand_stride, int n_seg,
                  const int32_t* cand_count, const double* bound, const double* nrm, const float4* norms4,
                  const double* scale, const int32_t* orig, int k,
                  void* out_dist, int32_t* out_idx, long long out_row0,
                  int32_t* fail_rows, int* n_fail, unsigned long long* totals,
                  cudaStream_t stream, const int32_t* slot_list = nullptr, const int* n_list = nullptr, int list_cap = 0,
                  const PeerOut* peers = nullptr);
// warp-per-row rerank straight from the filter's candidate lists (tensor path); rows it cannot take (too many
// candidates for its registers / survivors for its sort) are appended to retry_slots / n_retry for launch_rerank's
// retry mode, uncertified rows to fail_rows
int rerank_warp_cap(int k);
bool rerank_warp_supported(int dtype, int d, int k);
size_t row_order_workspace_bytes(long long n_rows);
// locality order of the chunk's query rows for the warp rerank: perm_out[i] = slot visited i-th (lives in `ws`);
// with a row list (`rows`, n_rows global row ids) the order is over those rows and perm_out holds row ids - q_begin
int launch_row_order(const void* X, int dtype, int d, long long ld, long long q_begin, long long n_rows,
                     const double* centre, void* ws, size_t ws_bytes, int32_t** perm_out, cudaStream_t stream,
                     const int32_t* rows = nullptr);
// row_slot (optional): candidate lists, thresholds and bounds are indexed by row_slot[row] instead of row - q_begin
// (the symmetric filter keeps them in staged-position order)
int launch_rerank_warp(const void* X, int dtype, int d, long long ld, long long q_begin, long long n_rows, const int32_t* perm,
                       const int32_t* row_slot,
                       const uint2* cand, const int32_t* cand_count, int cand_cap, const float* thr, const double* bound,
                       const double* nrm, const float4* norms4, const double* scale, const int32_t* orig, int k,
                       void* out_dist, int32_t* out_idx, long long out_row0, int32_t* fail_rows, int* n_fail,
                       int32_t* retry_slots, int* n_retry, int retry_cap, unsigned long long* totals, cudaStream_t stream,
                       const PeerOut* peers = nullptr);

// ---- tensor-core (tcgen05) filter: knn_tc.cu ----
struct TcPlan {
  int kpad;            // padded K (multiple of 16) = d + 3 norm columns + 2 per split column, rounded up
  int kpad_alloc;      // widest kpad the staging buffers are sized for
  int max_split;       // most columns tc_prepare may split
  int n_split;         // columns actually split (set by tc_prepare)
  int ksteps;          // kpad / 16
  long long n_tiles;   // reference tiles of TC_BN rows
  int sample_size;     // rows in the threshold sample (multiple of TC_BN)
  int order_stat;      // r: threshold = r-th smallest group minimum
  int n_groups;        // G
  int cand_cap;        // candidate slots per query row
  int cand_cap_planned; // what the expected candidate count asks for (the symmetric build may be given more room)
  int pool_limit;      // test hook: at most this many pool chunks (0 = sized for the build); forces the exhaustion fallback
  size_t stage_bytes;  // bytes of one staged reference tile
};
static constexpr int TC_BM = 128;   // query rows per CTA
static constexpr int TC_BN = 256;   // reference rows per MMA tile

struct TcScalars;

// ---- per-pair error model of the tensor-core score (shared by knn_tc.cu and the rerank) ----
// Allowance for the FP32 accumulation inside the tensor core: per MMA instruction (K = 16) each of
// the 16 products and the running sum may lose up to 2^-24 of the largest magnitude: 17 * 2^-24 < 2^-20
// relative to the sum of absolute values, per instruction.
__host__ __device__ static inline double tc_c_acc(int kpad) { return (double)(kpad / 16) * 9.5367431640625e-07; }
// norms4[j] = ( ||z_j||, ||delta_j||, ||a_j||, ||l_j||^2 ), every entry rounded up.
//   z = S (x - c) exact, a = FP16-representable part (hi, + lo on split columns), delta = z - a,
//   l = low parts of the split columns.
// exact score  e_ij = ||z_j||^2 - 2 <z_i, z_j>;  the MMA yields L_ij with  L_ij <= e_ij + ||l_i||^2  where it
// subtracted at least  G_ij + J_j:
//   G_ij = 2 ||delta_i|| ||z_j|| + ||a_i|| (2 ||delta_j|| + 2 c_acc ||a_j||)      (bilinear -> two K columns)
//   J_j  = c_acc ||z_j||^2 + ||l_j||^2                                             (folded into the norm pieces)
__host__ __device__ static inline double tc_pair_G(const float4& ni, const float4& nj, double c_acc) {
  return 2.0 * (double)ni.y * (double)nj.x + (double)ni.z * (2.0 * (double)nj.y + 2.0 * c_acc * (double)nj.z);
}
__host__ __device__ static inline double tc_pair_J(const float4& nj, double c_acc) {
  return c_acc * (double)nj.x * (double)nj.x + (double)nj.w;
}
// e_ij <= L_ij + tc_pair_upper_slack: what the MMA subtracted (<= (1+2^-9)^2 G + J + representation slack)
// plus the true error bound G + J + ||l_i||^2.
__host__ __device__ static inline double tc_pair_upper_slack(const float4& ni, const float4& nj, double c_acc) {
  const double G = tc_pair_G(ni, nj, c_acc), J = tc_pair_J(nj, c_acc);
  return 2.02 * G + 2.02 * J + (double)ni.w + 2.4e-7 * ((double)nj.x + (double)nj.y + (double)ni.z + (double)ni.y) + 4e-7;
}
// Sharded symmetric build (one process per GPU): GPU g owns the positions of tile block g.  A column hit whose row lives
// on another GPU goes into this GPU's outbox for that GPU (local memory); after a barrier the owner reads the outboxes
// addressed to it straight from the senders' memory (coalesced peer loads over NVLink; as 16-byte peer STORES from the
// transpose kernel the same records cost 65 ms at 1M on two GPUs) and merges them; `cap` = 0 means single-GPU.
struct TcMail {
  uint4* out[8];            // out[c]: this rank's outbox for rank c (records (j, i, s bits, -))
  int* out_cnt;             // [8] local counters of records sent to each rank
  int cap;                  // records per (source, destination) region
  long long blk_pos;        // positions per tile block (owner of position p = p / blk_pos)
  long long own_lo, own_hi; // positions this rank owns
};

struct TcState {
  __half* a_tiles; __half* b_tiles;
  double* centre; double* colsum; double* scale; double* nrm;
  float4* norms4; TcScalars* sc; int32_t* orig; float* thr;   // orig[p] = row staged at reference position p
  // symmetric self-join filter (queries == references): everything per staged POSITION
  int sym;                       // A tiles staged in position order too; slots are positions
  int32_t* rpos;                 // rpos[row] = position
  float2* cu;                    // (c_p, u_p): ||z||^2 - sigma rounded down / up (moves a score between the two rows' frames)
  float4* tq;                    // (c_p, u_p, t_p = thr_p + u_p, -) per position, refreshed with the thresholds
  float* t8;                     // per group of 8 positions: max of the column thresholds t_p = thr_p + u_p
  uint4* pool; int* pool_count; int pool_chunks;   // group records of one filter launch, in per-warp chunk streams
  unsigned long long* maxabs;
  int nstage; int cg; long long n_ref; int d;   // cg: CTAs per MMA (tcgen05 cta_group)
  long long perm_mult, perm_off, perm_minv;     // reference position p holds row (p * perm_mult + perm_off) mod n_ref
};

TcPlan tc_make_plan(long long n_ref, int d, int k);
long long tc_chunk_rows(const TcPlan& p, long long n_ref);
size_t tc_workspace_bytes(const TcPlan& p, long long n_ref, long long n_query, bool sym = false);
// whether the symmetric self-join filter can take a build of all n_ref rows (candidate lists of every row resident at once)
bool tc_sym_supported(const TcPlan& p, long long n_ref);
static constexpr long long TC_SYM_MIN_TILES = 1600;   // AUTO / TENSOR pick the symmetric filter from this many 256-row tiles on
static constexpr int TC_SYM_POOL_EXHAUSTED = 1000;   // tc_filter_sym: not an error, the caller falls back to the one-sided filter
// Stage operands, centre/scale, sample; `chunk` = max query rows per tc_filter_chunk call.
int tc_prepare(const void* X, int dtype, long long n_ref, int d, long long ld, TcPlan& p,
               long long chunk, void* ws, size_t ws_bytes, TcState* st, cudaStream_t stream, bool sym = false);
// Symmetric self-join filter: every unordered tile pair is multiplied once and tested for both rows (see knn_tc.cu).
// Thresholds must have been produced by tc_sample_chunk(st, p, 0, n_ref, ...) on a state prepared with sym = true;
// cand / cand_count / bound are indexed by staged position.
int tc_filter_sym(TcState* st, const TcPlan& p, int k, uint2* cand, int32_t* cand_count, double* bound,
                  int* n_launches, int* n_aux_launches, cudaStream_t stream);
// The same filter block row by block row (tc_sym_begin, tc_sym_block for r = 0 .. n_blocks-1, tc_sym_end): after block
// row r the rows at positions [r, r+1) * tc_sym_block_tiles * 256 are finished (complete candidate lists).
long long tc_sym_block_tiles(const TcPlan& p, int n_blocks);
int tc_sym_begin(TcState* st, const TcPlan& p, uint2* cand, int32_t* cand_count, double* bound, int* n_aux_launches,
                 cudaStream_t stream);
int tc_sym_block(TcState* st, const TcPlan& p, int k, uint2* cand, int32_t* cand_count, double* bound, int n_blocks, int r,
                 int* n_launches, int* n_aux_launches, cudaStream_t stream);
int tc_sym_end(TcState* st, cudaStream_t stream);
// ---- sharded symmetric build (rank of world GPUs; see TcMail) ----
long long tc_shard_block_tiles(const TcPlan& p, int world);
// own thresholds (positions of the rank's block) into every rank's exchange array
int tc_shard_bcast_thr(TcState* st, const TcPlan& p, int rank, int world, float* const* xthr, cudaStream_t stream);
// tile pairs of this rank: own block circulantly, the following (world-1)/2 blocks in full, and for an even world half
// of the opposite block; row hits and local column hits into the lists, remote column hits into the mailboxes
struct TcShardStep { int kind; long long q_lo, q_len, j_lo, j_len; };
std::vector<TcShardStep> tc_shard_schedule(long long n_tiles, int world, int rank);
void tc_trace(const char* what, cudaStream_t stream);
void tc_trace_dump(int rank);
int tc_sym_shard_filter(TcState* st, const TcPlan& p, int k, int rank, int world, uint2* cand, int32_t* cand_count, double* bound,
                        const TcMail& mail, int* const* peer_count_in, int* n_launches, int* n_aux_launches, cudaStream_t stream,
                        int (*mid_exchange)(void*, int) = nullptr, void* mid_arg = nullptr);
// t, T8 of every position from the current thresholds (after thresholds of other ranks arrived)
int tc_refresh_colthr(TcState* st, const TcPlan& p, cudaStream_t stream);
// records other ranks hold for this one (boxes[src]: the outbox of rank src for this rank, in src's memory; count_in[src])
// into the lists
int tc_mail_merge(TcState* st, const TcPlan& p, uint4* const* boxes, const int* count_in, int world, int cap,
                  uint2* cand, int32_t* cand_count, int* overflow_flag, cudaStream_t stream);
// Sample pass for query rows [row0, row0+n_rows): per-row thresholds (st->thr)
// and certificate bounds bound[slot].
int tc_sample_chunk(TcState* st, const TcPlan& p, long long row0, long long n_rows, double* bound,
                    cudaStream_t stream);
// Filter pass over all references: fills 4 segments of cand_cap/4 (column, score) pairs per slot,
// cand_count[slot*4 + segment].
// Candidate columns are reference POSITIONS (st->orig maps them to rows); the thresholds (and `bound`) are
// tightened between launches from the candidates seen so far.
int tc_filter_chunk(TcState* st, const TcPlan& p, long long row0, long long n_rows, int k,
                    uint2* cand, int32_t* cand_count, double* bound, int* n_launches, int* n_refines, cudaStream_t stream);
int tc_dump_scores(TcState* st, const TcPlan& p, long long n_q, float* out, long long out_ld, cudaStream_t stream);
int tc_debug_norms(TcState* st, const TcPlan& p, long long n, float* norms4_out, double* nrm_out, int32_t* info,
                   cudaStream_t stream);

}  // namespace sbk
