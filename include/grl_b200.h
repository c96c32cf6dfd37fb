/* grl_b200 — C ABI of the B200-native GRL hot path (sm_100a).
 *
 * The reference (flysnowtiger/GRL) is pure Python/PyTorch and has NO FFI: its seams for
 * this path are Python call signatures.  Each entry point below names the reference
 * interface it replaces (file:line under /root/reference); INTEGRATION.md shows the
 * ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - All tensor pointers are DEVICE pointers (fp32, contiguous, row-major / NCHW) unless
 *     the parameter name ends in `_host`.  The caller owns every buffer, including the
 *     workspace; the library performs no hidden device allocation on the hot path.
 *   - Work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*);
 *     calls return without synchronising unless documented.
 *   - Return 0 on success, a negative GRL_E* code otherwise; grl_last_error() gives text.
 *     Nothing throws across the boundary.  There is no CPU fallback.
 */
#ifndef GRL_B200_H
#define GRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define GRL_OK 0
#define GRL_EINVAL (-1)   /* bad shape / alignment / null pointer / unsupported mode */
#define GRL_ECUDA (-2)    /* CUDA runtime or driver error (message has the cudaError) */
#define GRL_EARCH (-3)    /* device is not sm_100 */
#define GRL_ENOMEM (-4)   /* caller-provided workspace too small */
#define GRL_ENCCL (-5)    /* NCCL missing (dlopen) or a collective failed (message has the ncclResult) */

typedef struct grl_handle grl_handle;

/* ---- lifecycle -------------------------------------------------------------------- */
int grl_create(int device, grl_handle** out);
void grl_destroy(grl_handle* h);
const char* grl_last_error(const grl_handle* h);   /* h may be NULL: last create() error */
const char* grl_version(void);
int grl_num_sms(const grl_handle* h);

/* ---- dense contraction primitive (test/debug surface of the tcgen05 kernel) -------- */
/* D[z][m][n] = sum_k A[z][m][k] * B[z][n][k], split-bf16 (3 MMA) with fp32 accumulation.
 * Operands are fp32; they are split into bf16 hi/lo planes inside `workspace`.
 * a_mn_major/b_mn_major = 0: operand stored [rows][K] (K contiguous); 1: stored [K][rows].
 * Epilogue: v = alpha*acc; v *= row_scale[m]; v += col_bias[n]; relu; optional C += v.
 * Optional outputs: col_sum/col_sq [4*ceil(M/128)][N] per-(tile,warp) partial column sums,
 * planes_hi/lo bf16 [M][N].  bn = 0 (auto) | 128 | 256 selects the N tile.                */
typedef struct grl_gemm_desc {
    int M, N, K, batch;
    int a_mn_major, b_mn_major;
    long long lda, ldb, ldc;            /* leading dimensions in elements */
    long long a_bstride, b_bstride, c_bstride;
    float alpha;
    const float* row_scale;             /* [M] or NULL */
    const float* col_bias;              /* [N] or NULL */
    int relu, accumulate, bn;
    float* col_sum;                     /* or NULL */
    float* col_sq;                      /* or NULL */
    uint16_t* planes_hi;                /* bf16 bits, or NULL */
    uint16_t* planes_lo;
} grl_gemm_desc;
size_t grl_gemm_workspace_bytes(const grl_gemm_desc* d);
int grl_gemm_bf16x3(grl_handle* h, const grl_gemm_desc* d, const float* A, const float* B, float* C,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- matching / evaluation ---------------------------------------------------------- */
/* cosin_dist(qf, gf) = -qf @ gf.T           reid/evaluator/attevaluator.py:44-46
 * pairwise_distance_tensor(qf, gf)          reid/evaluator/attevaluator.py:33-41
 * q [nq][dim], g [ng][dim] fp32 -> dist [nq][ng] fp32.  metric: 0 = negative dot, 1 = L2. */
#define GRL_METRIC_NEG_DOT 0
#define GRL_METRIC_L2 1
size_t grl_distance_workspace_bytes(int nq, int ng, int dim);
int grl_distance(grl_handle* h, int metric, const float* q, const float* g, int nq, int ng, int dim,
                 float* dist, void* workspace, size_t workspace_bytes, void* stream);

/* eva_functions.evaluate(distmat, q_pids, g_pids, q_camids, g_camids, max_rank)
 *                                            reid/evaluator/eva_functions.py:134-184
 * Sort-free: per query, rank of each positive = #kept gallery rows strictly closer (ties: lower
 * index first == stable argsort).  Junk (same pid & same cam) removed, queries without a match
 * skipped.  Outputs (device): cmc_hits int32[max_rank] (number of valid queries whose first hit
 * is at rank <= r), ap double[nq] (-1 for skipped queries), first_hit int32[nq] (-1 skipped).
 * The Python wrapper turns these into (np.float32[max_rank], float) exactly like :179-182.   */
int grl_cmc_map(grl_handle* h, const float* dist, long long ld_dist, const int64_t* q_pid, const int64_t* g_pid,
                const int64_t* q_cam, const int64_t* g_cam, int nq, int ng, int max_rank,
                int32_t* cmc_hits, double* ap, int32_t* first_hit, void* stream);

/* The same evaluation over a gallery whose rows are sharded over the ranks of the handle's communicator (grl_comm_*; one
 * rank without): `dist` [nq][ng_local] holds this rank's columns, g_pid / g_cam its rows, idx_base the global index of its
 * first row (shards are contiguous and in rank order); q_pid / q_cam are replicated.  Two exchanges: an all-gather of each
 * query's positives (distance, global index; at most max_pos per query and shard) and an all-reduce(sum) of the per-shard
 * (#kept, #positive) counts of rows ordered before each positive; every rank then holds the same ap / first_hit / cmc_hits,
 * bit-identical to grl_cmc_map on the concatenated matrix.  max_seen (device int32) returns the largest per-(query, shard)
 * positive count over all ranks: the result is valid iff max_seen <= max_pos (the caller re-runs with a larger max_pos).  */
size_t grl_cmc_map_sharded_workspace_bytes(const grl_handle* h, int nq, int max_pos);
int grl_cmc_map_sharded(grl_handle* h, const float* dist, long long ld_dist, const int64_t* q_pid, const int64_t* g_pid,
                        const int64_t* q_cam, const int64_t* g_cam, int nq, int ng_local, int64_t idx_base, int max_rank, int max_pos,
                        int32_t* cmc_hits, double* ap, int32_t* first_hit, int32_t* max_seen, void* workspace, size_t workspace_bytes,
                        void* stream);

/* np.argsort(distmat, axis=1)                reid/evaluator/eva_functions.py:139
 * Stable (distance, index) row sort, ng <= 16384.  order int32 [nq][ng].                    */
int grl_argsort_rows(grl_handle* h, const float* dist, long long ld_dist, int nq, int ng, int32_t* order, void* stream);

/* Gallery-sharded retrieval (BASELINE.json config 5; no reference counterpart: the reference
 * materialises the full matrix and argsorts it on the CPU, eva_functions.py:139).
 * grl_topk_rows: k smallest (distance, index) per row of a distance tile, merged into the running
 * per-query lists top_d/top_i [nq][k] (initialise top_d to +inf, top_i to -1 via grl_topk_init).
 * idx_base is added to column indices (global gallery index of column 0).
 * grl_topk_merge: merges `nshards` lists laid out [nshards][nq][k] into out_d/out_i [nq][k],
 * ties by lower global index, so the result does not depend on the shard count.              */
int grl_topk_init(grl_handle* h, float* top_d, int64_t* top_i, int nq, int k, void* stream);
int grl_topk_rows(grl_handle* h, const float* dist, long long ld_dist, int nq, int ncols, int k, int64_t idx_base,
                  float* top_d, int64_t* top_i, void* stream);
int grl_topk_merge(grl_handle* h, const float* all_d, const int64_t* all_i, int nshards, int nq, int k,
                   float* out_d, int64_t* out_i, void* stream);
/* One gallery shard, end to end: the k nearest gallery rows of every query without materialising the nq x ng matrix (40 GB
 * at 10k x 1M).  Two-stage exact search:
 *   1. grl_coarse_topk   fp16 operands (per-row power-of-two scaling), ONE tcgen05 MMA per k-step, candidate filter fused
 *                        into the GEMM epilogue, streaming top-K' (K' = grl_topk_kprime(k)) by COARSE distance;
 *   2. grl_rescore       fixed-order fp32 inner products of the K' candidates (the distance the result reports);
 *   3. grl_topk_finalize sort by (exact distance, global index), emit the top k, and PROVE per query that no row outside
 *                        the K' candidates can belong to them (coarse K'-th value minus the worst-case rounding error of
 *                        the coarse pass still exceeds the exact k-th value); queries where the proof fails are flagged;
 *   4. grl_exact_topk    brute force in the same fixed-order arithmetic for the flagged queries (rare).
 * The result is the exact stable top-k of the fixed-order fp32 distances, independent of chunking and of the number of
 * gallery shards.  g [ng][dim] is this rank's shard, idx_base the global index of its first row; top_d/top_i [nq][k] are
 * overwritten.  grl_dist_topk runs 1-4 for one shard and synchronises the stream once (it reads the flag count); it is the
 * single-rank form of grl_sharded_topk (below), which runs the whole sharded protocol incl. its NCCL collectives.
 * The stage entry points remain for hosts that run their own collectives between them: all-gather + grl_topk_merge of the
 * coarse lists, each rank re-scores the candidates it owns (grl_rescore writes 0 for rows of other shards, so the per-rank
 * results combine by a sum), finalize, and the flagged queries go through grl_exact_topk per shard + grl_topk_merge.
 * coarse_d holds -q.g (metric 0) or the SQUARED L2 distance (metric 1); gmax2 [1] = max |g|^2 over the shard (combine
 * shards with a max); dirty int32 [nq] = 1 where a per-chunk candidate buffer overflowed (only the first chunk's distance
 * tile is ever stored, so such a row cannot be rescanned: it is flagged and brute-forced; combine shards with a max; may be
 * NULL for grl_topk_finalize); flags int32 [nq], nflag int32 [1].                                                                */
/* A static gallery shard can be converted ONCE (fp16 rows with their per-row scales, squared norms, largest squared norm) and
 * searched many times: grl_gallery_prepare fills a caller-owned buffer of grl_gallery_prepared_bytes(ng, dim) bytes (256-byte
 * aligned); the *_prepared entry points take it in place of the per-search conversion.  The fp32 rows are still needed for
 * the re-score (grl_rescore / grl_dist_topk_prepared / grl_exact_topk).                                                    */
size_t grl_gallery_prepared_bytes(int ng, int dim);
int grl_gallery_prepare(grl_handle* h, const float* g, int ng, int dim, void* prepared, size_t prepared_bytes, void* stream);
int grl_coarse_topk_prepared(grl_handle* h, int metric, const float* q, const void* prepared, int nq, int ng, int dim, int kprime,
                             int64_t idx_base, float* coarse_d, int64_t* coarse_i, float* gmax2, int32_t* dirty, void* workspace,
                             size_t workspace_bytes, void* stream);
int grl_dist_topk_prepared(grl_handle* h, int metric, const float* q, const float* g, const void* prepared, int nq, int ng, int dim,
                           int k, int64_t idx_base, float* top_d, int64_t* top_i, void* workspace, size_t workspace_bytes,
                           void* stream);
int grl_topk_kprime(int k);
size_t grl_coarse_topk_workspace_bytes(int nq, int ng, int dim);
int grl_coarse_topk(grl_handle* h, int metric, const float* q, const float* g, int nq, int ng, int dim, int kprime,
                    int64_t idx_base, float* coarse_d, int64_t* coarse_i, float* gmax2, int32_t* dirty, void* workspace,
                    size_t workspace_bytes, void* stream);
int grl_rescore(grl_handle* h, int metric, const float* q, const float* g, int nq, int ng, int dim, int64_t idx_base,
                const int64_t* cand_i, int kprime, float* exact_d, void* stream);
int grl_topk_finalize(grl_handle* h, int metric, const float* q, int nq, int dim, const float* coarse_d, const int64_t* cand_i,
                      const float* exact_d, int kprime, const float* gmax2, const int32_t* dirty, int k, float* top_d,
                      int64_t* top_i, int32_t* flags, int32_t* nflag, void* stream);
size_t grl_exact_topk_workspace_bytes(int nq, int ng, int dim);
int grl_exact_topk(grl_handle* h, int metric, const float* q, const float* g, int nq, int ng, int dim, int k, int64_t idx_base,
                   float* top_d, int64_t* top_i, void* workspace, size_t workspace_bytes, void* stream);
size_t grl_dist_topk_workspace_bytes(int nq, int ng, int dim);
int grl_dist_topk(grl_handle* h, int metric, const float* q, const float* g, int nq, int ng, int dim, int k,
                  int64_t idx_base, float* top_d, int64_t* top_i, void* workspace, size_t workspace_bytes, void* stream);

/* ---- gallery-sharded retrieval over NCCL (BASELINE.json configs[4]: top-k merged over 1/2/4/8 GPUs) --------------------
 * No reference counterpart (the reference evaluates on one device: attevaluator.py:125-163, and drives its GPUs from one
 * process with nn.DataParallel: mars_train.py:80).  One process per GPU; NCCL is resolved at run time (dlopen of
 * libnccl.so.2 -- inside a PyTorch process that is the instance torch already loaded), so the library has no link-time
 * dependency on it.
 * Communicator: rank 0 calls grl_comm_unique_id, the host distributes the GRL_COMM_ID_BYTES bytes by any means, every rank
 * calls grl_comm_init (ncclCommInitRank; the handle owns the communicator) -- or hands in its own ncclComm_t with
 * grl_comm_attach (borrowed; NULL detaches).  grl_comm_info reports (1, 0) without a communicator.                        */
#define GRL_COMM_ID_BYTES 128
int grl_comm_unique_id(grl_handle* h, void* id_host, size_t id_bytes);
int grl_comm_init(grl_handle* h, const void* id_host, size_t id_bytes, int world, int rank);
int grl_comm_attach(grl_handle* h, void* nccl_comm);
int grl_comm_destroy(grl_handle* h);
int grl_comm_info(const grl_handle* h, int* world, int* rank);

/* One search over a gallery whose rows are split contiguously over the W ranks of the handle's communicator (W = 1 without
 * one): the whole protocol behind ONE call, every collective on the caller's stream.
 *   q         query rows (device, fp32).  q_is_slice == 0: all nq queries (replicated by the caller); q_is_slice != 0: this
 *             rank's slice, rows [rank*qs, min(nq, (rank+1)*qs)) with qs = ceil(nq / W) (may be empty, then q may be NULL) --
 *             the ranks all-gather the slices over NVLink instead of each uploading all nq rows over PCIe.  The flag selects
 *             which collectives run: every rank must pass the same value
 *   g_local   this rank's gallery rows [ng_local][dim] (fp32; needed for the exact re-score), idx_base = global index of row 0
 *   prepared  grl_gallery_prepare'd index of g_local, or NULL (the shard is then converted chunk by chunk in every search)
 *   top_d / top_i [nq][k]: the exact stable top-k of the fixed-order fp32 distances, identical on every rank and for every W
 *   max_flagged < 0: synchronous -- the stream is synchronised once to read how many queries failed the completeness proof,
 *             and exactly those go through the brute-force leg.  >= 0: asynchronous (graph-capturable): the brute-force leg is
 *             launched for max_flagged row slots gated on the device-side count; queries beyond stay as the coarse search
 *             left them and the caller checks stats[0] <= max_flagged afterwards
 *   stats     device int32[8] or NULL: [0] queries flagged (proof failed / buffer overflow), [1] rows with an overflow mark,
 *             [2] candidates re-scored by this rank, [3] candidates this rank skipped (provably outside the top k) -- both of
 *             the first pass --, [4] flagged queries proven by the second chance, [5] queries that went to brute force
 * Stages: S0 query all-gather, S1 fp16 conversion, S2 coarse pass (tcgen05 GEMM + candidate filter + list merges) over the
 * local shard, S3 all-reduce(max) of {overflow marks, max |g|^2} + all-to-all of the K' lists by query slice + merge +
 * all-gather of the merged slices, S4 owned re-score, S5 reduce-scatter(sum), S6 finalisation + proof of the own slice,
 * S7 all-gather of the result keys + flag compaction + unpack, S8 the queries without a proof: (synchronous mode) a second
 * chance -- S1..S7 again over those rows with the longest lists, K' = 1024, which reaches past a cluster of near-duplicates at
 * a fraction of the brute-force cost -- then brute force (per shard + all-gather + merge) for what is still unproven.
 * grl_search_profile(h, 1) records an event at every stage boundary of the following calls; grl_search_stage_ms returns the
 * nine stage durations of the last one (synchronises on its last event).                                                 */
size_t grl_sharded_topk_workspace_bytes(const grl_handle* h, int nq, int ng_local, int dim, int k, int prepared);
int grl_sharded_topk(grl_handle* h, int metric, const float* q, int q_is_slice, const float* g_local, const void* prepared, int nq,
                     int ng_local, int dim, int k, int64_t idx_base, int max_flagged, float* top_d, int64_t* top_i, int32_t* stats,
                     void* workspace, size_t workspace_bytes, void* stream);
#define GRL_SEARCH_STAGES 9
int grl_search_profile(grl_handle* h, int on);

/* The exchange step of grl_sharded_topk on its own (for callers that move the shards' lists themselves, and for the tests):
 * `lists` [nlists][rows][kp] holds, per shard and query row, kp packed keys in ascending order (orderable(distance) << 32 | global
 * gallery index, all-ones = empty slot -- the list format of the search); `out` [rows][kp] receives the kp smallest keys of every row,
 * ascending.  A merge tree over the sorted lists (not a sort of their union).  kp: a power of two <= 1024; nlists (rounded up to a
 * power of two) * kp <= 16384.  Replaces nothing in the reference (its evaluator holds one gallery, attevaluator.py:125-161).       */
int grl_merge_key_lists(grl_handle* h, const unsigned long long* lists, int nlists, int rows, int kp, unsigned long long* out, void* stream);
int grl_search_stage_ms(grl_handle* h, double* ms, int n);

/* k-reciprocal re-ranking: re_ranking(q_g_dist, q_q_dist, g_g_dist, k1, k2, lambda_value)
 *                                            reid/evaluator/rerank.py:37-104 (called at attevaluator.py:151-155)
 * q_g [nq][ng], q_q [nq][nq], g_g [ng][ng] fp32 distance matrices -> final_dist [nq][ng] fp32.  Same float32 operation
 * order as the reference (column-normalised squared distances, k-reciprocal sets with 2/3-overlap expansion, np.sum's
 * pairwise order, rank-ordered query-expansion mean, ascending-column Jaccard sums); ties in the neighbour ranking go
 * to the lower index.  Limits: nq + ng <= 65536, 1 <= k1 <= 31, 1 <= k2 <= 32.  Workspace: ~8 bytes x (nq+ng)^2.      */
size_t grl_rerank_workspace_bytes(int nq, int ng, int k1, int k2);
int grl_rerank(grl_handle* h, const float* q_g, const float* q_q, const float* g_g, int nq, int ng, int k1, int k2,
               double lambda_value, float* final_dist, void* workspace, size_t workspace_bytes, void* stream);

/* ---- GCE + TRL head ------------------------------------------------------------------ */
/* Parameter block: device pointers to the reference's own state_dict tensors (fp32), in the
 * reference's names (reid/models/basebranch.py:38-50, reid/models/grl_model.py:88-128).
 * BN buffers (running_mean / running_var) are updated in place in train mode exactly like
 * torch.nn.BatchNorm (momentum 0.1, unbiased variance); num_batches_tracked is advanced by the
 * Python wrapper (+1 per GCE BN, +T per TRL BN per forward).                                 */
typedef struct grl_bn_params {
    const float* weight;
    const float* bias;
    float* running_mean;
    float* running_var;
} grl_bn_params;

typedef struct grl_head_params {
    /* GCE (Backbone.glo_fc / Backbone.corr_atte) */
    const float* glo_fc_w;      /* [1024][2048] */
    const float* glo_fc_b;      /* [1024] */
    grl_bn_params glo_bn;       /* BatchNorm1d(1024) */
    const float* atte0_w;       /* [1024][3072]  conv1x1, no bias */
    grl_bn_params atte_bn1;     /* BatchNorm2d(1024) */
    const float* atte2_w;       /* [256][1024] */
    grl_bn_params atte_bn3;     /* BatchNorm2d(256) */
    const float* atte5_w;       /* [256] */
    grl_bn_params atte_bn6;     /* BatchNorm2d(1) */
    /* TRL, index 0 = forward direction, 1 = backward direction */
    const float* f1_w[2];       /* [2048][2048] */
    const float* f1_b[2];
    const float* f2_w[2];
    const float* f2_b[2];
    const float* se1_w[2];      /* channel_atte_*_corr.0.weight [128][2048] */
    const float* se2_w[2];      /* channel_atte_*_corr.2.weight [2048][128] */
    const float* memo_conv1_w[2];   /* [512][2048] */
    grl_bn_params memo_bn1[2];
    const float* memo_conv2_w[2];   /* [512][512] */
    grl_bn_params memo_bn2[2];
    const float* memo_conv3_w[2];   /* [2048][512] */
    grl_bn_params memo_bn3[2];
} grl_head_params;

/* Gradients w.r.t. the same tensors (fp32, same shapes); every pointer must be non-NULL.
 * They are OVERWRITTEN (not accumulated) by grl_head_backward.                              */
typedef struct grl_head_grads {
    float* glo_fc_w; float* glo_fc_b; float* glo_bn_w; float* glo_bn_b;
    float* atte0_w; float* atte_bn1_w; float* atte_bn1_b;
    float* atte2_w; float* atte_bn3_w; float* atte_bn3_b;
    float* atte5_w; float* atte_bn6_w; float* atte_bn6_b;
    float* f1_w[2]; float* f1_b[2]; float* f2_w[2]; float* f2_b[2];
    float* se1_w[2]; float* se2_w[2];
    float* memo_conv1_w[2]; float* memo_bn1_w[2]; float* memo_bn1_b[2];
    float* memo_conv2_w[2]; float* memo_bn2_w[2]; float* memo_bn2_b[2];
    float* memo_conv3_w[2]; float* memo_bn3_w[2]; float* memo_bn3_b[2];
} grl_head_grads;

/* Workspace for B clips of T frames.  save_for_backward != 0 keeps every per-step activation
 * (needed by grl_head_backward); 0 reuses one step slot (inference).                          */
size_t grl_head_workspace_bytes(int B, int T, int save_for_backward);

/* BasicBlock.forward(x1, x2)                 reid/models/grl_model.py:67-85 (the memory-update module of TRLBlock, :88-128)
 *   relu(bn3(conv3(relu(bn2(conv2(relu(bn1(conv1(x1 + x2)))))))) + (x1 + x2));  x1, x2, out [n][2048][16][8] (NCHW fp32).
 * conv1_w [512][2048], conv2_w [512][512], conv3_w [2048][512] (1x1, no bias); train = 1: batch statistics over the n*128
 * pixels + running-buffer update (like one step of the recurrence), 0: running statistics.  Forward only: inside the head
 * the block is differentiated by grl_head_backward / grl_trl_backward.                                                    */
size_t grl_basic_block_workspace_bytes(int n);
int grl_basic_block_forward(grl_handle* h, const float* conv1_w, const grl_bn_params* bn1, const float* conv2_w, const grl_bn_params* bn2,
                            const float* conv3_w, const grl_bn_params* bn3, const float* x1, const float* x2, int n, int train,
                            float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Fused GCE + TRL forward.
 *   Backbone.forward after self.base     reid/models/basebranch.py:56-68
 *   TRLBlock.forward                     reid/models/grl_model.py:131-180
 * x         [B*T][2048][16][8]  layer4 maps (NCHW)
 * train     1: BatchNorm batch statistics (+ running-buffer update), 0: running statistics
 * f_uncorr  [B][2048]     f_corr [B][T][2048]     corr_map [B*T][1][16][8]
 * x_uncorr / x_corr [B*T][2048][16][8]: optional (NULL to skip; the fused path never needs them). */
int grl_head_forward(grl_handle* h, const grl_head_params* p, const float* x, int B, int T, int train,
                     float* f_uncorr, float* f_corr, float* corr_map, float* x_uncorr, float* x_corr,
                     void* workspace, size_t workspace_bytes, int save_for_backward, void* stream);

/* Backward of grl_head_forward (autograd replacement for the same reference lines).  Needs the
 * workspace of a train=1, save_for_backward=1 forward with the same (B,T) and unchanged params.
 * d_f_uncorr [B][2048], d_f_corr [B][T][2048]; optional upstream grads of the stand-alone outputs
 * d_x_uncorr / d_x_corr [B*T][2048][16][8] and d_corr_map [B*T][128] (NULL = zero).
 * dx [B*T][2048][16][8]; grads: every parameter gradient.                                      */
int grl_head_backward(grl_handle* h, const grl_head_params* p, const float* x, int B, int T,
                      const float* d_f_uncorr, const float* d_f_corr,
                      const float* d_x_uncorr, const float* d_x_corr, const float* d_corr_map,
                      float* dx, const grl_head_grads* grads,
                      void* workspace, size_t workspace_bytes, void* stream);

/* The two halves as stand-alone operators (the reference's modules can be called separately):
 *   grl_gce_forward  = Backbone.forward after self.base (basebranch.py:56-68): x -> x_uncorr, x_corr, corr_map (all written)
 *   grl_trl_forward  = TRLBlock.forward (grl_model.py:131-180): x_uncorr, x_corr [B*T][2048][16][8] -> f_uncorr, f_corr
 * and their backwards (same workspace contract as the fused pair; grads: only the GCE resp. TRL members are read
 * and written).  d_x_uncorr / d_x_corr / d_corr_map of grl_gce_backward may be NULL (= zero).                      */
int grl_gce_forward(grl_handle* h, const grl_head_params* p, const float* x, int B, int T, int train,
                    float* x_uncorr, float* x_corr, float* corr_map,
                    void* workspace, size_t workspace_bytes, int save_for_backward, void* stream);
int grl_gce_backward(grl_handle* h, const grl_head_params* p, int B, int T,
                     const float* d_x_uncorr, const float* d_x_corr, const float* d_corr_map,
                     float* dx, const grl_head_grads* grads, void* workspace, size_t workspace_bytes, void* stream);
int grl_trl_forward(grl_handle* h, const grl_head_params* p, const float* x_uncorr, const float* x_corr, int B, int T,
                    int train, float* f_uncorr, float* f_corr,
                    void* workspace, size_t workspace_bytes, int save_for_backward, void* stream);
int grl_trl_backward(grl_handle* h, const grl_head_params* p, int B, int T, const float* d_f_uncorr, const float* d_f_corr,
                     float* d_x_uncorr, float* d_x_corr, const grl_head_grads* grads,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- eval feature tail (what sits between the head and the distance GEMM at evaluation time) --------------- */
/* Head outputs -> per-clip 6144-d descriptor cat(x_uncorr, self_attention(x_corr), mean_t x_corr), eval-mode BN:
 *   corr_bn / uncorr_bn + F.normalize            reid/models/grl_model.py:222-226
 *   Siamese.self_attention                       reid/models/Siamese.py:79-106   (featQ / featK: Linear(2048, 512) + BN1d)
 *   the concat                                   reid/evaluator/attevaluator.py:79-80
 * f_uncorr [n][2048], f_corr [n][T][2048] -> out [n][ld_out >= 6144].  apply_tail_bn = 0: the inputs are already the
 * model's normalised outputs (x_uncorr, x_corr) and corr_bn / uncorr_bn may be NULL; out[:, 2048:4096] is then exactly
 * Siamese.self_attention(x_corr).                                                                                    */
typedef struct grl_tail_params {
    grl_bn_params corr_bn, uncorr_bn;          /* BatchNorm1d(2048) of ResNet50_GRL_Model */
    const float* featQ_w; const float* featQ_b; grl_bn_params featQ_bn;     /* [512][2048], [512], BatchNorm1d(512) */
    const float* featK_w; const float* featK_b; grl_bn_params featK_bn;
} grl_tail_params;
size_t grl_eval_descriptor_workspace_bytes(int n, int T);
int grl_eval_descriptor(grl_handle* h, const grl_tail_params* p, const float* f_uncorr, const float* f_corr, int n, int T,
                        int apply_tail_bn, float* out, long long ld_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- loss neighbours of the head ------------------------------------------------------------------- */
/* OIMLoss (reid/loss/oim.py:8-58): logits = scalar * x @ lut^T [B][C] (the module's second return value), probs = softmax,
 * row_loss [B], loss [1] = mean cross-entropy (F.cross_entropy).  Backward = the legacy OIM.backward contract (:19-27):
 * dx = d_loss * scalar / B * (probs - onehot) @ lut with the table BEFORE its update (dx may be NULL), then
 * lut[y] = momentum * lut[y] + (1 - momentum) * x_i; lut[y] /= |lut[y]| for every sample in batch order (in place).       */
int grl_oim_forward(grl_handle* h, const float* x, const int64_t* targets, const float* lut, int B, int C, int D, float scalar,
                    float* logits, float* probs, float* row_loss, float* loss, void* stream);
int grl_oim_backward(grl_handle* h, const float* x, const int64_t* targets, float* lut, const float* probs, int B, int C, int D,
                     float scalar, float momentum, const float* d_loss, float* dx, void* stream);
/* TripletLoss(margin, batch_hard=True).forward(feat, id) (reid/loss/triplet.py:16-90; mode 'id', dis_func 'eu', n_dis 0):
 * loss [B] = log(1 + exp(z)) (soft != 0) or clamp(z + margin, 0), z = hardest positive - hardest negative distance.
 * The forward also returns what the backward needs: z, the two arg-extrema (pos_idx = -1 when the row has no positive)
 * and their distances.                                                                                                   */
int grl_triplet_forward(grl_handle* h, const float* feat, const int64_t* ids, int B, int D, int soft, float margin, float* loss,
                        float* z, int32_t* pos_idx, int32_t* neg_idx, float* pos_d, float* neg_d, void* stream);
int grl_triplet_backward(grl_handle* h, const float* feat, int B, int D, int soft, float margin, const float* z,
                         const int32_t* pos_idx, const int32_t* neg_idx, const float* pos_d, const float* neg_d,
                         const float* d_loss, float* dfeat, void* stream);

/* Siamese.forward in training (reid/models/Siamese.py:108-142): temporal self-attention pooling (:79-106) of the probe half
 * and of the gallery half (train-mode BatchNorm statistics per half, running buffers updated probe first), squared pair
 * differences -> classifierBN -> classifierlinear.  x [2n][T][2048] holds the n probe clips first, then the n gallery clips
 * (the reference pairs samples (2i, 2i+1): the Python mirror regroups); siamese_out [2n][2048] in the same order;
 * cls_encode [n][n][2].  The backward needs the forward's workspace and siamese_out; d_siamese_out may be NULL; grads are
 * overwritten.  featV / featV_bn exist in the module but take no part in forward (Siamese.py:99).  T <= 32.               */
typedef struct grl_siamese_params {
    const float* featQ_w; const float* featQ_b; grl_bn_params featQ_bn;     /* Linear(2048, 512), BatchNorm1d(512) */
    const float* featK_w; const float* featK_b; grl_bn_params featK_bn;
    grl_bn_params cls_bn; const float* cls_w; const float* cls_b;          /* BatchNorm1d(2048), Linear(2048, 2) */
} grl_siamese_params;
typedef struct grl_siamese_grads {
    float* featQ_w; float* featQ_b; float* featQ_bn_w; float* featQ_bn_b;
    float* featK_w; float* featK_b; float* featK_bn_w; float* featK_bn_b;
    float* cls_bn_w; float* cls_bn_b; float* cls_w; float* cls_b;
} grl_siamese_grads;
size_t grl_siamese_workspace_bytes(int n2, int T);
int grl_siamese_forward(grl_handle* h, const grl_siamese_params* p, const float* x, int n2, int T, int train, float* cls_encode,
                        float* siamese_out, void* workspace, size_t workspace_bytes, void* stream);
int grl_siamese_backward(grl_handle* h, const grl_siamese_params* p, const float* x, const float* siamese_out, int n2, int T,
                         const float* d_cls_encode, const float* d_siamese_out, float* dx, const grl_siamese_grads* g,
                         void* workspace, size_t workspace_bytes, void* stream);
/* PairLoss.forward(score [n][n], tar_probe [n], tar_gallery [n]) -> (loss, prec)   reid/loss/pairloss.py:19-48: mean binary
 * cross-entropy of the pair scores (log clamped at -100 like torch.nn.BCELoss) against label[p][g] =
 * (tar_probe[g] == tar_gallery[p]) -- the reference's expand/eq order -- and the fraction of correct decisions.          */
int grl_pair_loss_forward(grl_handle* h, const float* score, const int64_t* tar_probe, const int64_t* tar_gallery, int n,
                          float* loss, float* prec, void* stream);
int grl_pair_loss_backward(grl_handle* h, const float* score, const int64_t* tar_probe, const int64_t* tar_gallery, int n,
                           const float* d_loss, float* d_score, void* stream);

/* Debug/test: byte offset and size of a named intermediate inside the head workspace
 * (e.g. "xp_hi", "y1", "m", "f2", "memo_h1").  Returns GRL_EINVAL for unknown names.         */
int grl_head_ws_lookup(int B, int T, int save_for_backward, const char* name, size_t* offset, size_t* bytes);

/* Number of kernel launches issued by this handle since creation (bench.py's gpu_launches).   */
long long grl_launch_count(const grl_handle* h);

/* The head entry points fork work that is off the recurrence's critical path (f1 / f2 convolutions, weight
 * gradients) onto an internal low-priority stream and join it back into the caller's stream before returning, so
 * HBM-bound glue overlaps tensor-core work.  `mask`: bit 0 = forward, bit 1 = backward; default 3; 0 runs everything
 * on the caller's stream (debugging, per-kernel profiling).  Bit 3 (value 8) is a debugging switch of the retrieval search: it
 * replaces the CTA-pair (cta_group::2) coarse GEMM by the single-CTA 256 x 256-tile kernel.  Bits 4-7 are A/B switches of the head
 * step (tools/ab_head_pairs.py, tools/ab_bn_fuse.py): 16 = single-CTA GEMM kernels only, 32 = fp16 GEMMs on the single-CTA kernel,
 * 64 = CTA pairs for every 256-wide tile, 128 = separate BatchNorm-backward reduce passes for the memory block's bn1 / bn2.                  */
int grl_set_overlap(grl_handle* h, int mask);

/* Profiling aid for bench.py's roofline: while enabled, every tcgen05 GEMM launch is bracketed by CUDA
 * events on the caller's stream.  grl_profile_read waits for them and returns, since the last read, the
 * summed device time of those launches (ms), their algorithmic FLOPs (2*M*N*K*batch; the split-bf16
 * kernel issues 3x that in MMAs) and their count.  Off by default; not used on the product path.        */
int grl_profile_enable(grl_handle* h, int on);
int grl_profile_read(grl_handle* h, double* gemm_ms, double* gemm_flops, long long* gemm_launches);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* GRL_B200_H */
