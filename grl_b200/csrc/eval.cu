// grl_b200 — re-ID matching kernels (sm_100a): distance via the split-bf16 tcgen05 GEMM,
// sort-free CMC / mAP, stable row argsort, streaming top-k and shard merge.
//
// Reference lines replaced:
//   cosin_dist / pairwise_distance_tensor   reid/evaluator/attevaluator.py:33-46
//   evaluate (argsort + CMC + AP)           reid/evaluator/eva_functions.py:134-184
#include <math_constants.h>

#include <algorithm>

#include "api.h"
#include "comm.h"
#include "eval_common.cuh"

namespace grl {

// ------------------------------------------------------------------ row squared norms (L2 metric)
__global__ void row_sqnorm_kernel(const float* __restrict__ x, int rows, int dim, float* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= rows) return;
    const float* r = x + (long long)warp * dim;
    float acc = 0.f;
    for (int i = lane_id() * 4; i < dim; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(r + i);
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    acc = warp_sum(acc);
    if (lane_id() == 0) out[warp] = acc;
}

// ------------------------------------------------------------------ CMC / mAP without sorting
// One block per query.  For every positive gallery row g+ of the query:
//   ca = #{kept rows ordered before g+}, cp = #{positive rows ordered before g+}
// (order = (distance, index), i.e. a stable argsort).  Then
//   AP = mean over positives of (cp+1)/(ca+1)   and   first hit = min ca.
// This equals eva_functions.py:150-176 (cumsum / (i+1) * match summed, / num_rel).
constexpr int CMC_THREADS = 256;
constexpr int CMC_CHUNK = 512;   // positives staged in shared memory per round

__global__ void __launch_bounds__(CMC_THREADS) cmc_map_kernel(const float* __restrict__ dist, long long ld,
                                                              const int64_t* __restrict__ q_pid, const int64_t* __restrict__ g_pid,
                                                              const int64_t* __restrict__ q_cam, const int64_t* __restrict__ g_cam,
                                                              int ng, double* __restrict__ ap, int32_t* __restrict__ first_hit) {
    __shared__ float pd[CMC_CHUNK];
    __shared__ int pi[CMC_CHUNK];
    __shared__ int warp_cnt[CMC_THREADS / 32];
    __shared__ double warp_ap[CMC_THREADS / 32];
    __shared__ int warp_first[CMC_THREADS / 32];
    const int q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = CMC_THREADS / 32;
    const int64_t qp = q_pid[q], qc = q_cam[q];
    const float* drow = dist + (long long)q * ld;
    // contiguous gallery segment per warp, so ordinals follow index order
    const int seg = ((ng + NW - 1) / NW + 31) / 32 * 32;
    const int g_begin = warp * seg, g_end = min(ng, g_begin + seg);

    int cnt = 0;
    for (int g0 = g_begin; g0 < g_end; g0 += 32) {
        const int g = g0 + lane;
        const bool pos = g < g_end && g_pid[g] == qp && g_cam[g] != qc;
        cnt += __popc(__ballot_sync(0xffffffffu, pos));
    }
    if (lane == 0) warp_cnt[warp] = cnt;
    __syncthreads();
    int warp_base = 0, npos = 0;
    for (int w = 0; w < NW; ++w) {
        if (w < warp) warp_base += warp_cnt[w];
        npos += warp_cnt[w];
    }
    if (npos == 0) {                                  // eva_functions.py:159-161: query skipped
        if (threadIdx.x == 0) { ap[q] = -1.0; first_hit[q] = -1; }
        return;
    }
    double ap_acc = 0.0;
    int first = 0x7fffffff;
    for (int base = 0; base < npos; base += CMC_CHUNK) {
        __syncthreads();
        int ord = warp_base;
        for (int g0 = g_begin; g0 < g_end; g0 += 32) {
            const int g = g0 + lane;
            const bool pos = g < g_end && g_pid[g] == qp && g_cam[g] != qc;
            const unsigned m = __ballot_sync(0xffffffffu, pos);
            const int my = ord + __popc(m & ((1u << lane) - 1));
            if (pos && my >= base && my < base + CMC_CHUNK) { pd[my - base] = drow[g]; pi[my - base] = g; }
            ord += __popc(m);
        }
        __syncthreads();
        const int nchunk = min(CMC_CHUNK, npos - base);
        for (int j = warp; j < nchunk; j += NW) {
            const float dj = pd[j];
            const int ij = pi[j];
            int ca = 0, cp = 0;
            for (int g = lane; g < ng; g += 32) {
                const float d = drow[g];
                const bool before = (d < dj) || (d == dj && g < ij);
                if (before) {
                    const bool same = g_pid[g] == qp;
                    const bool junk = same && g_cam[g] == qc;
                    ca += junk ? 0 : 1;
                    cp += (same && !junk) ? 1 : 0;
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                ca += __shfl_xor_sync(0xffffffffu, ca, off);
                cp += __shfl_xor_sync(0xffffffffu, cp, off);
            }
            ap_acc += (double)(cp + 1) / (double)(ca + 1);
            first = min(first, ca);
        }
    }
    if (lane == 0) { warp_ap[warp] = ap_acc; warp_first[warp] = first; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        int f = 0x7fffffff;
        for (int w = 0; w < NW; ++w) { s += warp_ap[w]; f = min(f, warp_first[w]); }
        ap[q] = s / (double)npos;
        first_hit[q] = f;
    }
}

__global__ void cmc_hits_kernel(const int32_t* __restrict__ first_hit, int nq, int max_rank, int32_t* __restrict__ hits) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= max_rank) return;
    int c = 0;
    for (int q = 0; q < nq; ++q) {
        const int f = first_hit[q];
        c += (f >= 0 && f <= r) ? 1 : 0;
    }
    hits[r] = c;
}

// ------------------------------------------------------------------ CMC / mAP over a gallery sharded across ranks
// Same arithmetic as cmc_map_kernel, split in three steps around two collectives:
//   collect  per query, this shard's positives as (distance bits, GLOBAL gallery index) words in ascending index order,
//            slot [max_pos] of the row = their count  ->  all-gather over the ranks
//   count    for every positive of every shard: (#kept, #positive) LOCAL rows ordered before it  ->  all-reduce(sum)
//   reduce   global ranks -> AP (summed in cmc_map_kernel's order: eight interleaved partial sums over the positives in
//            ascending global index, then added in order) and the first hit
// Shards are contiguous row ranges in rank order, so (shard, slot) order == ascending global index order.
__global__ void __launch_bounds__(CMC_THREADS) cmc_collect_pos_kernel(const float* __restrict__ dist, long long ld, const int64_t* __restrict__ q_pid,
                                                                      const int64_t* __restrict__ g_pid, const int64_t* __restrict__ q_cam,
                                                                      const int64_t* __restrict__ g_cam, int ng, long long idx_base, int max_pos,
                                                                      unsigned long long* __restrict__ pos, int32_t* __restrict__ max_seen) {
    __shared__ int warp_cnt[CMC_THREADS / 32];
    const int q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = CMC_THREADS / 32;
    const int64_t qp = q_pid[q], qc = q_cam[q];
    const float* drow = dist + (long long)q * ld;
    unsigned long long* prow = pos + (long long)q * (max_pos + 1);
    const int seg = ((ng + NW - 1) / NW + 31) / 32 * 32;
    const int g_begin = warp * seg, g_end = min(ng, g_begin + seg);
    int cnt = 0;
    for (int g0 = g_begin; g0 < g_end; g0 += 32) {
        const int g = g0 + lane;
        const bool p = g < g_end && g_pid[g] == qp && g_cam[g] != qc;
        cnt += __popc(__ballot_sync(0xffffffffu, p));
    }
    if (lane == 0) warp_cnt[warp] = cnt;
    __syncthreads();
    int ord = 0, npos = 0;
    for (int w = 0; w < NW; ++w) {
        if (w < warp) ord += warp_cnt[w];
        npos += warp_cnt[w];
    }
    for (int g0 = g_begin; g0 < g_end; g0 += 32) {
        const int g = g0 + lane;
        const bool p = g < g_end && g_pid[g] == qp && g_cam[g] != qc;
        const unsigned m = __ballot_sync(0xffffffffu, p);
        const int my = ord + __popc(m & ((1u << lane) - 1));
        if (p && my < max_pos) prow[my] = ((unsigned long long)__float_as_uint(drow[g]) << 32) | (unsigned long long)(uint32_t)(idx_base + g);
        ord += __popc(m);
    }
    if (threadIdx.x == 0) { prow[max_pos] = (unsigned long long)npos; atomicMax(max_seen, npos); }
}

__global__ void __launch_bounds__(CMC_THREADS) cmc_count_kernel(const float* __restrict__ dist, long long ld, const int64_t* __restrict__ q_pid,
                                                                const int64_t* __restrict__ g_pid, const int64_t* __restrict__ q_cam,
                                                                const int64_t* __restrict__ g_cam, int ng, long long idx_base, int nq, int world,
                                                                int max_pos, const unsigned long long* __restrict__ pos_all, int32_t* __restrict__ cnt) {
    const int q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = CMC_THREADS / 32;
    const int64_t qp = q_pid[q], qc = q_cam[q];
    const float* drow = dist + (long long)q * ld;
    const int nslots = world * max_pos;
    for (int sl = warp; sl < nslots; sl += NW) {
        const int s = sl / max_pos, j = sl - s * max_pos;
        const unsigned long long* prow = pos_all + ((long long)s * nq + q) * (max_pos + 1);
        int ca = 0, cp = 0;
        if (j < min((int)prow[max_pos], max_pos)) {
            const float dj = __uint_as_float((uint32_t)(prow[j] >> 32));
            const long long ij = (long long)(uint32_t)(prow[j] & 0xFFFFFFFFull);
            for (int g = lane; g < ng; g += 32) {
                const float d = drow[g];
                const bool before = (d < dj) || (d == dj && idx_base + g < ij);
                if (before) {
                    const bool same = g_pid[g] == qp;
                    const bool junk = same && g_cam[g] == qc;
                    ca += junk ? 0 : 1;
                    cp += (same && !junk) ? 1 : 0;
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                ca += __shfl_xor_sync(0xffffffffu, ca, off);
                cp += __shfl_xor_sync(0xffffffffu, cp, off);
            }
        }
        if (lane == 0) { cnt[((long long)q * nslots + sl) * 2] = ca; cnt[((long long)q * nslots + sl) * 2 + 1] = cp; }
    }
}

__global__ void cmc_reduce_kernel(int nq, int world, int max_pos, const unsigned long long* __restrict__ pos_all, const int32_t* __restrict__ cnt,
                                  double* __restrict__ ap, int32_t* __restrict__ first_hit) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    constexpr int NW = CMC_THREADS / 32;
    double acc[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) acc[w] = 0.0;
    int first = 0x7fffffff, ord = 0;
    const int nslots = world * max_pos;
    for (int s = 0; s < world; ++s) {
        const int n = min((int)pos_all[((long long)s * nq + q) * (max_pos + 1) + max_pos], max_pos);
        for (int j = 0; j < n; ++j, ++ord) {
            const int32_t* c = cnt + ((long long)q * nslots + s * max_pos + j) * 2;
            const double term = (double)(c[1] + 1) / (double)(c[0] + 1);
#pragma unroll
            for (int w = 0; w < NW; ++w) if ((ord & (NW - 1)) == w) acc[w] += term;   // cmc_map_kernel: warp w sums ordinals == w mod 8
            first = min(first, c[0]);
        }
    }
    if (ord == 0) { ap[q] = -1.0; first_hit[q] = -1; return; }
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) sum += acc[w];
    ap[q] = sum / (double)ord;
    first_hit[q] = first;
}

// ------------------------------------------------------------------ stable argsort of rows (ng <= 16384)
__global__ void __launch_bounds__(1024) argsort_rows_kernel(const float* __restrict__ dist, long long ld, int ng, int npad,
                                                            int32_t* __restrict__ order) {
    extern __shared__ uint64_t keys[];
    const float* drow = dist + (long long)blockIdx.x * ld;
    for (int i = threadIdx.x; i < npad; i += blockDim.x)
        keys[i] = i < ng ? make_key(drow[i], (uint32_t)i) : ~0ull;
    block_bitonic_sort(keys, npad);
    int32_t* orow = order + (long long)blockIdx.x * ng;
    for (int i = threadIdx.x; i < ng; i += blockDim.x) orow[i] = (int32_t)(keys[i] & 0xFFFFFFFFu);
}

// ------------------------------------------------------------------ streaming top-k per row
// Keeps the k smallest (distance, global index) keys of everything seen so far.  Columns are
// filtered against the current k-th best key; survivors are staged in shared memory and folded
// into the running list by a block bitonic sort whenever the staging buffer may overflow.

__global__ void __launch_bounds__(TOPK_THREADS) topk_rows_kernel(const float* __restrict__ dist, long long ld, int ncols, int k,
                                                                 int64_t idx_base, float* __restrict__ top_d,
                                                                 int64_t* __restrict__ top_i) {
    __shared__ uint64_t keys[TOPK_BUF];
    __shared__ int count;
    __shared__ uint64_t thresh;
    const int row = blockIdx.x;
    const float* drow = dist + (long long)row * ld;
    float* td = top_d + (long long)row * k;
    int64_t* ti = top_i + (long long)row * k;
    // running list occupies keys[0..k); global indices must fit 32 bits inside a key
    for (int i = threadIdx.x; i < TOPK_BUF; i += blockDim.x) {
        uint64_t key = ~0ull;
        if (i < k && ti[i] >= 0) key = make_key(td[i], (uint32_t)ti[i]);
        keys[i] = key;
    }
    if (threadIdx.x == 0) count = k;
    __syncthreads();
    if (threadIdx.x == 0) thresh = keys[k - 1];      // lists are kept sorted, so this is the k-th best
    __syncthreads();
    // flush when a full wave might not fit -- and, for short lists, as soon as one wave is staged: the first flush is what
    // establishes a threshold, after which almost nothing passes the filter
    const int stage_cap = k <= 256 ? TOPK_WAVE : TOPK_BUF - TOPK_WAVE;
    int c0 = 0;
    while (c0 < ncols) {
        const uint64_t th = thresh;
        // a short list without a threshold yet: stage ONE column per thread and sort right away (a 512-key sort instead of a
        // 2048-key one); with the threshold of 256 columns in place only ~k * ln(ncols / 256) more keys are ever staged
        const bool boot = th == ~0ull && k <= 128;
        const int nu = boot ? 1 : 4;
        for (int u = 0; u < nu; ++u) {
            const int c = c0 + u * TOPK_THREADS + threadIdx.x;
            if (c < ncols) {
                const uint64_t key = make_key(drow[c], (uint32_t)(idx_base + c));
                if (key < th) keys[atomicAdd(&count, 1)] = key;
            }
        }
        __syncthreads();
        c0 += nu * TOPK_THREADS;
        const bool last = c0 >= ncols;
        // every thread takes ONE snapshot of the counter and a barrier follows before anyone can bump it again for the next
        // wave: the flush decision is block-uniform (a divergent decision would split the barriers inside the sort)
        const int n = count;
        __syncthreads();
        if (n > stage_cap || (boot && n > k) || (last && n > k)) {
            int npad = 2;
            while (npad < n) npad <<= 1;                 // sort no more than what is staged
            for (int i = n + threadIdx.x; i < npad; i += blockDim.x) keys[i] = ~0ull;
            block_bitonic_sort(keys, npad);
            if (threadIdx.x == 0) { count = k; thresh = keys[k - 1]; }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const uint64_t key = keys[i];
        if (key == ~0ull) { td[i] = CUDART_INF_F; ti[i] = -1; }
        else { td[i] = from_orderable((uint32_t)(key >> 32)); ti[i] = (int64_t)(key & 0xFFFFFFFFu); }
    }
}

__global__ void topk_init_kernel(float* top_d, int64_t* top_i, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) { top_d[i] = CUDART_INF_F; top_i[i] = -1; }
}

// merge nshards sorted lists [nshards][nq][k] -> [nq][k]; (distance, global index) order
__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ all_d, const int64_t* __restrict__ all_i,
                                                         int nshards, int nq, int k, int npad, float* __restrict__ out_d,
                                                         int64_t* __restrict__ out_i) {
    extern __shared__ uint64_t keys[];
    const int row = blockIdx.x;
    const int n = nshards * k;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        uint64_t key = ~0ull;
        if (i < n) {
            const int s = i / k, j = i - s * k;
            const long long src = ((long long)s * nq + row) * k + j;
            if (all_i[src] >= 0) key = make_key(all_d[src], (uint32_t)all_i[src]);
        }
        keys[i] = key;
    }
    block_bitonic_sort(keys, npad);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const uint64_t key = keys[i];
        if (key == ~0ull) { out_d[(long long)row * k + i] = CUDART_INF_F; out_i[(long long)row * k + i] = -1; }
        else {
            out_d[(long long)row * k + i] = from_orderable((uint32_t)(key >> 32));
            out_i[(long long)row * k + i] = (int64_t)(key & 0xFFFFFFFFu);
        }
    }
}


}  // namespace grl

using namespace grl;

// ------------------------------------------------------------------ C ABI
static int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

static void dist_ws_layout(int nq, int ng, int dim, size_t* q_pl, size_t* g_pl, size_t* norms) {
    *q_pl = align_up((size_t)nq * dim * 2, 1024);
    *g_pl = align_up((size_t)ng * dim * 2, 1024);
    *norms = align_up((size_t)(nq + ng) * 4, 1024);
}

extern "C" size_t grl_distance_workspace_bytes(int nq, int ng, int dim) {
    size_t a, b, c;
    dist_ws_layout(nq, ng, dim, &a, &b, &c);
    return 2 * a + 2 * b + c;
}

extern "C" int grl_distance(grl_handle* h, int metric, const float* q, const float* g, int nq, int ng, int dim, float* dist,
                            void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !q || !g || !dist || !workspace) return set_error(h, GRL_EINVAL, "grl_distance: NULL argument");
    if (nq <= 0 || ng <= 0 || dim <= 0 || (dim & 7)) return set_error(h, GRL_EINVAL, "grl_distance: need nq,ng > 0 and dim %% 8 == 0 (dim=%d)", dim);
    if (metric != GRL_METRIC_NEG_DOT && metric != GRL_METRIC_L2) return set_error(h, GRL_EINVAL, "grl_distance: unknown metric %d", metric);
    if (workspace_bytes < grl_distance_workspace_bytes(nq, ng, dim)) return set_error(h, GRL_ENOMEM, "grl_distance: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    size_t qb, gb, nb;
    dist_ws_layout(nq, ng, dim, &qb, &gb, &nb);
    uint8_t* w = (uint8_t*)workspace;
    __nv_bfloat16* q_hi = (__nv_bfloat16*)w; w += qb;
    __nv_bfloat16* q_lo = (__nv_bfloat16*)w; w += qb;
    __nv_bfloat16* g_hi = (__nv_bfloat16*)w; w += gb;
    __nv_bfloat16* g_lo = (__nv_bfloat16*)w; w += gb;
    float* qn = (float*)w;
    float* gn = qn + nq;
    GRL_TRY(split_planes(h, st, q, dim, q_hi, q_lo, dim, nq, dim));
    GRL_TRY(split_planes(h, st, g, dim, g_hi, g_lo, dim, ng, dim));
    GemmEpi e = epi_default();
    e.C = dist; e.ldc = ng;
    if (metric == GRL_METRIC_L2) {
        row_sqnorm_kernel<<<(nq * 32 + 255) / 256, 256, 0, st>>>(q, nq, dim, qn);
        GRL_LAUNCH_CHECK(h);
        row_sqnorm_kernel<<<(ng * 32 + 255) / 256, 256, 0, st>>>(g, ng, dim, gn);
        GRL_LAUNCH_CHECK(h);
        e.mode = 1; e.row_norm = qn; e.col_norm = gn;
    } else {
        e.alpha = -1.f;
    }
    Operand oa{q_hi, q_lo, dim, 0, 0}, ob{g_hi, g_lo, dim, 0, 0};
    return gemm_launch(h, st, nq, ng, dim, 1, oa, ob, e, 0);
}

extern "C" int grl_cmc_map(grl_handle* h, const float* dist, long long ld_dist, const int64_t* q_pid, const int64_t* g_pid,
                           const int64_t* q_cam, const int64_t* g_cam, int nq, int ng, int max_rank, int32_t* cmc_hits,
                           double* ap, int32_t* first_hit, void* stream) {
    if (!h || !dist || !q_pid || !g_pid || !q_cam || !g_cam || !cmc_hits || !ap || !first_hit)
        return set_error(h, GRL_EINVAL, "grl_cmc_map: NULL argument");
    if (nq <= 0 || ng <= 0 || max_rank <= 0 || ld_dist < ng) return set_error(h, GRL_EINVAL, "grl_cmc_map: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    cmc_map_kernel<<<nq, CMC_THREADS, 0, st>>>(dist, ld_dist, q_pid, g_pid, q_cam, g_cam, ng, ap, first_hit);
    GRL_LAUNCH_CHECK(h);
    cmc_hits_kernel<<<(max_rank + 127) / 128, 128, 0, st>>>(first_hit, nq, max_rank, cmc_hits);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

// ---- the same over a gallery sharded across the handle's communicator
struct CmcShardedLayout { size_t pos, pos_all, cnt, total; };
static CmcShardedLayout cmc_sharded_layout(int world, int nq, int max_pos) {
    CmcShardedLayout L;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    L.pos_all = take((size_t)world * nq * (max_pos + 1) * 8);     // this rank's rows live at slot `rank` (in-place all-gather)
    L.cnt = take((size_t)nq * world * max_pos * 2 * 4);
    L.pos = 0;
    L.total = off;
    return L;
}
extern "C" size_t grl_cmc_map_sharded_workspace_bytes(const grl_handle* h, int nq, int max_pos) {
    if (!h || nq <= 0 || max_pos <= 0) return 0;
    return cmc_sharded_layout(h->comm ? h->comm_world : 1, nq, max_pos).total;
}
extern "C" int grl_cmc_map_sharded(grl_handle* h, const float* dist, long long ld_dist, const int64_t* q_pid, const int64_t* g_pid,
                                   const int64_t* q_cam, const int64_t* g_cam, int nq, int ng_local, int64_t idx_base, int max_rank, int max_pos,
                                   int32_t* cmc_hits, double* ap, int32_t* first_hit, int32_t* max_seen, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    if (!h || !dist || !q_pid || !g_pid || !q_cam || !g_cam || !cmc_hits || !ap || !first_hit || !max_seen || !workspace)
        return set_error(h, GRL_EINVAL, "grl_cmc_map_sharded: NULL argument");
    if (nq <= 0 || ng_local <= 0 || max_rank <= 0 || max_pos <= 0 || ld_dist < ng_local) return set_error(h, GRL_EINVAL, "grl_cmc_map_sharded: bad sizes");
    if (idx_base < 0 || idx_base + ng_local > 0xFFFFFFFFll) return set_error(h, GRL_EINVAL, "grl_cmc_map_sharded: global index must fit 32 bits");
    const int world = h->comm ? h->comm_world : 1, rank = h->comm ? h->comm_rank : 0;
    const CmcShardedLayout L = cmc_sharded_layout(world, nq, max_pos);
    if (workspace_bytes < L.total) return set_error(h, GRL_ENOMEM, "grl_cmc_map_sharded: workspace %zu < %zu bytes", workspace_bytes, L.total);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* w = (uint8_t*)workspace;
    unsigned long long* pos_all = (unsigned long long*)(w + L.pos_all);
    int32_t* cnt = (int32_t*)(w + L.cnt);
    const size_t per_rank = (size_t)nq * (max_pos + 1);
    unsigned long long* mine = pos_all + (size_t)rank * per_rank;
    GRL_CUDA(h, cudaMemsetAsync(max_seen, 0, 4, st));
    cmc_collect_pos_kernel<<<nq, CMC_THREADS, 0, st>>>(dist, ld_dist, q_pid, g_pid, q_cam, g_cam, ng_local, idx_base, max_pos, mine, max_seen);
    GRL_LAUNCH_CHECK(h);
    const NcclApi* api = nullptr;
    if (world > 1) {
        api = nccl_api(h);
        if (!api) return GRL_ENCCL;
        GRL_NCCL(h, api, api->AllGather(mine, pos_all, per_rank, ncclUint64, (ncclComm_t)h->comm, st));
        GRL_NCCL(h, api, api->AllReduce(max_seen, max_seen, 1, ncclInt32, ncclMax, (ncclComm_t)h->comm, st));
    }
    cmc_count_kernel<<<nq, CMC_THREADS, 0, st>>>(dist, ld_dist, q_pid, g_pid, q_cam, g_cam, ng_local, idx_base, nq, world, max_pos, pos_all, cnt);
    GRL_LAUNCH_CHECK(h);
    if (world > 1) GRL_NCCL(h, api, api->AllReduce(cnt, cnt, (size_t)nq * world * max_pos * 2, ncclInt32, ncclSum, (ncclComm_t)h->comm, st));
    cmc_reduce_kernel<<<(nq + 127) / 128, 128, 0, st>>>(nq, world, max_pos, pos_all, cnt, ap, first_hit);
    GRL_LAUNCH_CHECK(h);
    cmc_hits_kernel<<<(max_rank + 127) / 128, 128, 0, st>>>(first_hit, nq, max_rank, cmc_hits);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

extern "C" int grl_argsort_rows(grl_handle* h, const float* dist, long long ld_dist, int nq, int ng, int32_t* order, void* stream) {
    if (!h || !dist || !order) return set_error(h, GRL_EINVAL, "grl_argsort_rows: NULL argument");
    if (nq <= 0 || ng <= 0 || ng > 16384) return set_error(h, GRL_EINVAL, "grl_argsort_rows: need 0 < ng <= 16384 (ng=%d)", ng);
    const int npad = next_pow2(ng < 2 ? 2 : ng);
    const size_t smem = (size_t)npad * 8;
    GRL_TRY(ensure_dyn_smem(h, (const void*)argsort_rows_kernel, 16384 * 8));
    argsort_rows_kernel<<<nq, 1024, smem, (cudaStream_t)stream>>>(dist, ld_dist, ng, npad, order);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

extern "C" int grl_topk_init(grl_handle* h, float* top_d, int64_t* top_i, int nq, int k, void* stream) {
    if (!h || !top_d || !top_i || nq <= 0 || k <= 0) return set_error(h, GRL_EINVAL, "grl_topk_init: bad argument");
    const long long n = (long long)nq * k;
    topk_init_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(top_d, top_i, n);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

extern "C" int grl_topk_rows(grl_handle* h, const float* dist, long long ld_dist, int nq, int ncols, int k, int64_t idx_base,
                             float* top_d, int64_t* top_i, void* stream) {
    if (!h || !dist || !top_d || !top_i) return set_error(h, GRL_EINVAL, "grl_topk_rows: NULL argument");
    if (nq <= 0 || ncols <= 0 || k <= 0 || k > TOPK_MAXK) return set_error(h, GRL_EINVAL, "grl_topk_rows: need 0 < k <= %d", TOPK_MAXK);
    if (idx_base < 0 || idx_base + ncols > 0xFFFFFFFFll) return set_error(h, GRL_EINVAL, "grl_topk_rows: global index must fit 32 bits");
    topk_rows_kernel<<<nq, TOPK_THREADS, 0, (cudaStream_t)stream>>>(dist, ld_dist, ncols, k, idx_base, top_d, top_i);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

extern "C" int grl_topk_merge(grl_handle* h, const float* all_d, const int64_t* all_i, int nshards, int nq, int k, float* out_d,
                              int64_t* out_i, void* stream) {
    if (!h || !all_d || !all_i || !out_d || !out_i) return set_error(h, GRL_EINVAL, "grl_topk_merge: NULL argument");
    if (nshards <= 0 || nq <= 0 || k <= 0 || (long long)nshards * k > 16384) return set_error(h, GRL_EINVAL, "grl_topk_merge: nshards*k must be <= 16384");
    const int npad = next_pow2(nshards * k < 2 ? 2 : nshards * k);
    GRL_TRY(ensure_dyn_smem(h, (const void*)topk_merge_kernel, 16384 * 8));
    topk_merge_kernel<<<nq, 256, (size_t)npad * 8, (cudaStream_t)stream>>>(all_d, all_i, nshards, nq, k, npad, out_d, out_i);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}
