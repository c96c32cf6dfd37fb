// grl_b200 — re-ID matching kernels (sm_100a): distance via the split-bf16 tcgen05 GEMM,
// sort-free CMC / mAP, stable row argsort, streaming top-k and shard merge.
//
// Reference lines replaced:
//   cosin_dist / pairwise_distance_tensor   reid/evaluator/attevaluator.py:33-46
//   evaluate (argsort + CMC + AP)           reid/evaluator/eva_functions.py:134-184
#include <math_constants.h>

#include "api.h"

namespace grl {

// ------------------------------------------------------------------ helpers (sort keys live in common.cuh)
// In-place ascending bitonic sort of n (power of two) keys in shared memory by the whole block.
__device__ void block_bitonic_sort(uint64_t* keys, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint64_t a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    __syncthreads();
}

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
constexpr int TOPK_THREADS = 256;
constexpr int TOPK_WAVE = TOPK_THREADS * 4;
constexpr int TOPK_BUF = 2048;                 // total sort size (running list + staged candidates)
constexpr int TOPK_MAXK = 512;

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
    const int stage_cap = TOPK_BUF - TOPK_WAVE;      // flush when a full wave might not fit
    for (int c0 = 0; c0 < ncols; c0 += TOPK_WAVE) {
        const uint64_t th = thresh;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = c0 + u * TOPK_THREADS + threadIdx.x;
            if (c < ncols) {
                const uint64_t key = make_key(drow[c], (uint32_t)(idx_base + c));
                if (key < th) keys[atomicAdd(&count, 1)] = key;
            }
        }
        __syncthreads();
        const bool last = c0 + TOPK_WAVE >= ncols;
        if (count > stage_cap || (last && count > k)) {
            const int n = count;
            for (int i = n + threadIdx.x; i < TOPK_BUF; i += blockDim.x) keys[i] = ~0ull;
            block_bitonic_sort(keys, TOPK_BUF);
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

// One column chunk folded into the running lists, fed by the distance GEMM's candidate filter.  One block per query row:
//   cnt == 0          : nothing in this chunk beats the row's k-th best -> exit
//   cnt <= cap        : sort (list U candidates) with a bitonic network sized next_pow2(k + cnt) -- typically 128 keys
//   cnt  > cap        : the candidate list overflowed (first chunk, adversarial order): rescan the row of the distance tile
// and publish the new k-th best distance as the row's filter threshold.
__global__ void __launch_bounds__(TOPK_THREADS) topk_update_kernel(const float* __restrict__ dist, long long ld, int ncols, int k,
                                                                   int64_t idx_base, float* __restrict__ top_d, int64_t* __restrict__ top_i,
                                                                   float* __restrict__ thresh_out, const unsigned long long* __restrict__ cand,
                                                                   int* __restrict__ cand_cnt, int cap) {
    __shared__ uint64_t keys[TOPK_BUF];
    __shared__ int count;
    __shared__ uint64_t thresh;
    const int row = blockIdx.x;
    const int cnt = cand_cnt[row];
    if (cnt == 0) return;
    float* td = top_d + (long long)row * k;
    int64_t* ti = top_i + (long long)row * k;
    int nsort;
    if (cnt <= cap) {
        int npad = 2;
        while (npad < k + cnt) npad <<= 1;
        for (int i = threadIdx.x; i < npad; i += blockDim.x) {
            uint64_t key = ~0ull;
            if (i < k) { if (ti[i] >= 0) key = make_key(td[i], (uint32_t)ti[i]); }
            else if (i < k + cnt) key = cand[(long long)row * cap + (i - k)];
            keys[i] = key;
        }
        block_bitonic_sort(keys, npad);
        nsort = npad;
    } else {
        const float* drow = dist + (long long)row * ld;
        for (int i = threadIdx.x; i < TOPK_BUF; i += blockDim.x) {
            uint64_t key = ~0ull;
            if (i < k && ti[i] >= 0) key = make_key(td[i], (uint32_t)ti[i]);
            keys[i] = key;
        }
        if (threadIdx.x == 0) count = k;
        __syncthreads();
        if (threadIdx.x == 0) thresh = keys[k - 1];
        __syncthreads();
        const int stage_cap = TOPK_BUF - TOPK_WAVE;
        for (int c0 = 0; c0 < ncols; c0 += TOPK_WAVE) {
            const uint64_t th = thresh;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + u * TOPK_THREADS + threadIdx.x;
                if (c < ncols) {
                    const uint64_t key = make_key(drow[c], (uint32_t)(idx_base + c));
                    if (key < th) keys[atomicAdd(&count, 1)] = key;
                }
            }
            __syncthreads();
            const bool last = c0 + TOPK_WAVE >= ncols;
            if (count > stage_cap || (last && count > k)) {
                const int n = count;
                for (int i = n + threadIdx.x; i < TOPK_BUF; i += blockDim.x) keys[i] = ~0ull;
                block_bitonic_sort(keys, TOPK_BUF);
                if (threadIdx.x == 0) { count = k; thresh = keys[k - 1]; }
                __syncthreads();
            }
        }
        nsort = TOPK_BUF;
    }
    (void)nsort;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const uint64_t key = keys[i];
        if (key == ~0ull) { td[i] = CUDART_INF_F; ti[i] = -1; }
        else { td[i] = from_orderable((uint32_t)(key >> 32)); ti[i] = (int64_t)(key & 0xFFFFFFFFu); }
    }
    if (threadIdx.x == 0) {
        const uint64_t kth = keys[k - 1];
        thresh_out[row] = kth == ~0ull ? CUDART_INF_F : from_orderable((uint32_t)(kth >> 32));
        cand_cnt[row] = 0;
    }
}

// Before the first chunk no threshold exists: mark every row "overflowed" (cnt = cap + 1, threshold -inf) so the first
// update rescans its tile row and the epilogue of the first GEMM appends nothing.
__global__ void topk_filter_init_kernel(float* thresh, int* cnt, int nq, int cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) { thresh[i] = -CUDART_INF_F; cnt[i] = cap + 1; }
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

// ------------------------------------------------------------------ gallery-shard search: distance tiles + streaming top-k
static int topk_chunk_cols(int nq, int ng) {
    long long c = (1ll << 28) / (4ll * (nq > 0 ? nq : 1));       // ~256 MB distance tile
    c = c / 256 * 256;
    if (c < 1024) c = 1024;
    if (c > 16384) c = 16384;
    if (c > ng) c = (ng + 7) / 8 * 8;
    return (int)c;
}

constexpr int TOPK_CAND_CAP = 256;      // candidates per query row per column chunk before the rescan path takes over

static void dist_topk_layout(int nq, int ng, int dim, size_t* q_pl, size_t* g_pl, size_t* norms, size_t* tile, int* chunk) {
    *chunk = topk_chunk_cols(nq, ng);
    *q_pl = align_up((size_t)nq * dim * 2, 1024);
    *g_pl = align_up((size_t)*chunk * dim * 2, 1024);
    *norms = align_up((size_t)(nq + *chunk) * 4, 1024);
    *tile = align_up((size_t)nq * *chunk * 4, 1024);
}
static size_t dist_topk_filter_bytes(int nq) {   // thresh f32[nq] | cnt i32[nq] | cand u64[nq][cap]
    return align_up((size_t)nq * 4, 1024) * 2 + align_up((size_t)nq * TOPK_CAND_CAP * 8, 1024);
}

extern "C" size_t grl_dist_topk_workspace_bytes(int nq, int ng, int dim) {
    size_t a, b, c, t; int chunk;
    dist_topk_layout(nq, ng, dim, &a, &b, &c, &t, &chunk);
    return 2 * a + 2 * b + c + t + dist_topk_filter_bytes(nq);
}

extern "C" int grl_dist_topk(grl_handle* h, int metric, const float* q, const float* g, int nq, int ng, int dim, int k,
                             int64_t idx_base, float* top_d, int64_t* top_i, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !q || !g || !top_d || !top_i || !workspace) return set_error(h, GRL_EINVAL, "grl_dist_topk: NULL argument");
    if (nq <= 0 || ng <= 0 || dim <= 0 || (dim & 7)) return set_error(h, GRL_EINVAL, "grl_dist_topk: need nq,ng > 0 and dim %% 8 == 0 (dim=%d)", dim);
    if (k <= 0 || k > TOPK_MAXK) return set_error(h, GRL_EINVAL, "grl_dist_topk: need 0 < k <= %d", TOPK_MAXK);
    if (metric != GRL_METRIC_NEG_DOT && metric != GRL_METRIC_L2) return set_error(h, GRL_EINVAL, "grl_dist_topk: unknown metric %d", metric);
    if (idx_base < 0 || idx_base + ng > 0xFFFFFFFFll) return set_error(h, GRL_EINVAL, "grl_dist_topk: global index must fit 32 bits");
    if (workspace_bytes < grl_dist_topk_workspace_bytes(nq, ng, dim)) return set_error(h, GRL_ENOMEM, "grl_dist_topk: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    size_t qb, gb, nb, tb; int chunk;
    dist_topk_layout(nq, ng, dim, &qb, &gb, &nb, &tb, &chunk);
    uint8_t* w = (uint8_t*)workspace;
    __nv_bfloat16* q_hi = (__nv_bfloat16*)w; w += qb;
    __nv_bfloat16* q_lo = (__nv_bfloat16*)w; w += qb;
    __nv_bfloat16* g_hi = (__nv_bfloat16*)w; w += gb;
    __nv_bfloat16* g_lo = (__nv_bfloat16*)w; w += gb;
    float* qn = (float*)w;
    float* gn = qn + nq; w += nb;
    float* tile = (float*)w; w += tb;
    float* thresh = (float*)w; w += align_up((size_t)nq * 4, 1024);
    int* cand_cnt = (int*)w; w += align_up((size_t)nq * 4, 1024);
    unsigned long long* cand = (unsigned long long*)w;
    topk_filter_init_kernel<<<(nq + 255) / 256, 256, 0, st>>>(thresh, cand_cnt, nq, TOPK_CAND_CAP);
    GRL_LAUNCH_CHECK(h);
    GRL_TRY(split_planes(h, st, q, dim, q_hi, q_lo, dim, nq, dim));
    if (metric == GRL_METRIC_L2) {
        row_sqnorm_kernel<<<(nq * 32 + 255) / 256, 256, 0, st>>>(q, nq, dim, qn);
        GRL_LAUNCH_CHECK(h);
    }
    GRL_TRY(grl_topk_init(h, top_d, top_i, nq, k, stream));
    for (int c0 = 0; c0 < ng; c0 += chunk) {
        const int nc = (ng - c0 < chunk) ? ng - c0 : chunk;
        const float* gc = g + (size_t)c0 * dim;
        GRL_TRY(split_planes(h, st, gc, dim, g_hi, g_lo, dim, nc, dim));
        GemmEpi e = epi_default();
        e.C = tile; e.ldc = chunk;
        if (metric == GRL_METRIC_L2) {
            row_sqnorm_kernel<<<(nc * 32 + 255) / 256, 256, 0, st>>>(gc, nc, dim, gn);
            GRL_LAUNCH_CHECK(h);
            e.mode = 1; e.row_norm = qn; e.col_norm = gn;
        } else {
            e.alpha = -1.f;
        }
        // the epilogue keeps only distances that can still enter a row's list (v <= current k-th best) as candidates
        e.tk_cand = cand; e.tk_cnt = cand_cnt; e.tk_thresh = thresh; e.tk_cap = TOPK_CAND_CAP; e.tk_idx_base = idx_base + c0;
        Operand oa{q_hi, q_lo, dim, 0, 0}, ob{g_hi, g_lo, dim, 0, 0};
        GRL_TRY(gemm_launch(h, st, nq, nc, dim, 1, oa, ob, e, 0));
        topk_update_kernel<<<nq, TOPK_THREADS, 0, st>>>(tile, chunk, nc, k, idx_base + c0, top_d, top_i, thresh, cand, cand_cnt,
                                                        TOPK_CAND_CAP);
        GRL_LAUNCH_CHECK(h);
    }
    return GRL_OK;
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

static int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

extern "C" int grl_argsort_rows(grl_handle* h, const float* dist, long long ld_dist, int nq, int ng, int32_t* order, void* stream) {
    if (!h || !dist || !order) return set_error(h, GRL_EINVAL, "grl_argsort_rows: NULL argument");
    if (nq <= 0 || ng <= 0 || ng > 16384) return set_error(h, GRL_EINVAL, "grl_argsort_rows: need 0 < ng <= 16384 (ng=%d)", ng);
    const int npad = next_pow2(ng < 2 ? 2 : ng);
    const size_t smem = (size_t)npad * 8;
    static bool configured = false;
    if (!configured) {
        GRL_CUDA(h, cudaFuncSetAttribute(argsort_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        configured = true;
    }
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
    static bool configured = false;
    if (!configured) {
        GRL_CUDA(h, cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        configured = true;
    }
    topk_merge_kernel<<<nq, 256, (size_t)npad * 8, (cudaStream_t)stream>>>(all_d, all_i, nshards, nq, k, npad, out_d, out_i);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}
