// grl_b200 — re-ID matching kernels (sm_100a): distance via the split-bf16 tcgen05 GEMM,
// sort-free CMC / mAP, stable row argsort, streaming top-k and shard merge.
//
// Reference lines replaced:
//   cosin_dist / pairwise_distance_tensor   reid/evaluator/attevaluator.py:33-46
//   evaluate (argsort + CMC + AP)           reid/evaluator/eva_functions.py:134-184
#include <math_constants.h>

#include <algorithm>

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
constexpr int TOPK_BUF = 4096;                 // total sort size (running list + staged candidates)
constexpr int TOPK_MAXK = 1024;

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
        if (count > stage_cap || (boot && count > k) || (last && count > k)) {
            const int n = count;
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

constexpr int TOPK_SMALL = 32;                  // candidate counts up to this are merged by one warp per row

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
                                                                   int* __restrict__ cand_cnt, int cap, int32_t* __restrict__ dirty) {
    __shared__ uint64_t keys[TOPK_BUF];
    __shared__ int count;
    __shared__ uint64_t thresh;
    const int row = blockIdx.x;
    const int cnt = cand_cnt[row];
    if (cnt <= TOPK_SMALL) return;                  // 0: nothing to do; <= TOPK_SMALL: folded by topk_update_small_kernel
    float* td = top_d + (long long)row * k;
    int64_t* ti = top_i + (long long)row * k;
    int nsort;
    if (cnt <= cap) {
        // sort the candidates alone (cnt <= cap keys), then merge them into the sorted list by rank: a list key moves down by
        // the number of candidates below it, a candidate lands at its rank plus the number of list keys below it
        uint64_t* cs = keys;                        // [npc] sorted candidates
        uint64_t* ls = keys + TOPK_BUF / 2;         // [k]   the running list (k <= TOPK_MAXK <= TOPK_BUF / 2)
        int npc = 2;
        while (npc < cnt) npc <<= 1;
        for (int i = threadIdx.x; i < npc; i += blockDim.x) cs[i] = i < cnt ? cand[(long long)row * cap + i] : ~0ull;
        for (int i = threadIdx.x; i < k; i += blockDim.x) ls[i] = ti[i] >= 0 ? make_key(td[i], (uint32_t)ti[i]) : ~0ull;
        block_bitonic_sort(cs, npc);
        for (int i = threadIdx.x; i < k + cnt; i += blockDim.x) {
            uint64_t key;
            int pos;
            if (i < k) {
                key = ls[i];
                int lo = 0, hi = cnt;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (cs[mid] < key) lo = mid + 1; else hi = mid; }
                pos = i + lo;
                if (lo == 0) pos = (pos == k - 1) ? pos : -1 - pos;      // unmoved: nothing to write unless it defines the threshold
            } else {
                key = cs[i - k];
                int lo = 0, hi = k;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (ls[mid] < key) lo = mid + 1; else hi = mid; }
                pos = (i - k) + lo;
            }
            if (pos >= 0 && pos < k) {
                if (key == ~0ull) { td[pos] = CUDART_INF_F; ti[pos] = -1; }
                else { td[pos] = from_orderable((uint32_t)(key >> 32)); ti[pos] = (int64_t)(key & 0xFFFFFFFFu); }
                if (pos == k - 1) thresh_out[row] = key == ~0ull ? CUDART_INF_F : from_orderable((uint32_t)(key >> 32));
            }
        }
        if (threadIdx.x == 0) cand_cnt[row] = 0;
        return;
    } else {
        if (dist == nullptr) {
            // The candidate list overflowed in a chunk whose tile was not stored (only the first, threshold-less chunk is):
            // this row's list can no longer be trusted.  Mark it; finalisation sends it to the brute-force search.
            if (threadIdx.x == 0) { dirty[row] = 1; thresh_out[row] = -CUDART_INF_F; cand_cnt[row] = 0; }
            return;
        }
        const float* drow = dist + (long long)row * ld;
        if (ti[0] < 0 && ncols <= TOPK_BUF) {
            // bootstrap (empty list, first chunk): one sort of the tile row, sized to the row
            int npad = 2;
            while (npad < ncols) npad <<= 1;
            for (int i = threadIdx.x; i < npad; i += blockDim.x) keys[i] = i < ncols ? make_key(drow[i], (uint32_t)(idx_base + i)) : ~0ull;
            block_bitonic_sort(keys, npad);
            for (int i = threadIdx.x; i < k; i += blockDim.x) {
                const uint64_t key = i < npad ? keys[i] : ~0ull;
                if (key == ~0ull) { td[i] = CUDART_INF_F; ti[i] = -1; }
                else { td[i] = from_orderable((uint32_t)(key >> 32)); ti[i] = (int64_t)(key & 0xFFFFFFFFu); }
            }
            if (threadIdx.x == 0) {
                const uint64_t kth = (k - 1 < npad) ? keys[k - 1] : ~0ull;
                thresh_out[row] = kth == ~0ull ? CUDART_INF_F : from_orderable((uint32_t)(kth >> 32));
                cand_cnt[row] = 0;
            }
            return;
        }
        for (int i = threadIdx.x; i < TOPK_BUF; i += blockDim.x) {
            uint64_t key = ~0ull;
            if (i < k && ti[i] >= 0) key = make_key(td[i], (uint32_t)ti[i]);
            keys[i] = key;
        }
        if (threadIdx.x == 0) count = k;
        __syncthreads();
        if (threadIdx.x == 0) thresh = keys[k - 1];
        __syncthreads();
        const int stage_cap = k <= 256 ? TOPK_WAVE : TOPK_BUF - TOPK_WAVE;
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
                int npad = 2;
                while (npad < n) npad <<= 1;
                for (int i = n + threadIdx.x; i < npad; i += blockDim.x) keys[i] = ~0ull;
                block_bitonic_sort(keys, npad);
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

// The common case once the thresholds have tightened: a handful of candidates per row and chunk.  One WARP per row, no block
// barriers: the candidates are sorted across the lanes with a shuffle network and merged into the (sorted) running list by
// rank -- every list key moves down by the number of candidates below it, every candidate lands at its own rank plus the
// number of list keys below it -- so the cost is one pass over the list instead of a 512-key bitonic sort.
__global__ void __launch_bounds__(256) topk_update_small_kernel(int nq, int k, float* __restrict__ top_d, int64_t* __restrict__ top_i,
                                                                float* __restrict__ thresh_out, const unsigned long long* __restrict__ cand,
                                                                int* __restrict__ cand_cnt, int cap) {
    extern __shared__ uint64_t small_keys[];          // [8 warps][k]
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int row = blockIdx.x * 8 + warp;
    if (row >= nq) return;
    const int cnt = cand_cnt[row];
    if (cnt == 0 || cnt > TOPK_SMALL) return;
    uint64_t* sl = small_keys + (size_t)warp * k;
    float* td = top_d + (long long)row * k;
    int64_t* ti = top_i + (long long)row * k;
    for (int t = lane; t < k; t += 32) sl[t] = ti[t] >= 0 ? make_key(td[t], (uint32_t)ti[t]) : ~0ull;
    uint64_t c = lane < cnt ? cand[(long long)row * cap + lane] : ~0ull;
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {      // bitonic sort across the lanes, ascending
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const uint64_t o = __shfl_xor_sync(0xffffffffu, c, stride);
            const bool up = (lane & size) == 0;
            const bool lower = (lane & stride) == 0;
            c = ((c < o) == (up == lower)) ? c : o;
        }
    }
    __syncwarp();
    // candidates: rank among the list keys (keys are unique: the index is part of the key)
    if (lane < cnt) {
        int lo = 0, hi = k;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sl[mid] < c) lo = mid + 1; else hi = mid; }
        const int pos = lo + lane;
        if (pos < k) {
            td[pos] = from_orderable((uint32_t)(c >> 32)); ti[pos] = (int64_t)(c & 0xFFFFFFFFu);
            if (pos == k - 1) thresh_out[row] = td[pos];
        }
    }
    // list keys: shifted down by the number of candidates below them
    for (int t = lane; t < k; t += 32) {
        const uint64_t key = sl[t];
        int below = 0;
        for (int j = 0; j < cnt; ++j) below += (__shfl_sync(0xffffffffu, c, j) < key) ? 1 : 0;
        const int pos = t + below;
        if (below > 0 && pos < k) {
            if (key == ~0ull) { td[pos] = CUDART_INF_F; ti[pos] = -1; }
            else { td[pos] = from_orderable((uint32_t)(key >> 32)); ti[pos] = (int64_t)(key & 0xFFFFFFFFu); }
        }
        if (pos == k - 1) thresh_out[row] = key == ~0ull ? CUDART_INF_F : from_orderable((uint32_t)(key >> 32));
    }
    if (lane == 0) cand_cnt[row] = 0;
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


// ------------------------------------------------------------------ two-stage exact search (retrieval)
// Stage 1 ranks every (query, gallery row) pair by a COARSE distance: operands rounded to fp16 after a per-row power-of-two
// scaling, one tcgen05 MMA per k-step instead of three.  For any pair the coarse and the exact inner product differ by at most
//        E(q, g) = CE * |q| * |g|,     CE = 2^-10 + 2^-17 + 2 * dim * 2^-24
// (two fp16 roundings of <= 2^-11 relative each, their product term, subnormal/flush slack, and fp32 accumulation of `dim` terms
// on either side), so the K' coarse-nearest rows of a query contain its k exact-nearest whenever
//        coarse[K'-th] - E_max  >  exact[k-th among the K' re-scored]                                       (*)
// Stage 2 re-scores the K' candidates with a fixed-order fp32 inner product and checks (*) per query; queries that fail it
// (near-duplicate galleries, k close to K') are searched by brute force in the same fixed-order arithmetic.  The result is
// therefore the exact (distance, index) top-k of the fixed-order fp32 distances, independent of chunking and of the shard count.

// fp32 row -> fp16 row scaled by a power of two so that max|x| lands in [2^14, 2^15); one warp per row.
// inv_scale[row] = 1/scale, sqnorm[row] = |x|^2 (fixed order), gmax2 = max over rows of sqnorm (uint-ordered atomicMax).
__global__ void f16_rows_kernel(const float* __restrict__ x, long long rows, int dim, __half* __restrict__ out,
                                float* __restrict__ inv_scale, float* __restrict__ sqnorm, unsigned int* __restrict__ gmax2) {
    const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const int lane = lane_id();
    const float4* r4 = reinterpret_cast<const float4*>(x + row * dim);
    const int nv = dim >> 2;
    float m = 0.f, acc = 0.f;
    for (int j = lane; j < nv; j += 32) {
        const float4 v = __ldg(r4 + j);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        acc = __fadd_rn(acc, __fmul_rn(v.x, v.x)); acc = __fadd_rn(acc, __fmul_rn(v.y, v.y));
        acc = __fadd_rn(acc, __fmul_rn(v.z, v.z)); acc = __fadd_rn(acc, __fmul_rn(v.w, v.w));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
    }
    int e = 0;
    if (m > 0.f && m < CUDART_INF_F) e = (int)((__float_as_uint(m) >> 23) & 0xff) - 127;
    e = max(-100, min(100, e));
    const float s = ldexpf(1.f, 14 - e);
    __half2* o2 = reinterpret_cast<__half2*>(out + row * dim);
    for (int j = lane; j < nv; j += 32) {
        const float4 v = __ldg(r4 + j);
        o2[2 * j] = __floats2half2_rn(v.x * s, v.y * s);
        o2[2 * j + 1] = __floats2half2_rn(v.z * s, v.w * s);
    }
    if (lane == 0) {
        inv_scale[row] = ldexpf(1.f, e - 14);
        sqnorm[row] = acc;
        atomicMax(gmax2, __float_as_uint(acc));
    }
}

// Fixed-order fp32 inner product of a shared-memory query row with a global gallery row, by one warp: lane l accumulates the
// float4 chunks l, l+32, ... component by component (separate multiply and add, no FMA), then an xor-shuffle tree.  Also returns
// |g|^2 in the same order.  The parity tests restate this order in numpy.
__device__ __forceinline__ void warp_dot_fixed(const float* __restrict__ qs, const float* __restrict__ g, int dim, float& dot, float& gg) {
    const int lane = lane_id();
    const float4* q4 = reinterpret_cast<const float4*>(qs);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const int nv = dim >> 2;
    float a = 0.f, b = 0.f;
    for (int j = lane; j < nv; j += 32) {
        const float4 x = q4[j];
        const float4 y = __ldg(g4 + j);
        a = __fadd_rn(a, __fmul_rn(x.x, y.x)); a = __fadd_rn(a, __fmul_rn(x.y, y.y));
        a = __fadd_rn(a, __fmul_rn(x.z, y.z)); a = __fadd_rn(a, __fmul_rn(x.w, y.w));
        b = __fadd_rn(b, __fmul_rn(y.x, y.x)); b = __fadd_rn(b, __fmul_rn(y.y, y.y));
        b = __fadd_rn(b, __fmul_rn(y.z, y.z)); b = __fadd_rn(b, __fmul_rn(y.w, y.w));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, off));
        b = __fadd_rn(b, __shfl_xor_sync(0xffffffffu, b, off));
    }
    dot = a; gg = b;
}
__device__ __forceinline__ float exact_distance(int metric, float dot, float qq, float gg) {
    if (metric == GRL_METRIC_L2) return sqrtf(fmaxf(__fsub_rn(__fadd_rn(qq, gg), __fmul_rn(2.f, dot)), 1e-12f));
    return -dot;
}
// |q|^2 of the shared-memory row in the same fixed order (every warp computes the same value)
__device__ __forceinline__ float warp_sqnorm_fixed(const float* __restrict__ qs, int dim) {
    const int lane = lane_id();
    const float4* q4 = reinterpret_cast<const float4*>(qs);
    float b = 0.f;
    for (int j = lane; j < (dim >> 2); j += 32) {
        const float4 y = q4[j];
        b = __fadd_rn(b, __fmul_rn(y.x, y.x)); b = __fadd_rn(b, __fmul_rn(y.y, y.y));
        b = __fadd_rn(b, __fmul_rn(y.z, y.z)); b = __fadd_rn(b, __fmul_rn(y.w, y.w));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) b = __fadd_rn(b, __shfl_xor_sync(0xffffffffu, b, off));
    return b;
}

// exact_d[row][t] = exact distance of query `row` to candidate cand_i[row][t] when that gallery row lives in this shard
// ([idx_base, idx_base + ng)), else 0 -- so that the per-shard results combine by a plain sum.  One block per query.
__global__ void __launch_bounds__(256) rescore_kernel(int metric, const float* __restrict__ q, const float* __restrict__ g, int ng, int dim,
                                                      long long idx_base, const int64_t* __restrict__ cand_i, int kp,
                                                      float* __restrict__ exact_d) {
    extern __shared__ float qs[];
    const int row = blockIdx.x;
    for (int j = threadIdx.x; j < (dim >> 2); j += blockDim.x)
        reinterpret_cast<float4*>(qs)[j] = __ldg(reinterpret_cast<const float4*>(q + (long long)row * dim) + j);
    __syncthreads();
    const float qq = (metric == GRL_METRIC_L2) ? warp_sqnorm_fixed(qs, dim) : 0.f;
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int t = warp; t < kp; t += nw) {
        const long long idx = cand_i[(long long)row * kp + t] - idx_base;
        float d = 0.f;
        if (idx >= 0 && idx < ng) {
            float dot, gg;
            warp_dot_fixed(qs, g + idx * dim, dim, dot, gg);
            d = exact_distance(metric, dot, qq, gg);
        }
        if (lane_id() == 0) exact_d[(long long)row * kp + t] = d;
    }
}

// Sort the K' re-scored candidates by (exact distance, index), emit the top k, and check condition (*) above.
// coarse_d holds -q.g (metric 0) or the SQUARED L2 distance (metric 1) of the coarse pass, ascending; gmax2 = max |g|^2.
__global__ void __launch_bounds__(256) topk_finalize_kernel(int metric, const float* __restrict__ q, int dim, const float* __restrict__ coarse_d,
                                                            const int64_t* __restrict__ cand_i, const float* __restrict__ exact_d, int kp,
                                                            int npad, const float* __restrict__ gmax2, float ce, int k,
                                                            const int32_t* __restrict__ dirty, float* __restrict__ top_d,
                                                            int64_t* __restrict__ top_i, int32_t* __restrict__ flags,
                                                            int32_t* __restrict__ nflag) {
    extern __shared__ uint64_t fkeys[];
    __shared__ float red[8];
    __shared__ int nvalid_s;
    const int row = blockIdx.x;
    if (threadIdx.x == 0) nvalid_s = 0;
    __syncthreads();
    int local_valid = 0;
    for (int t = threadIdx.x; t < npad; t += blockDim.x) {
        uint64_t key = ~0ull;
        if (t < kp) {
            const int64_t idx = cand_i[(long long)row * kp + t];
            if (idx >= 0) { key = make_key(exact_d[(long long)row * kp + t], (uint32_t)idx); ++local_valid; }
        }
        fkeys[t] = key;
    }
    if (local_valid) atomicAdd(&nvalid_s, local_valid);
    // |q|^2 (any order: only an upper bound is needed, a relative 1e-5 is added below)
    float qq = 0.f;
    for (int j = threadIdx.x; j < dim; j += blockDim.x) { const float v = q[(long long)row * dim + j]; qq += v * v; }
    qq = warp_sum(qq);
    if (lane_id() == 0) red[threadIdx.x >> 5] = qq;
    block_bitonic_sort(fkeys, npad);                  // also orders the writes above before the reads below
    qq = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) qq += red[w];
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const uint64_t key = fkeys[t];
        if (key == ~0ull) { top_d[(long long)row * k + t] = CUDART_INF_F; top_i[(long long)row * k + t] = -1; }
        else { top_d[(long long)row * k + t] = from_orderable((uint32_t)(key >> 32)); top_i[(long long)row * k + t] = (int64_t)(key & 0xFFFFFFFFu); }
    }
    if (threadIdx.x == 0) {
        bool ok = !(dirty && dirty[row]);             // a shard's candidate list overflowed: not provable
        const int nvalid = nvalid_s;
        if (ok && nvalid == kp && nvalid > k - 1 && fkeys[k - 1] != ~0ull) {   // a full list may have cut off relevant rows
            const float u = from_orderable((uint32_t)(fkeys[k - 1] >> 32));
            const float ck = coarse_d[(long long)row * kp + kp - 1];
            const float g2 = *gmax2;
            const float emax = ce * sqrtf(qq) * sqrtf(g2) * 1.00001f;
            if (metric == GRL_METRIC_L2) {
                const float lb = ck - 2.f * emax - 1e-5f * (qq + g2);
                ok = lb > 1e-12f && sqrtf(lb) * 0.999999f > u;
            } else {
                ok = ck - emax - 1e-6f * fabsf(ck) > u;
            }
            if (!(ok)) ok = false;                    // NaN anywhere -> brute force
        }
        flags[row] = ok ? 0 : 1;
        if (!ok) atomicAdd(nflag, 1);
    }
}

// Brute force in the fixed-order arithmetic: tile[r][c] = exact distance of query rows[r] (or r when rows == NULL) to gallery
// row c.  Each block keeps RQ query rows in shared memory and streams a slab of gallery rows once, one warp per gallery row.
__global__ void __launch_bounds__(256) exact_rows_kernel(int metric, const float* __restrict__ q, const int32_t* __restrict__ rows, int r0, int rq,
                                                         const float* __restrict__ g, int ng, int dim, float* __restrict__ tile, long long ld_tile) {
    extern __shared__ float qs[];                     // [rq][dim]
    __shared__ float qqs[16];
    for (int r = 0; r < rq; ++r) {
        const long long src = rows ? rows[r0 + r] : (r0 + r);
        for (int j = threadIdx.x; j < (dim >> 2); j += blockDim.x)
            reinterpret_cast<float4*>(qs + (long long)r * dim)[j] = __ldg(reinterpret_cast<const float4*>(q + src * dim) + j);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (warp == 0) for (int r = 0; r < rq; ++r) { const float v = warp_sqnorm_fixed(qs + (long long)r * dim, dim); if (lane_id() == 0) qqs[r] = v; }
    __syncthreads();
    for (long long c = (long long)blockIdx.x * nw + warp; c < ng; c += (long long)gridDim.x * nw) {
        for (int r = 0; r < rq; ++r) {                // the gallery row stays in L1 across the rq passes
            float dot, gg;
            warp_dot_fixed(qs + (long long)r * dim, g + c * dim, dim, dot, gg);
            if (lane_id() == 0) tile[(long long)r * ld_tile + c] = exact_distance(metric, dot, qqs[r], gg);
        }
    }
}

__global__ void flagged_rows_kernel(const int32_t* __restrict__ flags, int nq, int32_t* __restrict__ rows, int32_t* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq && flags[i]) rows[atomicAdd(count, 1)] = i;
}
__global__ void scatter_rows_kernel(const float* __restrict__ src_d, const int64_t* __restrict__ src_i, const int32_t* __restrict__ rows, int r0,
                                    int k, float* __restrict__ dst_d, int64_t* __restrict__ dst_i) {
    const int r = blockIdx.x;
    const long long dst = (long long)rows[r0 + r] * k;
    for (int t = threadIdx.x; t < k; t += blockDim.x) { dst_d[dst + t] = src_d[(long long)r * k + t]; dst_i[dst + t] = src_i[(long long)r * k + t]; }
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

// ------------------------------------------------------------------ gallery-shard search: coarse tiles + streaming top-K', exact re-score
// Column chunks of one search.  The first chunk has no thresholds yet: its tile IS stored and every row is rescanned, so it is
// kept small (TOPK_FIRST_CHUNK columns, one 1024-key sort per row).  Afterwards the K'-th best of n_seen columns lets ~K' * nc / n_seen candidates per row
// through, so chunks grow with n_seen (at most doubling the columns seen) up to the steady-state size, whose 256 x 256 tiles
// fill whole waves of the persistent grid; their tiles are never stored.
constexpr int TOPK_FIRST_CHUNK = 1024;
constexpr int TOPK_CAND_CAP = 512;      // candidates per query row per column chunk (expected <= K' = 256..1024 / growth factor)

static int topk_chunk_max(int nq, int ng, int num_sms) {
    long long c = 16384;
    if (nq >= 1024 && c < ng) {
        const long long mt = (nq + 255) / 256;
        double best = 0.0; long long best_n = c / 256;
        for (long long n = c / 256; n >= 16; --n) {
            const long long tiles = mt * n, waves = (tiles + num_sms - 1) / num_sms;
            const double eff = (double)tiles / (double)(waves * num_sms);
            if (eff > best + 1e-9) { best = eff; best_n = n; }
        }
        c = best_n * 256;
    }
    if (c > ng) c = (ng + 7) / 8 * 8;
    return (int)c;
}
static int topk_next_chunk(int c0, int ng, int chunk_max) {
    long long nc = c0 == 0 ? TOPK_FIRST_CHUNK : (c0 < chunk_max ? c0 : chunk_max);
    nc = (nc + 255) / 256 * 256;
    if (nc > chunk_max) nc = chunk_max;
    if (nc > ng - c0) nc = ng - c0;
    return (int)nc;
}

extern "C" int grl_topk_kprime(int k) { return k <= 128 ? 256 : (k <= 256 ? 512 : 1024); }

static float coarse_error_constant(int dim) {   // CE of the comment above f16_rows_kernel
    return 0x1p-10f + 0x1p-17f + 2.f * (float)dim * 0x1p-24f;
}

struct CoarseLayout {
    int chunk, first;
    size_t q16, g16, qf, gf, tile, thresh, cnt, cand, total;
};
static void coarse_layout(int nq, int ng, int dim, CoarseLayout* L) {
    L->chunk = topk_chunk_max(nq, ng, 148);
    L->first = ng < TOPK_FIRST_CHUNK ? (ng + 7) / 8 * 8 : TOPK_FIRST_CHUNK;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    L->q16 = take((size_t)nq * dim * 2);
    L->g16 = take((size_t)L->chunk * dim * 2);
    L->qf = take((size_t)nq * 4 * 2);                 // inv_scale | sqnorm
    L->gf = take((size_t)L->chunk * 4 * 2);
    L->tile = take((size_t)nq * L->first * 4);        // coarse distances of the first chunk only
    L->thresh = take((size_t)nq * 4);
    L->cnt = take((size_t)nq * 4);
    L->cand = take((size_t)nq * TOPK_CAND_CAP * 8);
    L->total = off;
}

extern "C" size_t grl_coarse_topk_workspace_bytes(int nq, int ng, int dim) {
    CoarseLayout L;
    coarse_layout(nq, ng, dim, &L);
    return L.total;
}

// A gallery shard converted once (fp16 rows with their per-row scales, squared norms, the largest squared norm): searches
// against a static gallery skip the per-chunk conversion.
struct PreparedLayout { size_t g16, inv, n2, gmax2, total; };
static PreparedLayout prepared_layout(int ng, int dim) {
    PreparedLayout P;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    P.g16 = take((size_t)ng * dim * 2);
    P.inv = take((size_t)ng * 4);
    P.n2 = take((size_t)ng * 4);
    P.gmax2 = take(4);
    P.total = off;
    return P;
}
extern "C" size_t grl_gallery_prepared_bytes(int ng, int dim) {
    if (ng <= 0 || dim <= 0 || (dim & 7)) return 0;
    return prepared_layout(ng, dim).total;
}
extern "C" int grl_gallery_prepare(grl_handle* h, const float* g, int ng, int dim, void* prepared, size_t prepared_bytes, void* stream) {
    if (!h || !g || !prepared) return set_error(h, GRL_EINVAL, "grl_gallery_prepare: NULL argument");
    if (ng <= 0 || dim <= 0 || (dim & 7)) return set_error(h, GRL_EINVAL, "grl_gallery_prepare: need ng > 0 and dim %% 8 == 0 (dim=%d)", dim);
    const PreparedLayout P = prepared_layout(ng, dim);
    if (prepared_bytes < P.total) return set_error(h, GRL_ENOMEM, "grl_gallery_prepare: buffer %zu < %zu bytes", prepared_bytes, P.total);
    if (reinterpret_cast<uintptr_t>(prepared) & 255) return set_error(h, GRL_EINVAL, "grl_gallery_prepare: buffer must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* b = (uint8_t*)prepared;
    GRL_CUDA(h, cudaMemsetAsync(b + P.gmax2, 0, 4, st));
    f16_rows_kernel<<<(int)(((long long)ng * 32 + 255) / 256), 256, 0, st>>>(g, ng, dim, (__half*)(b + P.g16), (float*)(b + P.inv), (float*)(b + P.n2),
                                                                             (unsigned int*)(b + P.gmax2));
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

static int coarse_topk_impl(grl_handle* h, int metric, const float* q, const float* g, const void* prepared, int nq, int ng, int dim, int kprime,
                            int64_t idx_base, float* coarse_d, int64_t* coarse_i, float* gmax2, int32_t* dirty, void* workspace,
                            size_t workspace_bytes, void* stream) {
    if (!h || !q || (!g && !prepared) || !coarse_d || !coarse_i || !gmax2 || !dirty || !workspace) return set_error(h, GRL_EINVAL, "grl_coarse_topk: NULL argument");
    if (nq <= 0 || ng <= 0 || dim <= 0 || (dim & 7)) return set_error(h, GRL_EINVAL, "grl_coarse_topk: need nq,ng > 0 and dim %% 8 == 0 (dim=%d)", dim);
    if (kprime <= 0 || kprime > TOPK_MAXK) return set_error(h, GRL_EINVAL, "grl_coarse_topk: need 0 < kprime <= %d", TOPK_MAXK);
    if (metric != GRL_METRIC_NEG_DOT && metric != GRL_METRIC_L2) return set_error(h, GRL_EINVAL, "grl_coarse_topk: unknown metric %d", metric);
    if (idx_base < 0 || idx_base + ng > 0xFFFFFFFFll) return set_error(h, GRL_EINVAL, "grl_coarse_topk: global index must fit 32 bits");
    CoarseLayout L;
    coarse_layout(nq, ng, dim, &L);
    if (workspace_bytes < L.total) return set_error(h, GRL_ENOMEM, "grl_coarse_topk: workspace %zu < %zu bytes", workspace_bytes, L.total);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* w = (uint8_t*)workspace;
    __half* q16 = (__half*)(w + L.q16);
    __half* g16 = (__half*)(w + L.g16);
    float* q_inv = (float*)(w + L.qf);
    float* q_n2 = q_inv + nq;
    float* g_inv = (float*)(w + L.gf);
    float* g_n2 = g_inv + L.chunk;
    float* tile = (float*)(w + L.tile);
    float* thresh = (float*)(w + L.thresh);
    int* cand_cnt = (int*)(w + L.cnt);
    unsigned long long* cand = (unsigned long long*)(w + L.cand);
    const PreparedLayout P = prepared_layout(ng, dim);
    const uint8_t* pb = (const uint8_t*)prepared;
    if (prepared) GRL_CUDA(h, cudaMemcpyAsync(gmax2, pb + P.gmax2, 4, cudaMemcpyDeviceToDevice, st));
    else GRL_CUDA(h, cudaMemsetAsync(gmax2, 0, 4, st));
    GRL_CUDA(h, cudaMemsetAsync(dirty, 0, (size_t)nq * 4, st));
    GRL_TRY(ensure_dyn_smem(h, (const void*)topk_update_small_kernel, 8 * kprime * 8));
    topk_filter_init_kernel<<<(nq + 255) / 256, 256, 0, st>>>(thresh, cand_cnt, nq, TOPK_CAND_CAP);
    GRL_LAUNCH_CHECK(h);
    // the query norms go to a scratch max (cand is free until the first GEMM): only the gallery maximum is reported
    f16_rows_kernel<<<(int)(((long long)nq * 32 + 255) / 256), 256, 0, st>>>(q, nq, dim, q16, q_inv, q_n2, (unsigned int*)cand);
    GRL_LAUNCH_CHECK(h);
    GRL_TRY(grl_topk_init(h, coarse_d, coarse_i, nq, kprime, stream));
    for (int c0 = 0; c0 < ng;) {
        const int nc = topk_next_chunk(c0, ng, L.chunk);
        const bool first = c0 == 0;
        const __half* g16c = g16;
        const float* g_invc = g_inv;
        const float* g_n2c = g_n2;
        if (prepared) {                               // chunk starts are multiples of 256 columns: the slices stay 16-byte aligned
            g16c = (const __half*)(pb + P.g16) + (size_t)c0 * dim;
            g_invc = (const float*)(pb + P.inv) + c0;
            g_n2c = (const float*)(pb + P.n2) + c0;
        } else {
            f16_rows_kernel<<<(int)(((long long)nc * 32 + 255) / 256), 256, 0, st>>>(g + (size_t)c0 * dim, nc, dim, g16, g_inv, g_n2,
                                                                                     (unsigned int*)gmax2);
            GRL_LAUNCH_CHECK(h);
        }
        GemmEpi e = epi_default();
        if (first) { e.C = tile; e.ldc = L.first; }   // later chunks never store their tile
        e.row_scale = q_inv; e.col_scale = g_invc;
        if (metric == GRL_METRIC_L2) { e.mode = 2; e.row_norm = q_n2; e.col_norm = g_n2c; }
        else e.alpha = -1.f;
        // the epilogue keeps only distances that can still enter a row's list (v <= current K'-th best) as candidates
        e.tk_cand = cand; e.tk_cnt = cand_cnt; e.tk_thresh = thresh; e.tk_cap = TOPK_CAND_CAP; e.tk_idx_base = idx_base + c0;
        GRL_TRY(coarse_gemm_launch(h, st, nq, nc, dim, q16, dim, g16c, dim, e));
        topk_update_small_kernel<<<(nq + 7) / 8, 256, (size_t)8 * kprime * 8, st>>>(nq, kprime, coarse_d, coarse_i, thresh, cand, cand_cnt,
                                                                                     TOPK_CAND_CAP);
        GRL_LAUNCH_CHECK(h);
        topk_update_kernel<<<nq, TOPK_THREADS, 0, st>>>(first ? tile : nullptr, L.first, nc, kprime, idx_base + c0, coarse_d, coarse_i, thresh,
                                                        cand, cand_cnt, TOPK_CAND_CAP, dirty);
        GRL_LAUNCH_CHECK(h);
        c0 += nc;
    }
    return GRL_OK;
}

extern "C" int grl_coarse_topk(grl_handle* h, int metric, const float* q, const float* g, int nq, int ng, int dim, int kprime,
                               int64_t idx_base, float* coarse_d, int64_t* coarse_i, float* gmax2, int32_t* dirty, void* workspace,
                               size_t workspace_bytes, void* stream) {
    if (!g) return set_error(h, GRL_EINVAL, "grl_coarse_topk: NULL argument");
    return coarse_topk_impl(h, metric, q, g, nullptr, nq, ng, dim, kprime, idx_base, coarse_d, coarse_i, gmax2, dirty, workspace, workspace_bytes,
                            stream);
}
extern "C" int grl_coarse_topk_prepared(grl_handle* h, int metric, const float* q, const void* prepared, int nq, int ng, int dim, int kprime,
                                        int64_t idx_base, float* coarse_d, int64_t* coarse_i, float* gmax2, int32_t* dirty,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    if (!prepared) return set_error(h, GRL_EINVAL, "grl_coarse_topk_prepared: NULL argument");
    return coarse_topk_impl(h, metric, q, nullptr, prepared, nq, ng, dim, kprime, idx_base, coarse_d, coarse_i, gmax2, dirty, workspace,
                            workspace_bytes, stream);
}

extern "C" int grl_rescore(grl_handle* h, int metric, const float* q, const float* g, int nq, int ng, int dim, int64_t idx_base,
                           const int64_t* cand_i, int kprime, float* exact_d, void* stream) {
    if (!h || !q || !g || !cand_i || !exact_d) return set_error(h, GRL_EINVAL, "grl_rescore: NULL argument");
    if (nq <= 0 || ng <= 0 || dim <= 0 || (dim & 7) || dim > 32768 || kprime <= 0) return set_error(h, GRL_EINVAL, "grl_rescore: bad sizes");
    const size_t smem = (size_t)dim * 4;
    GRL_TRY(ensure_dyn_smem(h, (const void*)rescore_kernel, (int)smem));
    rescore_kernel<<<nq, 256, smem, (cudaStream_t)stream>>>(metric, q, g, ng, dim, idx_base, cand_i, kprime, exact_d);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

extern "C" int grl_topk_finalize(grl_handle* h, int metric, const float* q, int nq, int dim, const float* coarse_d, const int64_t* cand_i,
                                 const float* exact_d, int kprime, const float* gmax2, const int32_t* dirty, int k, float* top_d,
                                 int64_t* top_i, int32_t* flags, int32_t* nflag, void* stream) {
    if (!h || !q || !coarse_d || !cand_i || !exact_d || !gmax2 || !top_d || !top_i || !flags || !nflag)
        return set_error(h, GRL_EINVAL, "grl_topk_finalize: NULL argument");
    if (nq <= 0 || dim <= 0 || kprime <= 0 || kprime > TOPK_MAXK || k <= 0 || k > kprime) return set_error(h, GRL_EINVAL, "grl_topk_finalize: need 0 < k <= kprime <= %d", TOPK_MAXK);
    cudaStream_t st = (cudaStream_t)stream;
    GRL_CUDA(h, cudaMemsetAsync(nflag, 0, 4, st));
    const int npad = next_pow2(kprime < 2 ? 2 : kprime);
    topk_finalize_kernel<<<nq, 256, (size_t)npad * 8, st>>>(metric, q, dim, coarse_d, cand_i, exact_d, kprime, npad, gmax2,
                                                            coarse_error_constant(dim), k, dirty, top_d, top_i, flags, nflag);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

// ---- brute force in the fixed-order arithmetic (fallback of the two-stage search; any query subset)
static int exact_group_rows(int dim) {
    int r = (160 * 1024) / (dim * 4);
    return r < 1 ? 1 : (r > 8 ? 8 : r);
}
extern "C" size_t grl_exact_topk_workspace_bytes(int nq, int ng, int dim) {
    const int R = std::min(exact_group_rows(dim), nq > 0 ? nq : 1);
    return align_up((size_t)R * ng * 4, 1024);
}
static int exact_topk_rows(grl_handle* h, int metric, const float* q, const int32_t* rows, int nrows, const float* g, int ng, int dim, int k,
                           int64_t idx_base, float* top_d, int64_t* top_i, float* tmp_d, int64_t* tmp_i, float* tile, cudaStream_t st) {
    // rows == NULL: query r is row r of q and results go to top_d/top_i[r]; else query rows[r], results scattered to row rows[r]
    const int R = exact_group_rows(dim);
    const size_t smem = (size_t)R * dim * 4;
    GRL_TRY(ensure_dyn_smem(h, (const void*)exact_rows_kernel, (int)smem));
    for (int r0 = 0; r0 < nrows; r0 += R) {
        const int rq = std::min(R, nrows - r0);
        exact_rows_kernel<<<h->num_sms * 4, 256, (size_t)rq * dim * 4, st>>>(metric, q, rows, r0, rq, g, ng, dim, tile, ng);
        GRL_LAUNCH_CHECK(h);
        float* od = rows ? tmp_d : top_d + (size_t)r0 * k;
        int64_t* oi = rows ? tmp_i : top_i + (size_t)r0 * k;
        GRL_TRY(grl_topk_init(h, od, oi, rq, k, st));
        GRL_TRY(grl_topk_rows(h, tile, ng, rq, ng, k, idx_base, od, oi, st));
        if (rows) {
            scatter_rows_kernel<<<rq, 128, 0, st>>>(tmp_d, tmp_i, rows, r0, k, top_d, top_i);
            GRL_LAUNCH_CHECK(h);
        }
    }
    return GRL_OK;
}
extern "C" int grl_exact_topk(grl_handle* h, int metric, const float* q, const float* g, int nq, int ng, int dim, int k, int64_t idx_base,
                              float* top_d, int64_t* top_i, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !q || !g || !top_d || !top_i || !workspace) return set_error(h, GRL_EINVAL, "grl_exact_topk: NULL argument");
    if (nq <= 0 || ng <= 0 || dim <= 0 || (dim & 7) || dim > 32768) return set_error(h, GRL_EINVAL, "grl_exact_topk: need nq,ng > 0, dim %% 8 == 0, dim <= 32768");
    if (k <= 0 || k > TOPK_MAXK) return set_error(h, GRL_EINVAL, "grl_exact_topk: need 0 < k <= %d", TOPK_MAXK);
    if (metric != GRL_METRIC_NEG_DOT && metric != GRL_METRIC_L2) return set_error(h, GRL_EINVAL, "grl_exact_topk: unknown metric %d", metric);
    if (idx_base < 0 || idx_base + ng > 0xFFFFFFFFll) return set_error(h, GRL_EINVAL, "grl_exact_topk: global index must fit 32 bits");
    if (workspace_bytes < grl_exact_topk_workspace_bytes(nq, ng, dim)) return set_error(h, GRL_ENOMEM, "grl_exact_topk: workspace too small");
    return exact_topk_rows(h, metric, q, nullptr, nq, g, ng, dim, k, idx_base, top_d, top_i, nullptr, nullptr, (float*)workspace, (cudaStream_t)stream);
}

// ---- one shard, end to end
struct DistTopkLayout {
    size_t coarse, cd, ci, ed, misc, flags, rows, dirty, tmp_d, tmp_i, brute, total;
    int kp;
};
static void dist_topk_layout(int nq, int ng, int dim, int k, DistTopkLayout* L) {
    L->kp = grl_topk_kprime(k);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    L->coarse = take(grl_coarse_topk_workspace_bytes(nq, ng, dim));
    L->cd = take((size_t)nq * L->kp * 4);
    L->ci = take((size_t)nq * L->kp * 8);
    L->ed = take((size_t)nq * L->kp * 4);
    L->misc = take(64);                               // gmax2 f32 | nflag i32 | nrows i32
    L->flags = take((size_t)nq * 4);
    L->rows = take((size_t)nq * 4);
    L->dirty = take((size_t)nq * 4);
    L->tmp_d = take((size_t)8 * k * 4);
    L->tmp_i = take((size_t)8 * k * 8);
    L->brute = take(grl_exact_topk_workspace_bytes(nq, ng, dim));
    L->total = off;
}

extern "C" size_t grl_dist_topk_workspace_bytes(int nq, int ng, int dim) {
    DistTopkLayout L;
    dist_topk_layout(nq, ng, dim, TOPK_MAXK / 2, &L);   // sized for the largest supported k
    return L.total;
}

static int dist_topk_impl(grl_handle* h, int metric, const float* q, const float* g, const void* prepared, int nq, int ng, int dim, int k,
                          int64_t idx_base, float* top_d, int64_t* top_i, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !q || !g || !top_d || !top_i || !workspace) return set_error(h, GRL_EINVAL, "grl_dist_topk: NULL argument");
    if (nq <= 0 || ng <= 0 || dim <= 0 || (dim & 7) || dim > 32768) return set_error(h, GRL_EINVAL, "grl_dist_topk: need nq,ng > 0, dim %% 8 == 0, dim <= 32768 (dim=%d)", dim);
    if (k <= 0 || k > TOPK_MAXK / 2) return set_error(h, GRL_EINVAL, "grl_dist_topk: need 0 < k <= %d", TOPK_MAXK / 2);
    if (metric != GRL_METRIC_NEG_DOT && metric != GRL_METRIC_L2) return set_error(h, GRL_EINVAL, "grl_dist_topk: unknown metric %d", metric);
    if (idx_base < 0 || idx_base + ng > 0xFFFFFFFFll) return set_error(h, GRL_EINVAL, "grl_dist_topk: global index must fit 32 bits");
    if (workspace_bytes < grl_dist_topk_workspace_bytes(nq, ng, dim)) return set_error(h, GRL_ENOMEM, "grl_dist_topk: workspace too small");
    DistTopkLayout L;
    dist_topk_layout(nq, ng, dim, k, &L);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* w = (uint8_t*)workspace;
    float* cd = (float*)(w + L.cd);
    int64_t* ci = (int64_t*)(w + L.ci);
    float* ed = (float*)(w + L.ed);
    float* gmax2 = (float*)(w + L.misc);
    int32_t* nflag = (int32_t*)(w + L.misc) + 1;
    int32_t* nrows = (int32_t*)(w + L.misc) + 2;
    int32_t* flags = (int32_t*)(w + L.flags);
    int32_t* rows = (int32_t*)(w + L.rows);
    int32_t* dirty = (int32_t*)(w + L.dirty);
    GRL_CUDA(h, cudaMemsetAsync(w + L.misc, 0, 64, st));
    GRL_TRY(coarse_topk_impl(h, metric, q, g, prepared, nq, ng, dim, L.kp, idx_base, cd, ci, gmax2, dirty, w + L.coarse, L.total - L.coarse, stream));
    GRL_TRY(grl_rescore(h, metric, q, g, nq, ng, dim, idx_base, ci, L.kp, ed, stream));
    GRL_TRY(grl_topk_finalize(h, metric, q, nq, dim, cd, ci, ed, L.kp, gmax2, dirty, k, top_d, top_i, flags, nflag, stream));
    // rows whose candidate list could not be proven complete: brute force (needs their count on the host: one 4-byte read)
    int host_nflag = 0;
    GRL_CUDA(h, cudaMemcpyAsync(&host_nflag, nflag, 4, cudaMemcpyDeviceToHost, st));
    GRL_CUDA(h, cudaStreamSynchronize(st));
    if (host_nflag > 0) {
        flagged_rows_kernel<<<(nq + 255) / 256, 256, 0, st>>>(flags, nq, rows, nrows);
        GRL_LAUNCH_CHECK(h);
        GRL_TRY(exact_topk_rows(h, metric, q, rows, host_nflag, g, ng, dim, k, idx_base, top_d, top_i, (float*)(w + L.tmp_d),
                                (int64_t*)(w + L.tmp_i), (float*)(w + L.brute), st));
    }
    return GRL_OK;
}

extern "C" int grl_dist_topk(grl_handle* h, int metric, const float* q, const float* g, int nq, int ng, int dim, int k,
                             int64_t idx_base, float* top_d, int64_t* top_i, void* workspace, size_t workspace_bytes, void* stream) {
    return dist_topk_impl(h, metric, q, g, nullptr, nq, ng, dim, k, idx_base, top_d, top_i, workspace, workspace_bytes, stream);
}
extern "C" int grl_dist_topk_prepared(grl_handle* h, int metric, const float* q, const float* g, const void* prepared, int nq, int ng, int dim,
                                      int k, int64_t idx_base, float* top_d, int64_t* top_i, void* workspace, size_t workspace_bytes,
                                      void* stream) {
    if (!prepared) return set_error(h, GRL_EINVAL, "grl_dist_topk_prepared: NULL argument");
    return dist_topk_impl(h, metric, q, g, prepared, nq, ng, dim, k, idx_base, top_d, top_i, workspace, workspace_bytes, stream);
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
