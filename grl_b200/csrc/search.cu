// grl_b200 — gallery-sharded exact top-k retrieval (sm_100a + NCCL): BASELINE.json configs[4].
//
// The reference has no counterpart beyond `-qf @ gf.T` + a full argsort on the CPU (reid/evaluator/attevaluator.py:44-46,
// reid/evaluator/eva_functions.py:139); what this file promises is the exact stable top-k of fixed-order fp32 distances,
// independent of chunking and of the number of gallery shards.
//
// Two-stage exact search.  Stage 1 ranks every (query, gallery row) pair by a COARSE distance: operands rounded to fp16 after a
// per-row power-of-two scaling, one tcgen05 MMA per k-step instead of three.  For any pair the coarse and the exact inner
// product differ by at most
//        E(q, g) = CE * |q| * |g|,     CE = 2^-10 + 2^-17 + 2 * dim * 2^-24
// (two fp16 roundings of <= 2^-11 relative each, their product term, subnormal/flush slack, and fp32 accumulation of `dim`
// terms on either side), so the K' coarse-nearest rows of a query contain its k exact-nearest whenever
//        coarse[K'-th] - E_max  >  exact[k-th among the K' re-scored]                                       (*)
// Stage 2 re-scores the candidates with a fixed-order fp32 inner product and checks (*) per query; queries that fail it
// (near-duplicate galleries, overflowed candidate buffers) are searched by brute force in the same fixed-order arithmetic.
// Only candidates that can still reach the top k are re-scored: a candidate whose coarse distance exceeds the coarse k-th
// by more than 2 E_max has at least k rows strictly closer in exact arithmetic and is skipped (it sorts last).
//
// Lists are packed 8-byte keys (orderable(distance) << 32 | global row index), ascending, KEY_EMPTY for unused slots: one
// array to sort, merge and send.
//
// Sharded protocol (grl_sharded_topk; W ranks, gallery rows split contiguously, rank r owns query slice r for the list work):
//   S0  every rank contributes nq/W query rows, ncclAllGather -> all nq rows          (or takes the full block as given)
//   S1  queries -> fp16 plane + scales + fixed-order |q|^2
//   S2  coarse pass over the local shard: K' coarse-nearest keys per query
//   S3  ncclAllReduce(max) of {overflow marks, max |g|^2}; all-to-all (grouped ncclSend/ncclRecv) of the lists by query slice;
//       rank r merges the W lists of its slice -> the global coarse K'; ncclAllGather of the merged slices
//   S4  every rank re-scores the candidates whose gallery rows it owns (0 elsewhere)
//   S5  ncclReduceScatter(sum) of the disjoint contributions -> exact distances of the own slice
//   S6  finalize the own slice: sort by (exact distance, index), top k, completeness proof -> flags
//   S7  ncclAllGather of the [nq/W, k+1] result keys (flag in the extra column); flagged rows compacted on the device;
//       keys -> top_d / top_i on every rank
//   S8  flagged rows: brute force per shard + all-gather + merge, scattered over the results (synchronous mode reads the
//       count on the host -- the one synchronisation; asynchronous mode gates a fixed number of row slots on the device)
#include <algorithm>

#include "api.h"
#include "comm.h"
#include "eval_common.cuh"

namespace grl {

// ------------------------------------------------------------------ operand conversion
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

// ------------------------------------------------------------------ streaming top-K' lists (packed keys)

__global__ void fill_keys_kernel(uint64_t* keys, long long n) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) keys[i] = KEY_EMPTY;
}

// Before the first chunk no threshold exists: mark every row "overflowed" (cnt = cap + 1, threshold -inf) so the first
// update rescans its tile row and the epilogue of the first GEMM appends nothing.
__global__ void topk_filter_init_kernel(float* thresh, int* cnt, int nq, int cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) { thresh[i] = -CUDART_INF_F; cnt[i] = cap + 1; }
}

__device__ __forceinline__ float thresh_of(uint64_t kth) { return kth == KEY_EMPTY ? CUDART_INF_F : key_value(kth); }

// Rows whose candidate count exceeds what one warp sorts in registers (possible only when the buffer is larger than
// LIST_WARP_MAX, i.e. K' >= 512): one block per row sorts the candidates alone in shared memory, then merges them into the
// sorted list by rank.  Defined after the warp kernel's constants.
__global__ void __launch_bounds__(TOPK_THREADS) list_update_kernel(int k, int min_cnt, uint64_t* __restrict__ list, float* __restrict__ thresh_out,
                                                                   const unsigned long long* __restrict__ cand, int* __restrict__ cand_cnt, int cap) {
    __shared__ uint64_t keys[TOPK_BUF];
    const int row = blockIdx.x;
    const int cnt = cand_cnt[row];
    if (cnt <= min_cnt || cnt > cap) return;          // the warp kernel's share / overflow (marked dirty there)
    uint64_t* lrow = list + (long long)row * k;
    uint64_t* cs = keys;                              // [npc] sorted candidates (cap <= TOPK_BUF / 2)
    uint64_t* ls = keys + TOPK_BUF / 2;               // [k]   the running list (k <= TOPK_MAXK <= TOPK_BUF / 2)
    int npc = 2;
    while (npc < cnt) npc <<= 1;
    for (int i = threadIdx.x; i < npc; i += blockDim.x) cs[i] = i < cnt ? cand[(long long)row * cap + i] : KEY_EMPTY;
    for (int i = threadIdx.x; i < k; i += blockDim.x) ls[i] = lrow[i];
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
            lrow[pos] = key;
            if (pos == k - 1) thresh_out[row] = thresh_of(key);
        }
    }
    if (threadIdx.x == 0) cand_cnt[row] = 0;
}

// ---- one WARP per row: the candidates are sorted in REGISTERS (NK keys per lane, element e = lane * NK + r; compare-exchange
// partners at distance < NK live in the same lane, the others one shuffle away), written to shared memory and merged into the
// sorted running list by rank -- every list key moves down by the number of candidates below it, every candidate lands at its
// own rank plus the number of list keys below it (keys are unique: the index is part of the key).  No block barrier anywhere.
template <int NK>
__device__ __forceinline__ void warp_sort_keys(uint64_t (&v)[NK], int lane) {
    constexpr int N = 32 * NK;
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride < NK) {
#pragma unroll
                for (int r = 0; r < NK; ++r) {
                    if ((r & stride) == 0) {
                        const int e = lane * NK + r;
                        const bool up = (size == N) || ((e & size) == 0);
                        const uint64_t a = v[r], b = v[r | stride];
                        const bool sw = (a > b) == up;
                        v[r] = sw ? b : a;
                        v[r | stride] = sw ? a : b;
                    }
                }
            } else {
                const int lstride = stride / NK;
                const bool lower = (lane & lstride) == 0;
#pragma unroll
                for (int r = 0; r < NK; ++r) {
                    const int e = lane * NK + r;
                    const bool up = (size == N) || ((e & size) == 0);
                    const uint64_t o = __shfl_xor_sync(0xffffffffu, v[r], lstride);
                    v[r] = ((v[r] < o) == (up == lower)) ? v[r] : o;
                }
            }
        }
    }
}

// Sorts this row's `cnt` candidates (from the candidate buffer, or -- first chunk -- straight from the stored tile row) and
// merges them into the list.  sl: the warp's copy of the list [k], cs: the warp's sorted candidates [32 * NK].
template <int NK>
__device__ __forceinline__ void warp_merge_row(int cnt, const unsigned long long* __restrict__ crow, const float* __restrict__ drow,
                                               int64_t idx_base, uint64_t* __restrict__ lrow, int k, uint64_t* sl, uint64_t* cs,
                                               float* __restrict__ thresh_row, int lane) {
    uint64_t v[NK];
#pragma unroll
    for (int r = 0; r < NK; ++r) {
        const int e = lane * NK + r;
        uint64_t key = KEY_EMPTY;
        if (e < cnt) key = drow ? make_key(drow[e], (uint32_t)(idx_base + e)) : crow[e];
        v[r] = key;
    }
    warp_sort_keys<NK>(v, lane);
#pragma unroll
    for (int r = 0; r < NK; ++r) cs[lane * NK + r] = v[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NK; ++r) {                    // candidates: rank among the list keys
        const int e = lane * NK + r;
        if (e < cnt) {
            int lo = 0, hi = k;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (sl[mid] < v[r]) lo = mid + 1; else hi = mid; }
            const int pos = lo + e;
            if (pos < k) {
                lrow[pos] = v[r];
                if (pos == k - 1) *thresh_row = key_value(v[r]);
            }
        }
    }
    for (int t = lane; t < k; t += 32) {              // list keys: shifted down by the number of candidates below them
        const uint64_t key = sl[t];
        int lo = 0, hi = cnt;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (cs[mid] < key) lo = mid + 1; else hi = mid; }
        const int pos = t + lo;
        if (lo > 0 && pos < k) lrow[pos] = key;
        if (pos == k - 1) *thresh_row = thresh_of(key);
    }
}

// ---- the first chunk of a search: no threshold exists yet, so its distance tile IS stored and every row selects its k smallest
// entries from it.  One block per row; the row (<= BOOT_MAX columns) lives in registers as orderable 32-bit distances, value
// j = 4 i + e of thread t being column 4 * BOOT_THREADS * i + 4 t + e (coalesced 16-byte loads).  O(ncols) work per row instead
// of a full sort, which is what lets the first chunk be thousands of columns wide and spares the search the first four
// "doubling" list merges.  Two selection paths, both exact (the k smallest (value, column) keys, whatever the data):
//  * exact path (boot_select_exact): histogram the values over BOOT_BINS ordered bins, find the bin that holds the k-th
//    smallest; everything in a lower bin is selected, the entries of that boundary bin are ranked and the smallest still needed
//    are taken; a boundary bin too large for that (degenerate data) is refined by another round on its own range.
//  * sampled pre-filter (boot_select_sampled, full chunks only): the same bin search over a 1-in-8 SAMPLE of the row gives a cut
//    t0 that about 2 k of the row's values pass; those are compacted into shared memory (one compare per value instead of a
//    bin computation and an atomic) and the exact path's bin search then runs over the few hundred survivors.  The cut is only
//    a work-saving guess: whenever fewer than k or more than BOOT_BUF values pass it (a few rows in 10^5 for continuous data;
//    every row of a gallery of duplicates) the row takes the exact path from its registers.
constexpr int BOOT_THREADS = 256, BOOT_PER = 32, BOOT_MAX = BOOT_THREADS * BOOT_PER, BOOT_BINS = 2048;
constexpr int BOOT_EDGE = 1024;        // boundary-bin entries resolved by ranking/sorting (more: another refinement round)
constexpr int BOOT_BUF = 2048;         // values the sampled pre-filter may pass on
constexpr int BOOT_SAMPLE = 8;         // the pre-filter looks at one value in BOOT_SAMPLE

struct BootSmem {
    unsigned int hist[BOOT_BINS];
    uint64_t sel[TOPK_MAXK];
    uint64_t edge[BOOT_EDGE];
    uint64_t buf[BOOT_BUF];
    int warp_tmp[BOOT_THREADS / 32];
    unsigned int lo, hi, bin, before, t0;
    int n_sel, n_edge, n_buf;
};

__device__ __forceinline__ int block_excl_scan_256(int v, int* smem_warp, int& total) {
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
    if (lane == 31) smem_warp[warp] = x;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < BOOT_THREADS / 32; ++w) { const int s = smem_warp[w]; if (w < warp) base += s; tot += s; }
    __syncthreads();
    total = tot;
    return base + x - v;
}

// A monotone map of [lo, hi] onto the bins 0 .. BOOT_BINS-1 (float arithmetic: the bins need not be equal, only ordered, with
// lo and hi in different ones -- an integer division per value would cost more than the whole selection); values below lo
// fall into bin 0.
struct BootBins {
    uint32_t lo;
    float scale;
    __device__ __forceinline__ BootBins(uint32_t lo_, uint32_t hi_) : lo(lo_), scale(hi_ > lo_ ? (float)(BOOT_BINS - 1) / (float)(hi_ - lo_) : 0.f) {}
    __device__ __forceinline__ unsigned int operator()(uint32_t x) const {
        return x <= lo ? 0u : min((unsigned int)(BOOT_BINS - 1), (unsigned int)((float)(x - lo) * scale));
    }
};

// With the histogram complete and visible: S.bin = the bin that holds the need-th smallest entry (1 <= need <= entries),
// S.before = the number of entries in lower bins.  Ends with a barrier.
__device__ __forceinline__ void boot_find_bin(BootSmem& S, int need) {
    constexpr int PER = BOOT_BINS / BOOT_THREADS;             // each thread owns PER consecutive bins
    static_assert(PER == 8, "boot_find_bin reads two uint4 per thread");
    const uint4 h0 = reinterpret_cast<const uint4*>(S.hist)[2 * threadIdx.x], h1 = reinterpret_cast<const uint4*>(S.hist)[2 * threadIdx.x + 1];
    const int c[PER] = {(int)h0.x, (int)h0.y, (int)h0.z, (int)h0.w, (int)h1.x, (int)h1.y, (int)h1.z, (int)h1.w};
    int mine = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) mine += c[j];
    int total;
    const int before = block_excl_scan_256(mine, S.warp_tmp, total);
    if (before < need && need <= before + mine) {
        int run = before;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (run < need && need <= run + c[j]) { S.bin = threadIdx.x * PER + j; S.before = run; }
            run += c[j];
        }
    }
    __syncthreads();
}

// The `need` smallest of the distinct keys edge[0 .. ne) -> sel[base ..] (any order).  Callers put a barrier on both sides.
__device__ __forceinline__ void boot_take_edge(BootSmem& S, int ne, int base, int need) {
    if (ne <= 64) {                                   // the normal case, a handful of keys: rank each against the others
        if ((int)threadIdx.x < ne) {
            const uint64_t key = S.edge[threadIdx.x];
            int rank = 0;
            for (int l = 0; l < ne; ++l) rank += S.edge[l] < key ? 1 : 0;
            if (rank < need) S.sel[base + rank] = key;
        }
        return;
    }
    int npe = 128;
    while (npe < ne) npe <<= 1;
    for (int i = ne + threadIdx.x; i < npe; i += BOOT_THREADS) S.edge[i] = KEY_EMPTY;
    block_bitonic_sort(S.edge, npe);
    for (int i = threadIdx.x; i < need; i += BOOT_THREADS) S.sel[base + i] = S.edge[i];
}

#define BOOT_COL(j) (((j) >> 2) * (4 * BOOT_THREADS) + 4 * (int)threadIdx.x + ((j) & 3))
#define BOOT_KEY(j) (((uint64_t)v[j] << 32) | (uint32_t)((uint32_t)idx_base + (uint32_t)BOOT_COL(j)))

// The k smallest keys of the row in registers -> S.sel[0 .. k) (any order); k < ncols.
__device__ __forceinline__ void boot_select_exact(BootSmem& S, const uint32_t (&v)[BOOT_PER], int ncols, int k, int64_t idx_base) {
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
#pragma unroll
    for (int j = 0; j < BOOT_PER; ++j)
        if (BOOT_COL(j) < ncols) { lo = min(lo, v[j]); hi = max(hi, v[j]); }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off)); }
    __syncthreads();                                  // (a row sent here by the pre-filter: every thread is done with S)
    if (threadIdx.x == 0) { S.lo = 0xFFFFFFFFu; S.hi = 0u; S.n_sel = 0; S.n_edge = 0; }
    __syncthreads();
    if (lane_id() == 0) { atomicMin(&S.lo, lo); atomicMax(&S.hi, hi); }
    __syncthreads();
    lo = S.lo; hi = S.hi;
    int need = k;                                     // entries still to take from [lo, hi]
    // Rounds of: histogram of the values still in range -> the bin holding the `need`-th of them.  Everything in a lower bin is
    // selected; a boundary bin of at most BOOT_EDGE entries (the normal case after ONE round: a handful) is resolved by
    // ranking its keys, otherwise the range narrows to that bin and the round repeats (the range strictly shrinks: lo and hi
    // always land in different bins).
    bool resolve;
    while (true) {
        for (int i = threadIdx.x; i < BOOT_BINS; i += BOOT_THREADS) S.hist[i] = 0;
        __syncthreads();
        const BootBins bins(lo, hi);
#pragma unroll
        for (int j = 0; j < BOOT_PER; ++j)
            if (BOOT_COL(j) < ncols && v[j] >= lo && v[j] <= hi) atomicAdd(&S.hist[bins(v[j])], 1u);
        __syncthreads();
        boot_find_bin(S, need);
        const unsigned int bin = S.bin;
        resolve = (int)S.hist[bin] <= BOOT_EDGE || lo == hi;         // (lo == hi: all remaining values are equal -> ties by column)
        uint32_t nlo = 0xFFFFFFFFu, nhi = 0u;
#pragma unroll
        for (int j = 0; j < BOOT_PER; ++j) {
            if (BOOT_COL(j) < ncols && v[j] >= lo && v[j] <= hi) {
                const unsigned int bj = bins(v[j]);
                if (bj < bin) S.sel[atomicAdd(&S.n_sel, 1)] = BOOT_KEY(j);
                else if (bj == bin) {
                    if (resolve) { const int e = atomicAdd(&S.n_edge, 1); if (e < BOOT_EDGE) S.edge[e] = BOOT_KEY(j); }
                    else { nlo = min(nlo, v[j]); nhi = max(nhi, v[j]); }
                }
            }
        }
        need -= (int)S.before;
        if (resolve) break;
        if (threadIdx.x == 0) { S.lo = 0xFFFFFFFFu; S.hi = 0u; }
        __syncthreads();
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) { nlo = min(nlo, __shfl_xor_sync(0xffffffffu, nlo, off)); nhi = max(nhi, __shfl_xor_sync(0xffffffffu, nhi, off)); }
        if (lane_id() == 0 && nlo <= nhi) { atomicMin(&S.lo, nlo); atomicMax(&S.hi, nhi); }
        __syncthreads();
        lo = S.lo; hi = S.hi;
        __syncthreads();
    }
    __syncthreads();
    const int base = S.n_sel;
    if (S.n_edge <= BOOT_EDGE) { boot_take_edge(S, S.n_edge, base, need); return; }
    // more EQUAL values than the edge buffer holds: the lowest columns among them, by ordered scans (column order = i, thread, e)
    int taken = 0;
    for (int i = 0; i < BOOT_PER / 4 && taken < need; ++i) {
        int n_eq = 0;
#pragma unroll
        for (int j = 0; j < BOOT_PER; ++j)
            if ((j >> 2) == i && BOOT_COL(j) < ncols && v[j] == lo) ++n_eq;
        int tot;
        int off = taken + block_excl_scan_256(n_eq, S.warp_tmp, tot);
#pragma unroll
        for (int j = 0; j < BOOT_PER; ++j)
            if ((j >> 2) == i && BOOT_COL(j) < ncols && v[j] == lo) { if (off < need) S.sel[base + off] = BOOT_KEY(j); ++off; }
        taken += tot;
    }
}

// The sampled pre-filter (see above).  Returns false, block-uniformly and with S.sel untouched, when the row must take the
// exact path; true with the k smallest keys in S.sel[0 .. k).  ncols == BOOT_MAX.
__device__ __forceinline__ bool boot_select_sampled(BootSmem& S, const uint32_t (&v)[BOOT_PER], int k, int64_t idx_base) {
    // the sample: one component of every other 16-byte load (columns spread over the whole chunk)
    const uint32_t sv[4] = {v[0], v[9], v[18], v[27]};
    static_assert(BOOT_PER / 4 == BOOT_SAMPLE, "the sample below is 4 of a thread's 32 values");
    uint32_t lo = min(min(sv[0], sv[1]), min(sv[2], sv[3])), hi = max(max(sv[0], sv[1]), max(sv[2], sv[3]));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off)); }
    if (threadIdx.x == 0) { S.lo = 0xFFFFFFFFu; S.hi = 0u; S.t0 = 0u; S.n_sel = 0; S.n_edge = 0; S.n_buf = 0; }
    for (int i = threadIdx.x; i < BOOT_BINS; i += BOOT_THREADS) S.hist[i] = 0;
    __syncthreads();
    if (lane_id() == 0) { atomicMin(&S.lo, lo); atomicMax(&S.hi, hi); }
    __syncthreads();
    lo = S.lo; hi = S.hi;
    if (lo == hi) return false;
    // sample rank whose value about k + 4.5 sigma of the row's values pass (sigma^2 ~ BOOT_SAMPLE * count)
    const int want = k + (int)(4.5f * sqrtf((float)(k * BOOT_SAMPLE))) + 2 * BOOT_SAMPLE;
    const int rank = (want + BOOT_SAMPLE - 1) / BOOT_SAMPLE;            // <= 181 of the 1024 samples for k <= TOPK_MAXK
    unsigned int sb[4];
    {
        const BootBins bins(lo, hi);
#pragma unroll
        for (int j = 0; j < 4; ++j) { sb[j] = bins(sv[j]); atomicAdd(&S.hist[sb[j]], 1u); }
    }
    __syncthreads();
    boot_find_bin(S, rank);
    {
        const unsigned int bin = S.bin;
        uint32_t t = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) if (sb[j] <= bin) t = max(t, sv[j]);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) t = max(t, __shfl_xor_sync(0xffffffffu, t, off));
        if (lane_id() == 0) atomicMax(&S.t0, t);
    }
    __syncthreads();
    const uint32_t t0 = S.t0;
    // compaction of the values <= t0 (any order: warp-aggregated reservation, then per-thread writes)
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < BOOT_PER; ++j) cnt += v[j] <= t0 ? 1 : 0;
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, off); if (lane_id() >= off) incl += y; }
    int wbase = 0;
    if (lane_id() == 31) wbase = atomicAdd(&S.n_buf, incl);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    int pos = wbase + incl - cnt;
#pragma unroll
    for (int j = 0; j < BOOT_PER; ++j)
        if (v[j] <= t0) { if (pos < BOOT_BUF) S.buf[pos] = BOOT_KEY(j); ++pos; }
    for (int i = threadIdx.x; i < BOOT_BINS; i += BOOT_THREADS) S.hist[i] = 0;      // (the sample's bins were last read before the barrier above)
    __syncthreads();
    const int n = S.n_buf;
    if (n < k || n > BOOT_BUF) return false;
    // the exact bin search over the survivors
    const BootBins bins(lo, t0);
    for (int i = threadIdx.x; i < n; i += BOOT_THREADS) atomicAdd(&S.hist[bins((uint32_t)(S.buf[i] >> 32))], 1u);
    __syncthreads();
    boot_find_bin(S, k);
    const unsigned int bin = S.bin;
    const int before = (int)S.before, in_bin = (int)S.hist[bin];
    if (in_bin > BOOT_EDGE) return false;
    for (int i = threadIdx.x; i < n; i += BOOT_THREADS) {
        const uint64_t key = S.buf[i];
        const unsigned int bj = bins((uint32_t)(key >> 32));
        if (bj < bin) S.sel[atomicAdd(&S.n_sel, 1)] = key;
        else if (bj == bin) S.edge[atomicAdd(&S.n_edge, 1)] = key;
    }
    __syncthreads();
    boot_take_edge(S, in_bin, before, k - before);
    return true;
}

// (four resident blocks per SM = 64 registers per thread: the row load of one block hides behind the selection of the others)
__global__ void __launch_bounds__(BOOT_THREADS, 4) list_boot_select_kernel(const float* __restrict__ tile, long long ld, int ncols, int k,
                                                                        int64_t idx_base, uint64_t* __restrict__ list, float* __restrict__ thresh_out,
                                                                        int* __restrict__ cand_cnt) {
    __shared__ BootSmem S;
    const int row = blockIdx.x;
    const float* drow = tile + (long long)row * ld;
    uint32_t v[BOOT_PER];                             // columns past ncols: the largest orderable value
#pragma unroll
    for (int i = 0; i < BOOT_PER / 4; ++i) {
        const int c = BOOT_COL(4 * i);
        const float big = __uint_as_float(0x7FFFFFFFu);
        float4 f = make_float4(big, big, big, big);
        if (c + 3 < ncols) f = __ldg(reinterpret_cast<const float4*>(drow + c));
        else { if (c < ncols) f.x = drow[c]; if (c + 1 < ncols) f.y = drow[c + 1]; if (c + 2 < ncols) f.z = drow[c + 2]; }
        v[4 * i] = orderable(f.x); v[4 * i + 1] = orderable(f.y); v[4 * i + 2] = orderable(f.z); v[4 * i + 3] = orderable(f.w);
    }
    uint64_t* lrow = list + (long long)row * k;
    if (ncols <= k) {                                 // fewer columns than list slots: everything is selected
        for (int i = threadIdx.x; i < k; i += BOOT_THREADS) S.sel[i] = KEY_EMPTY;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < BOOT_PER; ++j)
            if (BOOT_COL(j) < ncols) S.sel[BOOT_COL(j)] = BOOT_KEY(j);
    } else {
        bool done = false;
        if (ncols == BOOT_MAX) done = boot_select_sampled(S, v, k, idx_base);
        if (!done) boot_select_exact(S, v, ncols, k, idx_base);
    }
    __syncthreads();
    if (k == 256) {                                   // the common list length: one warp sorts the selection in registers, no barriers
        if (threadIdx.x < 32) {
            uint64_t r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = S.sel[threadIdx.x * 8 + j];
            warp_sort_keys<8>(r, threadIdx.x);
#pragma unroll
            for (int j = 0; j < 8; ++j) lrow[threadIdx.x * 8 + j] = r[j];
            if (threadIdx.x == 31) { thresh_out[row] = thresh_of(r[7]); cand_cnt[row] = 0; }
        }
        return;
    }
    int npad = 2;
    while (npad < k) npad <<= 1;
    for (int i = k + threadIdx.x; i < npad; i += BOOT_THREADS) S.sel[i] = KEY_EMPTY;      // (k <= TOPK_MAXK, a power of two in practice)
    block_bitonic_sort(S.sel, npad);
    for (int i = threadIdx.x; i < k; i += BOOT_THREADS) lrow[i] = S.sel[i];
    if (threadIdx.x == 0) { thresh_out[row] = thresh_of(S.sel[k - 1]); cand_cnt[row] = 0; }
}
#undef BOOT_COL
#undef BOOT_KEY

constexpr int LIST_WARPS = 4;          // rows per block of list_update_warp_kernel
constexpr int LIST_WARP_MAX = 512;     // candidates one warp sorts in registers (16 keys per lane)

// `dist` != NULL: first chunk -- every row takes all `ncols` (<= LIST_WARP_MAX) columns of its stored tile row.  Otherwise rows
// with 1..LIST_WARP_MAX buffered candidates are merged here; more than `cap`: the buffer overflowed and the row's list can no
// longer be trusted -> dirty (finalisation sends it to brute force); in between (cap > LIST_WARP_MAX): list_update_kernel.
__global__ void __launch_bounds__(LIST_WARPS * 32) list_update_warp_kernel(int nq, int k, const float* __restrict__ dist, long long ld, int ncols,
                                                                           int64_t idx_base, uint64_t* __restrict__ list, float* __restrict__ thresh_out,
                                                                           const unsigned long long* __restrict__ cand, int* __restrict__ cand_cnt,
                                                                           int cap, uint32_t* __restrict__ dirty) {
    extern __shared__ uint64_t warp_keys[];           // [LIST_WARPS][k + LIST_WARP_MAX]
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int row = blockIdx.x * LIST_WARPS + warp;
    if (row >= nq) return;
    int cnt = dist ? ncols : cand_cnt[row];
    if (cnt == 0) return;
    if (!dist && cnt > cap) {
        if (lane == 0) { dirty[row] = 1u; thresh_out[row] = -CUDART_INF_F; cand_cnt[row] = 0; }
        return;
    }
    if (cnt > LIST_WARP_MAX) return;                  // list_update_kernel's share
    uint64_t* sl = warp_keys + (size_t)warp * (k + LIST_WARP_MAX);
    uint64_t* cs = sl + k;
    uint64_t* lrow = list + (long long)row * k;
    for (int t = lane; t < k; t += 32) sl[t] = lrow[t];
    __syncwarp();
    const unsigned long long* crow = cand + (long long)row * cap;
    const float* drow = dist ? dist + (long long)row * ld : nullptr;
    float* th = thresh_out + row;
    if (cnt <= 32) warp_merge_row<1>(cnt, crow, drow, idx_base, lrow, k, sl, cs, th, lane);
    else if (cnt <= 128) warp_merge_row<4>(cnt, crow, drow, idx_base, lrow, k, sl, cs, th, lane);
    else if (cnt <= 256) warp_merge_row<8>(cnt, crow, drow, idx_base, lrow, k, sl, cs, th, lane);
    else warp_merge_row<16>(cnt, crow, drow, idx_base, lrow, k, sl, cs, th, lane);
    if (lane == 0) cand_cnt[row] = 0;
}

// keys [nq][k] -> (distance f32, index i64) [nq][k]; empty slots become (+inf, -1)
__global__ void unpack_keys_kernel(const uint64_t* __restrict__ keys, long long ld_keys, int nq, int k, float* __restrict__ out_d,
                                   int64_t* __restrict__ out_i) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)nq * k) return;
    const long long r = i / k;
    const int t = (int)(i - r * k);
    const uint64_t key = keys[r * ld_keys + t];
    if (key == KEY_EMPTY) { out_d[i] = CUDART_INF_F; out_i[i] = -1; }
    else { out_d[i] = key_value(key); out_i[i] = (int64_t)key_index(key); }
}

// ------------------------------------------------------------------ exact re-score
// Staged API: exact_d[row][t] = exact distance of query `row` to candidate cand_i[row][t] when that gallery row lives in this
// shard ([idx_base, idx_base + ng)), else 0 -- so that the per-shard results combine by a plain sum.  One block per query.
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

// E_max of condition (*) for one query: CE * |q| * max|g| (slightly inflated for the rounding of this very product)
__device__ __forceinline__ float coarse_emax(float ce, float qq, float g2) { return ce * sqrtf(qq) * sqrtf(g2) * 1.00001f; }

// The fused path: candidates come as the (merged) coarse key list of the row.  Candidate t >= k is SKIPPED (exact = +inf,
// it sorts behind every re-scored one) when its coarse distance exceeds the coarse k-th by more than twice the worst-case
// coarse error: each of the first k candidates is then strictly closer in exact arithmetic, so it cannot be among the k nearest.
// Rows owned by another shard get 0 (sum-combinable); stats[2] / stats[3] count re-scored / skipped candidates of this rank.
__global__ void __launch_bounds__(256) rescore_keys_kernel(int metric, const float* __restrict__ q, const float* __restrict__ q_n2,
                                                           const float* __restrict__ g, int ng, int dim, long long idx_base,
                                                           const uint64_t* __restrict__ list, int kp, int k, const uint32_t* __restrict__ gmax2_bits,
                                                           float ce, float* __restrict__ exact_d, int32_t* __restrict__ stats) {
    extern __shared__ float qs[];
    __shared__ int n_done, n_skip;
    const int row = blockIdx.x;
    if (threadIdx.x == 0) { n_done = 0; n_skip = 0; }
    for (int j = threadIdx.x; j < (dim >> 2); j += blockDim.x)
        reinterpret_cast<float4*>(qs)[j] = __ldg(reinterpret_cast<const float4*>(q + (long long)row * dim) + j);
    __syncthreads();
    const float qq = q_n2[row];                       // fixed-order |q|^2 (f16_rows_kernel), the value warp_sqnorm_fixed returns
    const float g2 = __uint_as_float(*gmax2_bits);
    const float emax = coarse_emax(ce, qq, g2);
    const uint64_t* lrow = list + (long long)row * kp;
    const uint64_t kth = (k - 1 < kp) ? lrow[k - 1] : KEY_EMPTY;
    float cut = CUDART_INF_F;                         // coarse values above `cut` cannot reach the top k
    if (kth != KEY_EMPTY) {
        const float ck = key_value(kth);
        cut = (metric == GRL_METRIC_L2) ? ck + 4.f * emax + 2e-5f * (qq + g2) + 4e-12f : ck + 2.f * emax + 2e-6f * fabsf(ck);
        if (!(cut == cut)) cut = CUDART_INF_F;        // NaN anywhere: re-score everything
    }
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int done = 0, skip = 0;
    for (int t = warp; t < kp; t += nw) {
        const uint64_t key = lrow[t];
        float d = 0.f;
        if (key != KEY_EMPTY) {
            const long long idx = (long long)key_index(key) - idx_base;
            if (idx >= 0 && idx < ng) {
                if (t >= k && key_value(key) > cut) { d = CUDART_INF_F; ++skip; }
                else {
                    float dot, gg;
                    warp_dot_fixed(qs, g + idx * dim, dim, dot, gg);
                    d = exact_distance(metric, dot, qq, gg);
                    ++done;
                }
            }
        }
        if (lane_id() == 0) exact_d[(long long)row * kp + t] = d;
    }
    if (stats) {
        if (lane_id() == 0) { atomicAdd(&n_done, done); atomicAdd(&n_skip, skip); }
        __syncthreads();
        if (threadIdx.x == 0) { atomicAdd(stats + 2, n_done); atomicAdd(stats + 3, n_skip); }
    }
}

// ------------------------------------------------------------------ finalisation + completeness proof
// Staged API.  coarse_d holds -q.g (metric 0) or the SQUARED L2 distance (metric 1) of the coarse pass, ascending.
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
        uint64_t key = KEY_EMPTY;
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
        if (key == KEY_EMPTY) { top_d[(long long)row * k + t] = CUDART_INF_F; top_i[(long long)row * k + t] = -1; }
        else { top_d[(long long)row * k + t] = key_value(key); top_i[(long long)row * k + t] = (int64_t)key_index(key); }
    }
    if (threadIdx.x == 0) {
        bool ok = !(dirty && dirty[row]);             // a shard's candidate list overflowed: not provable
        const int nvalid = nvalid_s;
        if (ok && nvalid == kp && nvalid > k - 1 && fkeys[k - 1] != KEY_EMPTY) {   // a full list may have cut off relevant rows
            const float u = key_value(fkeys[k - 1]);
            const float ck = coarse_d[(long long)row * kp + kp - 1];
            const float g2 = *gmax2;
            const float emax = coarse_emax(ce, qq, g2);
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

// The fused path: row r of this rank's query slice.  list = merged global coarse keys of the slice rows, exact_d their
// re-scored distances (+inf for skipped candidates); writes k result keys + the flag (0/1) as the (k+1)-th 8-byte word.
__global__ void __launch_bounds__(256) finalize_keys_kernel(int metric, const float* __restrict__ q_n2, const uint64_t* __restrict__ list,
                                                            const float* __restrict__ exact_d, int kp, int npad, const uint32_t* __restrict__ gmax2_bits,
                                                            float ce, int k, const uint32_t* __restrict__ dirty, int rows_valid,
                                                            uint64_t* __restrict__ out) {
    extern __shared__ uint64_t fkeys[];
    __shared__ int nvalid_s;
    const int row = blockIdx.x;
    uint64_t* orow = out + (long long)row * (k + 1);
    if (row >= rows_valid) {                          // padding rows of the last slice: empty result, never flagged
        for (int t = threadIdx.x; t < k; t += blockDim.x) orow[t] = KEY_EMPTY;
        if (threadIdx.x == 0) orow[k] = 0;
        return;
    }
    if (threadIdx.x == 0) nvalid_s = 0;
    __syncthreads();
    const uint64_t* lrow = list + (long long)row * kp;
    int local_valid = 0;
    for (int t = threadIdx.x; t < npad; t += blockDim.x) {
        uint64_t key = KEY_EMPTY;
        if (t < kp) {
            const uint64_t c = lrow[t];
            if (c != KEY_EMPTY) { key = make_key(exact_d[(long long)row * kp + t], key_index(c)); ++local_valid; }
        }
        fkeys[t] = key;
    }
    if (local_valid) atomicAdd(&nvalid_s, local_valid);
    block_bitonic_sort(fkeys, npad);
    for (int t = threadIdx.x; t < k; t += blockDim.x) orow[t] = fkeys[t];
    if (threadIdx.x == 0) {
        bool ok = !(dirty && dirty[row]);
        const int nvalid = nvalid_s;
        if (ok && nvalid == kp && nvalid > k - 1 && fkeys[k - 1] != KEY_EMPTY) {
            const float u = key_value(fkeys[k - 1]);
            const float ck = key_value(lrow[kp - 1]);
            const float qq = q_n2[row] * 1.00001f, g2 = __uint_as_float(*gmax2_bits);
            const float emax = coarse_emax(ce, qq, g2);
            if (metric == GRL_METRIC_L2) {
                const float lb = ck - 2.f * emax - 1e-5f * (qq + g2);
                ok = lb > 1e-12f && sqrtf(lb) * 0.999999f > u;
            } else {
                ok = ck - emax - 1e-6f * fabsf(ck) > u;
            }
            if (!(ok)) ok = false;
        }
        orow[k] = ok ? 0ull : 1ull;
    }
}

static int next_pow2_host(int n) { int p = 1; while (p < n) p <<= 1; return p; }

// lists [nlists][rows][kp] (ascending keys) -> out [rows][kp]: the kp smallest of each row's nlists * kp keys, ascending.
// A merge TREE over the sorted lists, not a sort of their union: for ascending A and B, C[i] = min(A[i], B[kp-1-i]) holds the kp
// smallest keys of both as a bitonic sequence, which log2(kp) half-cleaner stages sort -- (1 + log2 kp) passes over kp keys per
// pair and level instead of the log2^2 stages of a bitonic sort over all nlists * kp keys (8 lists of 256: 7.5x fewer compare-
// exchanges).  smem: npl * kp keys, npl = nlists rounded up to a power of two (missing lists are empty); kp is a power of two.
__global__ void __launch_bounds__(256) merge_key_lists_kernel(const uint64_t* __restrict__ lists, int nlists, int rows, int kp, int npl,
                                                              uint64_t* __restrict__ out) {
    extern __shared__ uint64_t mkeys[];               // [npl][kp]
    const int row = blockIdx.x;
    const int lkp = 31 - __clz(kp);
    for (int i = threadIdx.x; i < npl * kp; i += blockDim.x) {
        const int s = i >> lkp, j = i & (kp - 1);
        mkeys[i] = s < nlists ? lists[((long long)s * rows + row) * kp + j] : KEY_EMPTY;
    }
    __syncthreads();
    for (int gap = 1; gap < npl; gap <<= 1) {         // this level merges list a + gap into list a, a = 0, 2 gap, 4 gap, ...
        const int npairs = npl / (2 * gap);
        for (int i = threadIdx.x; i < npairs * kp; i += blockDim.x) {
            const int pr = i >> lkp, t = i & (kp - 1);
            uint64_t* A = mkeys + (size_t)pr * 2 * gap * kp;
            const uint64_t a = A[t], b = A[(size_t)gap * kp + (kp - 1 - t)];
            A[t] = a < b ? a : b;                      // (A[t] is read and written by this thread only; the other list is read-only here)
        }
        __syncthreads();
        for (int stride = kp >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < npairs * (kp >> 1); i += blockDim.x) {
                const int pr = i >> (lkp - 1), j = i & ((kp >> 1) - 1);
                uint64_t* A = mkeys + (size_t)pr * 2 * gap * kp;
                const int lo = 2 * j - (j & (stride - 1)), hi = lo + stride;
                const uint64_t a = A[lo], b = A[hi];
                if (a > b) { A[lo] = b; A[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < kp; i += blockDim.x) out[(long long)row * kp + i] = mkeys[i];
}

static int merge_key_lists(grl_handle* h, cudaStream_t st, const uint64_t* lists, int nlists, int rows, int kp, uint64_t* out) {
    const int npl = next_pow2_host(nlists);
    const size_t smem = (size_t)npl * kp * 8;
    GRL_TRY(ensure_dyn_smem(h, (const void*)merge_key_lists_kernel, (int)smem));
    merge_key_lists_kernel<<<rows, 256, smem, st>>>(lists, nlists, rows, kp, npl, out);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

// Flag column of the result keys -> ascending list of flagged rows + counters (ONE block: the order must be the same on every
// rank).  stats[0] = flagged queries, stats[1] = rows with an overflow mark.
__global__ void __launch_bounds__(1024) compact_flags_kernel(const uint64_t* __restrict__ out, int nq, int k, const uint32_t* __restrict__ dirty,
                                                             int32_t* __restrict__ rows, int32_t* __restrict__ nflag, int32_t* __restrict__ stats) {
    __shared__ int part[1024];
    __shared__ int dpart[32];
    const int per = (nq + blockDim.x - 1) / blockDim.x;
    const int r0 = threadIdx.x * per, r1 = min(nq, r0 + per);
    int c = 0, dcount = 0;
    for (int r = r0; r < r1; ++r) {
        c += out[(long long)r * (k + 1) + k] ? 1 : 0;
        dcount += (dirty && dirty[r]) ? 1 : 0;
    }
    part[threadIdx.x] = c;
    dcount = (int)warp_sum((float)dcount);
    if (lane_id() == 0) dpart[threadIdx.x >> 5] = dcount;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < (int)blockDim.x; ++i) { const int v = part[i]; part[i] = run; run += v; }
        int dsum = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) dsum += dpart[i];
        *nflag = run;
        if (stats) { stats[0] = run; stats[1] = dsum; }
    }
    __syncthreads();
    int pos = part[threadIdx.x];
    for (int r = r0; r < r1; ++r)
        if (out[(long long)r * (k + 1) + k]) rows[pos++] = r;
}

// ------------------------------------------------------------------ brute force in the fixed-order arithmetic
// tile[r][c] = exact distance of query rows[r0 + r] (or r0 + r when rows == NULL) to gallery row c.  Each block keeps rq query
// rows in shared memory and streams a slab of gallery rows once, one warp per gallery row.  nrows_dev (optional) gates the
// launch on a device-side row count (asynchronous mode: the host does not know how many rows are flagged).
__global__ void __launch_bounds__(256) exact_rows_kernel(int metric, const float* __restrict__ q, const int32_t* __restrict__ rows, int r0, int rq,
                                                         const int32_t* __restrict__ nrows_dev, const float* __restrict__ g, int ng, int dim,
                                                         float* __restrict__ tile, long long ld_tile) {
    extern __shared__ float qs[];                     // [rq][dim]
    __shared__ float qqs[16];
    if (nrows_dev) {
        const int n = *nrows_dev;
        if (r0 >= n) return;
        rq = min(rq, n - r0);
    }
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

// dst[rows[r0 + r]] = src[r] for r < min(n, *nrows_dev - r0)
__global__ void scatter_rows_kernel(const float* __restrict__ src_d, const int64_t* __restrict__ src_i, const int32_t* __restrict__ rows, int r0,
                                    const int32_t* __restrict__ nrows_dev, int k, float* __restrict__ dst_d, int64_t* __restrict__ dst_i) {
    const int r = blockIdx.x;
    if (nrows_dev && r0 + r >= *nrows_dev) return;
    const long long dst = (long long)rows[r0 + r] * k;
    for (int t = threadIdx.x; t < k; t += blockDim.x) { dst_d[dst + t] = src_d[(long long)r * k + t]; dst_i[dst + t] = src_i[(long long)r * k + t]; }
}

// dst[r] = src[rows[r]] (float4 granularity), r < nrows
__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ rows, int nrows, int dim, float* __restrict__ dst) {
    const int nv = dim >> 2;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)nrows * nv) return;
    const long long r = i / nv;
    const int j = (int)(i - r * nv);
    reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src + (long long)rows[r] * dim) + j);
}
// result keys of the second chance -> rows[r] of the outputs, for the rows that now carry a proof
__global__ void scatter_keys_kernel(const uint64_t* __restrict__ out2, const int32_t* __restrict__ rows, int k, float* __restrict__ top_d,
                                    int64_t* __restrict__ top_i) {
    const int r = blockIdx.x;
    const uint64_t* orow = out2 + (long long)r * (k + 1);
    if (orow[k]) return;                              // still flagged: the brute force writes this row
    const long long dst = (long long)rows[r] * k;
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const uint64_t key = orow[t];
        if (key == KEY_EMPTY) { top_d[dst + t] = CUDART_INF_F; top_i[dst + t] = -1; }
        else { top_d[dst + t] = key_value(key); top_i[dst + t] = (int64_t)key_index(key); }
    }
}
// rows2[i] (an index into `rows`) -> the original row id
__global__ void remap_rows_kernel(int32_t* __restrict__ rows2, const int32_t* __restrict__ n, const int32_t* __restrict__ rows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *n) rows2[i] = rows[rows2[i]];
}

}  // namespace grl

using namespace grl;

// ==================================================================================================== host side
static int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

// Column chunks of one coarse pass.  The first chunk has no thresholds yet: its tile IS stored and every row is rescanned, so it
// holds TOPK_FIRST_CHUNK columns (one block selects a row's K' smallest in O(columns), list_boot_select_kernel).  Afterwards the K'-th best of n_seen columns lets
// ~K' * nc / n_seen candidates per row through, so chunks grow with n_seen (at most doubling the columns seen) up to the
// steady-state size, whose 256 x 256 tiles fill whole waves of the persistent grid; their tiles are never stored.
constexpr int TOPK_FIRST_CHUNK = BOOT_MAX;          // 8192 columns: list_boot_select_kernel keeps a row in registers
// candidates per query row and column chunk: a chunk at most doubles the columns seen, so ~K' candidates per row are expected;
// the buffer holds twice that (an overflow outside the first chunk marks the row dirty -> brute force)
static int cand_cap(int kprime) { return std::max(512, 2 * kprime); }

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

static float coarse_error_constant(int dim) {   // CE of the header comment
    return 0x1p-10f + 0x1p-17f + 2.f * (float)dim * 0x1p-24f;
}

// ---- workspace of one coarse pass (nq query rows against a shard of ng rows)
struct CoarseLayout {
    int chunk, chunk_alloc, first, cap;
    size_t q16, g16, qf, gf, tile, thresh, cnt, cand, total;
};
static void coarse_layout(int nq, int ng, int dim, int kprime, bool prepared, CoarseLayout* L) {
    L->chunk = topk_chunk_max(nq, ng, 148);
    // the conversion buffers of an unprepared shard are sized for the largest chunk any query count can pick, so that a layout
    // computed for fewer rows (the second chance over the flagged queries) always fits the reservation made for all of them
    L->chunk_alloc = (int)std::min<long long>(16384, ((long long)ng + 7) / 8 * 8);
    L->first = (topk_next_chunk(0, ng, L->chunk) + 7) / 8 * 8;     // leading dimension of the stored first tile
    L->cap = cand_cap(kprime);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    L->q16 = take((size_t)nq * dim * 2);
    L->g16 = take(prepared ? 0 : (size_t)L->chunk_alloc * dim * 2);
    L->qf = take((size_t)nq * 4 * 2);                 // inv_scale | sqnorm
    L->gf = take(prepared ? 0 : (size_t)L->chunk_alloc * 4 * 2);
    L->tile = take((size_t)nq * L->first * 4);        // coarse distances of the first chunk only
    L->thresh = take((size_t)nq * 4);
    L->cnt = take((size_t)nq * 4);
    L->cand = take((size_t)nq * L->cap * 8);
    L->total = off;
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

// Queries -> fp16 plane + scales + fixed-order squared norms (stage S1).  `scratch_max` receives the largest query norm (unused).
static int convert_queries(grl_handle* h, cudaStream_t st, const float* q, int nq, int dim, uint8_t* w, const CoarseLayout& L) {
    float* q_inv = (float*)(w + L.qf);
    f16_rows_kernel<<<(int)(((long long)nq * 32 + 255) / 256), 256, 0, st>>>(q, nq, dim, (__half*)(w + L.q16), q_inv, q_inv + nq,
                                                                             (unsigned int*)(w + L.cand));   // cand is free until the first GEMM
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

// Stage S2: the K' coarse-nearest rows of the shard per query, as sorted keys list [nq][kprime].  Queries must have been
// converted (convert_queries).  gmax2_bits (uint-ordered float) is max-combined with the shard's largest |g|^2; dirty[row] is
// set where a candidate buffer overflowed outside the first chunk.  Both must be initialised by the caller.
static int coarse_pass(grl_handle* h, cudaStream_t st, int metric, const float* g, const void* prepared, int nq, int ng, int dim, int kprime,
                       int64_t idx_base, uint64_t* list, uint32_t* gmax2_bits, uint32_t* dirty, uint8_t* w, const CoarseLayout& L) {
    __half* q16 = (__half*)(w + L.q16);
    __half* g16 = (__half*)(w + L.g16);
    float* q_inv = (float*)(w + L.qf);
    float* q_n2 = q_inv + nq;
    float* g_inv = (float*)(w + L.gf);
    float* g_n2 = g_inv + L.chunk_alloc;
    float* tile = (float*)(w + L.tile);
    float* thresh = (float*)(w + L.thresh);
    int* cand_cnt = (int*)(w + L.cnt);
    unsigned long long* cand = (unsigned long long*)(w + L.cand);
    const PreparedLayout P = prepared_layout(ng, dim);
    const uint8_t* pb = (const uint8_t*)prepared;
    const size_t list_smem = (size_t)LIST_WARPS * (kprime + LIST_WARP_MAX) * 8;
    GRL_TRY(ensure_dyn_smem(h, (const void*)list_update_warp_kernel, (int)list_smem));
    topk_filter_init_kernel<<<(nq + 255) / 256, 256, 0, st>>>(thresh, cand_cnt, nq, L.cap);
    GRL_LAUNCH_CHECK(h);
    {
        const long long n = (long long)nq * kprime;
        fill_keys_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(list, n);
        GRL_LAUNCH_CHECK(h);
    }
    int c_merge = 0;                                  // columns seen at the last merge
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
            f16_rows_kernel<<<(int)(((long long)nc * 32 + 255) / 256), 256, 0, st>>>(g + (size_t)c0 * dim, nc, dim, g16, g_inv, g_n2, gmax2_bits);
            GRL_LAUNCH_CHECK(h);
        }
        GemmEpi e = epi_default();
        if (first) { e.C = tile; e.ldc = L.first; }   // later chunks never store their tile
        e.row_scale = q_inv; e.col_scale = g_invc;
        if (metric == GRL_METRIC_L2) { e.mode = 2; e.row_norm = q_n2; e.col_norm = g_n2c; }
        else e.alpha = -1.f;
        // the epilogue keeps only distances that can still enter a row's list (v <= current K'-th best) as candidates
        e.tk_cand = cand; e.tk_cnt = cand_cnt; e.tk_thresh = thresh; e.tk_cap = L.cap; e.tk_idx_base = idx_base + c0;
        GRL_TRY(coarse_gemm_launch(h, st, nq, nc, dim, q16, dim, g16c, dim, e));
        // Lazy merging: candidates may stay in the row buffers over several chunks (the filter threshold then dates from the
        // last merge at c_merge columns, so ~K' * cols / c_merge candidates per row arrive over `cols` further columns).
        // Merge when the buffers could overflow during the next chunk (1.6x the expectation + 48 of slack), and at the end.
        const int c_end = c0 + nc;
        const bool last = c_end >= ng;
        bool merge = last;
        if (!merge) {
            const int nc_next = topk_next_chunk(c_end, ng, L.chunk);
            const double pending = (double)kprime * (double)(c_end - c_merge + nc_next) / (double)c_merge;
            merge = 1.6 * pending + 48.0 > (double)L.cap;
        }
        if (first) {
            list_boot_select_kernel<<<nq, BOOT_THREADS, 0, st>>>(tile, L.first, nc, kprime, idx_base, list, thresh, cand_cnt);
            GRL_LAUNCH_CHECK(h);
            c_merge = c_end;
        } else if (merge) {
            list_update_warp_kernel<<<(nq + LIST_WARPS - 1) / LIST_WARPS, LIST_WARPS * 32, list_smem, st>>>(
                nq, kprime, nullptr, 0, nc, idx_base + c0, list, thresh, cand, cand_cnt, L.cap, dirty);
            GRL_LAUNCH_CHECK(h);
            if (L.cap > LIST_WARP_MAX && !first) {    // K' >= 512: a row can hold more candidates than one warp sorts
                list_update_kernel<<<nq, TOPK_THREADS, 0, st>>>(kprime, LIST_WARP_MAX, list, thresh, cand, cand_cnt, L.cap);
                GRL_LAUNCH_CHECK(h);
            }
            c_merge = c_end;
        }
        c0 += nc;
    }
    if (prepared) {   // the prepared index carries the shard's largest squared norm
        // max-combine (the value is a non-negative float: its bit pattern orders like an unsigned integer)
        GRL_CUDA(h, cudaMemcpyAsync(gmax2_bits, pb + P.gmax2, 4, cudaMemcpyDeviceToDevice, st));
    }
    return GRL_OK;
}

// ---- staged API (one shard; the caller runs its own collectives between the stages)
struct StagedCoarseLayout { CoarseLayout C; size_t coarse, list, dirty, total; };
static void staged_coarse_layout(int nq, int ng, int dim, int kprime, StagedCoarseLayout* S) {
    coarse_layout(nq, ng, dim, kprime, false, &S->C);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    S->coarse = take(S->C.total);
    S->list = take((size_t)nq * kprime * 8);
    S->dirty = take((size_t)nq * 4);
    S->total = off;
}
extern "C" size_t grl_coarse_topk_workspace_bytes(int nq, int ng, int dim) {
    StagedCoarseLayout S;
    staged_coarse_layout(nq, ng, dim, TOPK_MAXK, &S);     // sized for the largest K'
    return S.total;
}

static int coarse_topk_impl(grl_handle* h, int metric, const float* q, const float* g, const void* prepared, int nq, int ng, int dim, int kprime,
                            int64_t idx_base, float* coarse_d, int64_t* coarse_i, float* gmax2, int32_t* dirty, void* workspace,
                            size_t workspace_bytes, void* stream) {
    if (!h || !q || (!g && !prepared) || !coarse_d || !coarse_i || !gmax2 || !dirty || !workspace) return set_error(h, GRL_EINVAL, "grl_coarse_topk: NULL argument");
    if (nq <= 0 || ng <= 0 || dim <= 0 || (dim & 7)) return set_error(h, GRL_EINVAL, "grl_coarse_topk: need nq,ng > 0 and dim %% 8 == 0 (dim=%d)", dim);
    if (kprime <= 0 || kprime > TOPK_MAXK) return set_error(h, GRL_EINVAL, "grl_coarse_topk: need 0 < kprime <= %d", TOPK_MAXK);
    if (metric != GRL_METRIC_NEG_DOT && metric != GRL_METRIC_L2) return set_error(h, GRL_EINVAL, "grl_coarse_topk: unknown metric %d", metric);
    if (idx_base < 0 || idx_base + ng > 0xFFFFFFFFll) return set_error(h, GRL_EINVAL, "grl_coarse_topk: global index must fit 32 bits");
    StagedCoarseLayout S;
    staged_coarse_layout(nq, ng, dim, kprime, &S);
    if (workspace_bytes < S.total) return set_error(h, GRL_ENOMEM, "grl_coarse_topk: workspace %zu < %zu bytes", workspace_bytes, S.total);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* w = (uint8_t*)workspace;
    uint64_t* list = (uint64_t*)(w + S.list);
    GRL_CUDA(h, cudaMemsetAsync(gmax2, 0, 4, st));
    GRL_CUDA(h, cudaMemsetAsync(dirty, 0, (size_t)nq * 4, st));
    GRL_TRY(convert_queries(h, st, q, nq, dim, w + S.coarse, S.C));
    GRL_TRY(coarse_pass(h, st, metric, g, prepared, nq, ng, dim, kprime, idx_base, list, (uint32_t*)gmax2, (uint32_t*)dirty, w + S.coarse, S.C));
    const long long n = (long long)nq * kprime;
    unpack_keys_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(list, kprime, nq, kprime, coarse_d, coarse_i);
    GRL_LAUNCH_CHECK(h);
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
// Queries rows[r0 .. r0 + nrows) (or r0 .. when rows == NULL) against the shard: out_d / out_i [nrows][k], contiguous.
// nrows_dev != NULL: device-side row count gating the work (the host passes an upper bound as nrows).
static int exact_topk_rows(grl_handle* h, int metric, const float* q, const int32_t* rows, int r0, int nrows, const int32_t* nrows_dev,
                           const float* g, int ng, int dim, int k, int64_t idx_base, float* out_d, int64_t* out_i, float* tile, cudaStream_t st) {
    const int R = exact_group_rows(dim);
    GRL_TRY(ensure_dyn_smem(h, (const void*)exact_rows_kernel, (int)((size_t)R * dim * 4)));
    for (int r = 0; r < nrows; r += R) {
        const int rq = std::min(R, nrows - r);
        exact_rows_kernel<<<h->num_sms * 4, 256, (size_t)rq * dim * 4, st>>>(metric, q, rows, r0 + r, rq, nrows_dev, g, ng, dim, tile, ng);
        GRL_LAUNCH_CHECK(h);
        float* od = out_d + (size_t)r * k;
        int64_t* oi = out_i + (size_t)r * k;
        GRL_TRY(grl_topk_init(h, od, oi, rq, k, st));
        GRL_TRY(grl_topk_rows(h, tile, ng, rq, ng, k, idx_base, od, oi, st));
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
    return exact_topk_rows(h, metric, q, nullptr, 0, nq, nullptr, g, ng, dim, k, idx_base, top_d, top_i, (float*)workspace, (cudaStream_t)stream);
}

// ==================================================================================================== the fused search
constexpr int SEARCH_STAGES = 9;       // S0..S8 of the header comment (S9, the unpacking, is timed with S7)
constexpr int BRUTE_BATCH = 1024;      // flagged rows handled per round of the brute-force leg

// Buffers of one pass of the protocol (stages S1..S7) over `nqp` (padded) query rows with lists of kp keys.
struct CoreLayout {
    int nqp, qs, kp;
    CoarseLayout C;
    size_t coarse, L, meta, R, MA, E, Es, OUT, total;
};
static void core_layout(int world, int nq, int ng, int dim, int k, int kp, bool prepared, CoreLayout* S) {
    S->qs = (nq + world - 1) / world;
    S->nqp = S->qs * world;
    S->kp = kp;
    coarse_layout(S->nqp, ng, dim, kp, prepared, &S->C);
    const size_t nqp = S->nqp, qs = S->qs;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    S->coarse = take(S->C.total);
    S->L = take(nqp * kp * 8);                                  // local coarse lists
    S->meta = take((nqp + 1) * 4);                              // overflow marks [nqp] | max |g|^2 bits
    S->R = take(world > 1 ? nqp * kp * 8 : 0);                  // all-to-all receive: [world][qs][kp]
    S->MA = take(world > 1 ? nqp * kp * 8 : 0);                 // merged global lists of all slices (all-gathered)
    S->E = take(nqp * kp * 4);                                  // re-scored distances (owned candidates, 0 elsewhere)
    S->Es = take(world > 1 ? qs * kp * 4 : 0);                  // reduce-scattered: exact distances of the own slice
    S->OUT = take(nqp * (size_t)(k + 1) * 8);                   // result keys + flag word
    S->total = off;
}

struct SearchLayout {
    int world, fb;
    CoreLayout P1, P2;                                          // first pass (K' of k), second chance (K' = TOPK_MAXK)
    size_t Q, core1, Q2, core2, misc, rows, rows2, tile, btd, bti, bad, bai, bmd, bmi, total;
};
static void search_layout(int world, int nq, int ng, int dim, int k, bool prepared, SearchLayout* S) {
    S->world = world;
    S->fb = std::min(nq, BRUTE_BATCH);
    const int kp = grl_topk_kprime(k);
    core_layout(world, nq, ng, dim, k, kp, prepared, &S->P1);
    const bool second = kp < TOPK_MAXK;
    if (second) core_layout(world, nq, ng, dim, k, TOPK_MAXK, prepared, &S->P2);
    else memset(&S->P2, 0, sizeof(S->P2));
    const size_t nqp = S->P1.nqp;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    S->Q = take(world > 1 ? nqp * dim * 4 : 0);                 // all query rows (gathered / copied + zero padding)
    S->core1 = take(S->P1.total);
    S->Q2 = take(second ? nqp * dim * 4 : 0);                   // flagged query rows, gathered for the second chance
    S->core2 = take(S->P2.total);
    S->misc = take(64);                                         // nflag | nflag2
    S->rows = take(nqp * 4);                                    // flagged rows after the first pass
    S->rows2 = take(nqp * 4);                                   // flagged rows after the second chance (original row ids)
    S->tile = take(grl_exact_topk_workspace_bytes(nq, ng, dim));
    S->btd = take((size_t)S->fb * k * 4);
    S->bti = take((size_t)S->fb * k * 8);
    S->bad = take(world > 1 ? (size_t)world * S->fb * k * 4 : 0);
    S->bai = take(world > 1 ? (size_t)world * S->fb * k * 8 : 0);
    S->bmd = take(world > 1 ? (size_t)S->fb * k * 4 : 0);
    S->bmi = take(world > 1 ? (size_t)S->fb * k * 8 : 0);
    S->total = off;
}

static int stage_mark(grl_handle* h, cudaStream_t st, int i) {
    if (!h->stage_prof) return GRL_OK;
    if (!h->stage_ev) {
        h->stage_ev = new cudaEvent_t[SEARCH_STAGES + 1];
        for (int j = 0; j <= SEARCH_STAGES; ++j) GRL_CUDA(h, cudaEventCreate(&h->stage_ev[j]));
        h->n_stage_ev = SEARCH_STAGES + 1;
    }
    GRL_CUDA(h, cudaEventRecord(h->stage_ev[i], st));
    return GRL_OK;
}

extern "C" int grl_merge_key_lists(grl_handle* h, const unsigned long long* lists, int nlists, int rows, int kp, unsigned long long* out, void* stream) {
    if (!h || !lists || !out) return set_error(h, GRL_EINVAL, "grl_merge_key_lists: NULL argument");
    if (nlists <= 0 || rows <= 0 || kp < 2 || kp > TOPK_MAXK || (kp & (kp - 1))) return set_error(h, GRL_EINVAL, "grl_merge_key_lists: need nlists, rows > 0 and kp a power of two in [2, %d]", TOPK_MAXK);
    if ((long long)next_pow2_host(nlists) * kp > 16384) return set_error(h, GRL_EINVAL, "grl_merge_key_lists: nlists * kp must be <= 16384");
    return merge_key_lists(h, (cudaStream_t)stream, reinterpret_cast<const uint64_t*>(lists), nlists, rows, kp, reinterpret_cast<uint64_t*>(out));
}

extern "C" int grl_search_profile(grl_handle* h, int on) {
    if (!h) return GRL_EINVAL;
    h->stage_prof = on ? 1 : 0;
    return GRL_OK;
}
extern "C" int grl_search_stage_ms(grl_handle* h, double* ms, int n) {
    if (!h || !ms || n < SEARCH_STAGES) return set_error(h, GRL_EINVAL, "grl_search_stage_ms: need room for %d stages", SEARCH_STAGES);
    if (!h->stage_ev) return set_error(h, GRL_EINVAL, "grl_search_stage_ms: no profiled search yet (grl_search_profile)");
    GRL_CUDA(h, cudaEventSynchronize(h->stage_ev[SEARCH_STAGES]));
    for (int i = 0; i < SEARCH_STAGES; ++i) {
        float t = 0.f;
        GRL_CUDA(h, cudaEventElapsedTime(&t, h->stage_ev[i], h->stage_ev[i + 1]));
        ms[i] = t;
    }
    return GRL_OK;
}

// Stages S1..S7 for `nq` query rows Q (all of them on every rank; rows nq..nqp-1 are zero padding): leaves the result keys of
// every row + the flag word in OUT [nqp][k+1] on every rank and the (all-reduced) overflow marks in the pass's meta buffer.
// mark_stages: record the per-stage events of the handle (first pass only).
static int search_core(grl_handle* h, const NcclApi* api, ncclComm_t comm, int world, int rank, cudaStream_t st, int metric, const float* Q,
                       int nq, const float* g, const void* prepared, int ng, int dim, int k, int64_t idx_base, uint8_t* w, const CoreLayout& S,
                       int32_t* stats, bool mark_stages) {
    const int qs = S.qs, nqp = S.nqp, kp = S.kp;
    const int my_rows = std::max(0, std::min(qs, nq - rank * qs));
    uint8_t* cw = w + S.coarse;
    uint64_t* L = (uint64_t*)(w + S.L);
    uint32_t* dirty = (uint32_t*)(w + S.meta);
    uint32_t* gmax2_bits = dirty + nqp;
    float* E = (float*)(w + S.E);
    uint64_t* OUT = (uint64_t*)(w + S.OUT);
    const float ce = coarse_error_constant(dim);
    auto mark = [&](int i) { return mark_stages ? stage_mark(h, st, i) : GRL_OK; };

    // ---- S1 + S2: conversion and coarse pass over the local shard
    GRL_CUDA(h, cudaMemsetAsync(dirty, 0, (size_t)(nqp + 1) * 4, st));
    GRL_TRY(convert_queries(h, st, Q, nqp, dim, cw, S.C));
    const float* q_n2 = (const float*)(cw + S.C.qf) + nqp;
    GRL_TRY(mark(2));
    GRL_TRY(coarse_pass(h, st, metric, g, prepared, nqp, ng, dim, kp, idx_base, L, gmax2_bits, dirty, cw, S.C));
    GRL_TRY(mark(3));

    // ---- S3: exchange by query slice, merge, all-gather the merged lists
    const uint64_t* MA = L;
    if (world > 1) {
        uint64_t* R = (uint64_t*)(w + S.R);
        uint64_t* MAw = (uint64_t*)(w + S.MA);
        GRL_NCCL(h, api, api->AllReduce(dirty, dirty, (size_t)nqp + 1, ncclUint32, ncclMax, comm, st));
        const size_t cnt = (size_t)qs * kp;
        GRL_NCCL(h, api, api->GroupStart());
        for (int p = 0; p < world; ++p) {
            GRL_NCCL(h, api, api->Send(L + (size_t)p * cnt, cnt, ncclUint64, p, comm, st));
            GRL_NCCL(h, api, api->Recv(R + (size_t)p * cnt, cnt, ncclUint64, p, comm, st));
        }
        GRL_NCCL(h, api, api->GroupEnd());
        uint64_t* mine = MAw + (size_t)rank * cnt;
        GRL_TRY(merge_key_lists(h, st, R, world, qs, kp, mine));
        GRL_NCCL(h, api, api->AllGather(mine, MAw, cnt, ncclUint64, comm, st));                      // in place
        MA = MAw;
    }
    GRL_TRY(mark(4));

    // ---- S4: re-score the candidates this rank owns
    {
        const size_t smem = (size_t)dim * 4;
        GRL_TRY(ensure_dyn_smem(h, (const void*)rescore_keys_kernel, (int)smem));
        rescore_keys_kernel<<<nqp, 256, smem, st>>>(metric, Q, q_n2, g, ng, dim, idx_base, MA, kp, k, gmax2_bits, ce, E, stats);
        GRL_LAUNCH_CHECK(h);
    }
    GRL_TRY(mark(5));

    // ---- S5: exact distances of the own slice
    const float* Es = E;
    if (world > 1) {
        float* Esw = (float*)(w + S.Es);
        GRL_NCCL(h, api, api->ReduceScatter(E, Esw, (size_t)qs * kp, ncclFloat, ncclSum, comm, st));
        Es = Esw;
    }
    GRL_TRY(mark(6));

    // ---- S6: finalize the own slice
    {
        const int npad = next_pow2(kp < 2 ? 2 : kp);
        const size_t row0 = (size_t)rank * qs;
        finalize_keys_kernel<<<qs, 256, (size_t)npad * 8, st>>>(metric, q_n2 + row0, MA + row0 * kp, Es, kp, npad, gmax2_bits, ce, k, dirty + row0,
                                                                world > 1 ? my_rows : nq, OUT + row0 * (k + 1));
        GRL_LAUNCH_CHECK(h);
    }
    GRL_TRY(mark(7));

    // ---- S7 (first half): results of every slice on every rank
    if (world > 1) GRL_NCCL(h, api, api->AllGather(OUT + (size_t)rank * qs * (k + 1), OUT, (size_t)qs * (k + 1), ncclUint64, comm, st));
    return GRL_OK;
}

static int search_impl(grl_handle* h, int world, int rank, int metric, const float* q, int q_is_slice, const float* g, const void* prepared, int nq,
                       int ng, int dim, int k, int64_t idx_base, int max_flagged, float* top_d, int64_t* top_i, int32_t* stats, void* workspace,
                       size_t workspace_bytes, void* stream, const char* who) {
    if (!h || !g || !top_d || !top_i || !workspace) return set_error(h, GRL_EINVAL, "%s: NULL argument", who);
    if (nq <= 0 || ng <= 0 || dim <= 0 || (dim & 7) || dim > 32768) return set_error(h, GRL_EINVAL, "%s: need nq,ng > 0, dim %% 8 == 0, dim <= 32768 (dim=%d)", who, dim);
    if (k <= 0 || k > TOPK_MAXK / 2) return set_error(h, GRL_EINVAL, "%s: need 0 < k <= %d", who, TOPK_MAXK / 2);
    if (metric != GRL_METRIC_NEG_DOT && metric != GRL_METRIC_L2) return set_error(h, GRL_EINVAL, "%s: unknown metric %d", who, metric);
    if (idx_base < 0 || idx_base + ng > 0xFFFFFFFFll) return set_error(h, GRL_EINVAL, "%s: global index must fit 32 bits", who);
    SearchLayout S;
    search_layout(world, nq, ng, dim, k, prepared != nullptr, &S);
    if (workspace_bytes < S.total) return set_error(h, GRL_ENOMEM, "%s: workspace %zu < %zu bytes", who, workspace_bytes, S.total);
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return set_error(h, GRL_EINVAL, "%s: workspace must be 256-byte aligned", who);
    if ((long long)world * TOPK_MAXK > 16384) return set_error(h, GRL_EINVAL, "%s: world * K' must be <= 16384", who);
    const int qs = S.P1.qs, nqp = S.P1.nqp;
    const int my_rows = std::max(0, std::min(qs, nq - rank * qs));           // valid rows of this rank's query slice
    // q_is_slice decides which collectives run, so every rank must pass the same value (it cannot be inferred from a row count: a
    // slice can be as long as the whole block, e.g. one query on two ranks)
    if (!q && !(q_is_slice && my_rows == 0)) return set_error(h, GRL_EINVAL, "%s: NULL argument", who);   // (an empty query slice may be NULL)
    const NcclApi* api = nullptr;
    ncclComm_t comm = nullptr;
    if (world > 1) {
        api = nccl_api(h);
        if (!api) return GRL_ENCCL;
        comm = (ncclComm_t)h->comm;
    }
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* w = (uint8_t*)workspace;
    int32_t* nflag = (int32_t*)(w + S.misc);
    int32_t* nflag2 = nflag + 1;
    int32_t* rows = (int32_t*)(w + S.rows);
    if (stats) GRL_CUDA(h, cudaMemsetAsync(stats, 0, 8 * sizeof(int32_t), st));
    GRL_TRY(stage_mark(h, st, 0));

    // ---- S0: all query rows on every rank
    const float* Q = q;
    if (world > 1) {
        float* Qw = (float*)(w + S.Q);
        if (!q_is_slice) {
            GRL_CUDA(h, cudaMemcpyAsync(Qw, q, (size_t)nq * dim * 4, cudaMemcpyDeviceToDevice, st));
            if (nqp > nq) GRL_CUDA(h, cudaMemsetAsync(Qw + (size_t)nq * dim, 0, (size_t)(nqp - nq) * dim * 4, st));
        } else {
            float* mine = Qw + (size_t)rank * qs * dim;
            if (my_rows > 0) GRL_CUDA(h, cudaMemcpyAsync(mine, q, (size_t)my_rows * dim * 4, cudaMemcpyDeviceToDevice, st));
            if (my_rows < qs) GRL_CUDA(h, cudaMemsetAsync(mine + (size_t)my_rows * dim, 0, (size_t)(qs - my_rows) * dim * 4, st));
            GRL_NCCL(h, api, api->AllGather(mine, Qw, (size_t)qs * dim, ncclFloat, comm, st));      // in place
        }
        Q = Qw;
    }
    GRL_TRY(stage_mark(h, st, 1));

    // ---- S1..S7: the first pass
    uint8_t* w1 = w + S.core1;
    GRL_TRY(search_core(h, api, comm, world, rank, st, metric, Q, nq, g, prepared, ng, dim, k, idx_base, w1, S.P1, stats, true));
    uint64_t* OUT = (uint64_t*)(w1 + S.P1.OUT);
    const uint32_t* dirty = (const uint32_t*)(w1 + S.P1.meta);
    compact_flags_kernel<<<1, 1024, 0, st>>>(OUT, nq, k, dirty, rows, nflag, stats);
    GRL_LAUNCH_CHECK(h);
    {
        const long long n = (long long)nq * k;
        unpack_keys_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(OUT, k + 1, nq, k, top_d, top_i);
        GRL_LAUNCH_CHECK(h);
    }
    GRL_TRY(stage_mark(h, st, 8));

    // ---- S8: queries without a proof (identical row list on every rank)
    int todo = max_flagged;
    const int32_t* gate = nflag;
    const int32_t* brute_rows = rows;
    if (max_flagged < 0) {                            // synchronous mode: one 4-byte read tells the host how many rows there are
        int host_nflag = 0;
        GRL_CUDA(h, cudaMemcpyAsync(&host_nflag, nflag, 4, cudaMemcpyDeviceToHost, st));
        GRL_CUDA(h, cudaStreamSynchronize(st));
        todo = host_nflag;
        gate = nullptr;
        if (todo > 0 && S.P2.total > 0) {
            // Second chance: the same protocol over the flagged rows with the longest lists (K' = TOPK_MAXK).  A proof fails
            // when more than K' - k gallery rows sit within the coarse error of the k-th neighbour (clusters of near-duplicates,
            // the re-ID case); four times the list usually reaches past the cluster, at a fraction of the brute-force cost.
            float* Q2 = (float*)(w + S.Q2);
            CoreLayout P2;
            core_layout(world, todo, ng, dim, k, TOPK_MAXK, prepared != nullptr, &P2);   // offsets for `todo` rows (<= the reserved size)
            const long long n4 = (long long)todo * (dim >> 2);
            gather_rows_kernel<<<(int)((n4 + 255) / 256), 256, 0, st>>>(Q, rows, todo, dim, Q2);
            GRL_LAUNCH_CHECK(h);
            if (P2.nqp > todo) GRL_CUDA(h, cudaMemsetAsync(Q2 + (size_t)todo * dim, 0, (size_t)(P2.nqp - todo) * dim * 4, st));
            uint8_t* w2 = w + S.core2;
            GRL_TRY(search_core(h, api, comm, world, rank, st, metric, Q2, todo, g, prepared, ng, dim, k, idx_base, w2, P2, nullptr, false));
            uint64_t* OUT2 = (uint64_t*)(w2 + P2.OUT);
            int32_t* rows2 = (int32_t*)(w + S.rows2);
            // rows proven now: scatter their results; rows still flagged: compacted (as ORIGINAL row ids) for the brute force
            scatter_keys_kernel<<<todo, 128, 0, st>>>(OUT2, rows, k, top_d, top_i);
            GRL_LAUNCH_CHECK(h);
            compact_flags_kernel<<<1, 1024, 0, st>>>(OUT2, todo, k, nullptr, rows2, nflag2, nullptr);
            GRL_LAUNCH_CHECK(h);
            remap_rows_kernel<<<(todo + 255) / 256, 256, 0, st>>>(rows2, nflag2, rows);
            GRL_LAUNCH_CHECK(h);
            int host_nflag2 = 0;
            GRL_CUDA(h, cudaMemcpyAsync(&host_nflag2, nflag2, 4, cudaMemcpyDeviceToHost, st));
            GRL_CUDA(h, cudaStreamSynchronize(st));
            if (stats) {
                const int32_t v[2] = {todo - host_nflag2, host_nflag2};       // [4] proven by the second chance, [5] brute-forced
                GRL_CUDA(h, cudaMemcpyAsync(stats + 4, v, 8, cudaMemcpyHostToDevice, st));
                GRL_CUDA(h, cudaStreamSynchronize(st));                       // `v` lives on this stack frame
            }
            todo = host_nflag2;
            brute_rows = rows2;
        } else if (stats && todo > 0) {
            const int32_t v[2] = {0, todo};
            GRL_CUDA(h, cudaMemcpyAsync(stats + 4, v, 8, cudaMemcpyHostToDevice, st));
            GRL_CUDA(h, cudaStreamSynchronize(st));
        }
    } else if (todo > nq) todo = nq;
    float* btd = (float*)(w + S.btd);
    int64_t* bti = (int64_t*)(w + S.bti);
    for (int r0 = 0; r0 < todo; r0 += S.fb) {
        const int nb = std::min(S.fb, todo - r0);
        GRL_TRY(exact_topk_rows(h, metric, Q, brute_rows, r0, nb, gate, g, ng, dim, k, idx_base, btd, bti, (float*)(w + S.tile), st));
        const float* fd = btd;
        const int64_t* fi = bti;
        if (world > 1) {
            float* bad = (float*)(w + S.bad);
            int64_t* bai = (int64_t*)(w + S.bai);
            GRL_NCCL(h, api, api->GroupStart());
            GRL_NCCL(h, api, api->AllGather(btd, bad, (size_t)nb * k, ncclFloat, comm, st));
            GRL_NCCL(h, api, api->AllGather(bti, bai, (size_t)nb * k, ncclInt64, comm, st));
            GRL_NCCL(h, api, api->GroupEnd());
            GRL_TRY(grl_topk_merge(h, bad, bai, world, nb, k, (float*)(w + S.bmd), (int64_t*)(w + S.bmi), st));
            fd = (const float*)(w + S.bmd);
            fi = (const int64_t*)(w + S.bmi);
        }
        scatter_rows_kernel<<<nb, 128, 0, st>>>(fd, fi, brute_rows, r0, gate, k, top_d, top_i);
        GRL_LAUNCH_CHECK(h);
    }
    GRL_TRY(stage_mark(h, st, 9));
    return GRL_OK;
}

extern "C" size_t grl_sharded_topk_workspace_bytes(const grl_handle* h, int nq, int ng_local, int dim, int k, int prepared) {
    if (!h || nq <= 0 || ng_local <= 0 || dim <= 0 || k <= 0 || k > TOPK_MAXK / 2) return 0;
    SearchLayout S;
    search_layout(h->comm ? h->comm_world : 1, nq, ng_local, dim, k, prepared != 0, &S);
    return S.total;
}

extern "C" int grl_sharded_topk(grl_handle* h, int metric, const float* q, int q_is_slice, const float* g_local, const void* prepared, int nq,
                                int ng_local, int dim, int k, int64_t idx_base, int max_flagged, float* top_d, int64_t* top_i, int32_t* stats,
                                void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return GRL_EINVAL;
    const int world = h->comm ? h->comm_world : 1, rank = h->comm ? h->comm_rank : 0;
    return search_impl(h, world, rank, metric, q, q_is_slice, g_local, prepared, nq, ng_local, dim, k, idx_base, max_flagged, top_d, top_i, stats,
                       workspace, workspace_bytes, stream, "grl_sharded_topk");
}

// ---- one shard, end to end (the single-rank form of the same code path; ignores the handle's communicator)
extern "C" size_t grl_dist_topk_workspace_bytes(int nq, int ng, int dim) {
    if (nq <= 0 || ng <= 0 || dim <= 0) return 0;
    size_t need = 0;
    for (int k : {128, 256, TOPK_MAXK / 2}) {                   // one k per list length K' = 256 / 512 / 1024: sized for any supported k
        SearchLayout S;
        search_layout(1, nq, ng, dim, k, false, &S);              // unprepared gallery (the larger layout)
        need = std::max(need, S.total);
    }
    return need;
}
extern "C" int grl_dist_topk(grl_handle* h, int metric, const float* q, const float* g, int nq, int ng, int dim, int k,
                             int64_t idx_base, float* top_d, int64_t* top_i, void* workspace, size_t workspace_bytes, void* stream) {
    return search_impl(h, 1, 0, metric, q, 0, g, nullptr, nq, ng, dim, k, idx_base, -1, top_d, top_i, nullptr, workspace, workspace_bytes, stream,
                       "grl_dist_topk");
}
extern "C" int grl_dist_topk_prepared(grl_handle* h, int metric, const float* q, const float* g, const void* prepared, int nq, int ng, int dim,
                                      int k, int64_t idx_base, float* top_d, int64_t* top_i, void* workspace, size_t workspace_bytes,
                                      void* stream) {
    if (!prepared) return set_error(h, GRL_EINVAL, "grl_dist_topk_prepared: NULL argument");
    return search_impl(h, 1, 0, metric, q, 0, g, prepared, nq, ng, dim, k, idx_base, -1, top_d, top_i, nullptr, workspace, workspace_bytes, stream,
                       "grl_dist_topk_prepared");
}
