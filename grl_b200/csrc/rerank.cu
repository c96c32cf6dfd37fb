// grl_b200 — k-reciprocal re-ranking on the device (sm_100a).
//
// Replaces re_ranking(q_g_dist, q_q_dist, g_g_dist, k1, k2, lambda_value), reid/evaluator/rerank.py:37-104, called from
// ATTEvaluator.evaluate under `rerank` (reid/evaluator/attevaluator.py:151-155).  The reference builds three dense
// N x N float32 matrices (N = nq + ng = 11,310 on MARS) and walks them with Python loops for minutes; here
//
//   stage 1  rr_colmax_* + rr_scale_*    O[i][j] = fl32(D[j][i]^2 / max_r D[r][i]^2)  from the three blocks, coalesced both ways
//   stage 2  grl_topk_rows               (distance, index) top-max(k1+1, k2) of every row of O  (rerank.py:48 reads no more)
//   stage 3  rr_vrow_kernel              one CTA per row: k-reciprocal set, 2/3-overlap expansion, sort-unique, exp / sum
//   stage 4  rr_qe_kernel                one CTA per row: union of the k2 neighbour rows, float32 mean in rank order; the
//                                        result is stored sparse (query rows) and as the dense transpose Vt[c][row]
//   stage 5  rr_jaccard_kernel           thread == gallery column: t += min(V[i][c], Vt[c][j]) over the query's non-zero
//                                        columns in ascending order (coalesced rows of Vt), Jaccard, lambda blend
//
// are HBM/L2-bound integer and float32 work that reproduces the reference's float32 operation order (np.sum's pairwise
// scheme included); the only non-bit-exact step is exp (numpy's float32 exp is ~2 ulp, here exp is evaluated in fp64).
// Ties in stage 2 are broken by the lower index (numpy's introsort leaves them implementation-defined).
#include <math_constants.h>

#include <algorithm>

#include "api.h"

namespace grl {

constexpr int RR_THREADS = 128;

// ------------------------------------------------------------------ stage 1: column maxima of the squared block matrix
// Values are >= 0, so their float bit patterns order like unsigned integers.
__global__ void rr_colmax_cols_kernel(const float* __restrict__ src, long long ld, int rows, int cols, int rows_per_block,
                                      unsigned int* __restrict__ cmax) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(rows, r0 + rows_per_block);
    float m = 0.f;
    for (int r = r0; r < r1; ++r) {
        const float d = src[(long long)r * ld + c];
        m = fmaxf(m, d * d);
    }
    atomicMax(cmax + c, __float_as_uint(m));
}
__global__ void rr_colmax_rows_kernel(const float* __restrict__ src, long long ld, int rows, int cols,
                                      unsigned int* __restrict__ cmax) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    float m = 0.f;
    for (int c = lane_id(); c < cols; c += 32) {
        const float d = src[(long long)r * ld + c];
        m = fmaxf(m, d * d);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (lane_id() == 0) atomicMax(cmax + r, __float_as_uint(m));
}
// dst[i][j] = src[j][i]^2 / cmax[i]   (i < n_i, j < n_j), 32x32 tiles through shared memory
__global__ void rr_scale_transposed_kernel(const float* __restrict__ src, long long ld_src, int n_i, int n_j,
                                           const float* __restrict__ cmax, float* __restrict__ dst, long long ld_dst) {
    __shared__ float tile[32][33];
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = j0 + r, i = i0 + threadIdx.x;
        if (j < n_j && i < n_i) {
            const float d = src[(long long)j * ld_src + i];
            tile[r][threadIdx.x] = d * d;
        }
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, j = j0 + threadIdx.x;
        if (i < n_i && j < n_j) dst[(long long)i * ld_dst + j] = __fdiv_rn(tile[threadIdx.x][r], cmax[i]);
    }
}
// dst[i][j] = src[i][j]^2 / cmax[i]
__global__ void rr_scale_direct_kernel(const float* __restrict__ src, long long ld_src, int n_i, int n_j,
                                       const float* __restrict__ cmax, float* __restrict__ dst, long long ld_dst) {
    const int i = blockIdx.y;
    const float cm = cmax[i];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_j; j += gridDim.x * blockDim.x) {
        const float d = src[(long long)i * ld_src + j];
        dst[(long long)i * ld_dst + j] = __fdiv_rn(d * d, cm);
    }
}

// ------------------------------------------------------------------ helpers
// np.sum over a contiguous float32 vector (numpy's pairwise scheme: 8 strided partial sums per <=128-element block,
// halves split on multiples of 8).  Verified bit-exact against numpy 2.3 in tests/test_oracle_rerank.py's restatement.
__device__ float np_pairwise_sum(const float* a, int n) {
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
        return r;
    }
    if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
}

// ascending bitonic sort of n (power of two) 32-bit keys in shared memory by the whole block
__device__ void block_sort_u32(uint32_t* keys, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint32_t a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    __syncthreads();
}
// In-place removal of duplicates from sorted keys[0..n) (0xFFFFFFFF padding is dropped); returns the new length to all
// threads.  `scratch` holds blockDim.x/32 + 1 ints.  Chunks of blockDim.x keys are compacted in order, so a write never
// overtakes an unread key.
__device__ int block_unique_u32(uint32_t* keys, int n, int* scratch) {
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = lane_id();
    int base = 0;
    for (int c0 = 0; c0 < n; c0 += blockDim.x) {
        const int t = c0 + threadIdx.x;
        uint32_t key = 0xFFFFFFFFu;
        bool keep = false;
        if (t < n) {
            key = keys[t];
            keep = key != 0xFFFFFFFFu && (t == 0 || keys[t - 1] != key);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) scratch[warp] = __popc(ballot);
        __syncthreads();                                   // all reads of this chunk are done
        int off = base;
        for (int w = 0; w < warp; ++w) off += scratch[w];
        int total = 0;
        for (int w = 0; w < nwarps; ++w) total += scratch[w];
        if (keep) keys[off + __popc(ballot & ((1u << lane) - 1))] = key;
        base += total;
        __syncthreads();
    }
    return base;
}

// ------------------------------------------------------------------ stage 3: one row of V (rerank.py:54-76)
// rank [N][ldr] int64 (sorted neighbours, column 0 is normally the row itself), O [N][N].
// v_idx/v_w [N][cap]: ascending unique expansion indices and normalised weights, v_cnt [N].
__global__ void __launch_bounds__(RR_THREADS) rr_vrow_kernel(const int64_t* __restrict__ rank, int ldr, const float* __restrict__ O,
                                                             int N, int k1, int half, int cap, int npad,
                                                             int32_t* __restrict__ v_idx, float* __restrict__ v_w,
                                                             int32_t* __restrict__ v_cnt) {
    extern __shared__ uint32_t rr_smem[];
    uint32_t* exp_idx = rr_smem;                              // [npad]
    float* wts = reinterpret_cast<float*>(rr_smem + npad);    // [cap]
    __shared__ int R[32];
    __shared__ int nR, nE;
    __shared__ int scratch[RR_THREADS / 32 + 1];
    __shared__ float wsum;
    const int i = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int kf = min(k1 + 1, N), kh = min(half + 1, N);
    for (int t = threadIdx.x; t < npad; t += blockDim.x) exp_idx[t] = 0xFFFFFFFFu;
    __syncthreads();
    if (warp == 0) {                                          // R(i, k1): forward neighbours that list i back (:56-59)
        bool in = false;
        int f = -1;
        if (lane < kf) {
            f = (int)rank[(long long)i * ldr + lane];
            const int64_t* rf = rank + (long long)f * ldr;
            for (int u = 0; u < kf; ++u) in |= (rf[u] == i);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, in);
        if (in) {
            const int pos = __popc(ballot & ((1u << lane) - 1));
            R[pos] = f;
            exp_idx[pos] = (uint32_t)f;
        }
        if (lane == 0) { nR = __popc(ballot); nE = __popc(ballot); }
    }
    __syncthreads();
    const int nr = nR;
    for (int a = warp; a < nr; a += (blockDim.x >> 5)) {      // expansion by R(c, k1/2) of every member c (:61-70)
        const int c = R[a];
        bool in = false;
        int f = -1;
        if (lane < kh) {
            f = (int)rank[(long long)c * ldr + lane];
            const int64_t* rf = rank + (long long)f * ldr;
            for (int u = 0; u < kh; ++u) in |= (rf[u] == c);
        }
        bool common = false;
        if (in) for (int u = 0; u < nr; ++u) common |= (R[u] == f);
        const unsigned bin = __ballot_sync(0xffffffffu, in);
        const unsigned bcm = __ballot_sync(0xffffffffu, common);
        const int len = __popc(bin);
        if ((double)__popc(bcm) > 2. / 3 * (double)len) {     // same double comparison as :68-69
            int base = 0;
            if (lane == 0) base = atomicAdd(&nE, len);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (in) exp_idx[base + __popc(bin & ((1u << lane) - 1))] = (uint32_t)f;
        }
    }
    __syncthreads();
    block_sort_u32(exp_idx, npad);
    const int n = block_unique_u32(exp_idx, min(nE, npad), scratch);          // np.unique (:72)
    const float* orow = O + (long long)i * N;
    for (int t = threadIdx.x; t < n; t += blockDim.x) wts[t] = (float)exp(-(double)orow[exp_idx[t]]);   // :73
    __syncthreads();
    if (threadIdx.x == 0) wsum = np_pairwise_sum(wts, n);
    __syncthreads();
    const float s = wsum;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        v_idx[(long long)i * cap + t] = (int32_t)exp_idx[t];
        v_w[(long long)i * cap + t] = __fdiv_rn(wts[t], s);                    // :74
    }
    if (threadIdx.x == 0) v_cnt[i] = n;
}

// ------------------------------------------------------------------ stage 4: local query expansion (rerank.py:78-83)
// Row i of V_qe = float32 mean of the V rows of its k2 nearest neighbours, added in rank order.  Written as the dense
// transpose Vt[c][i] (zero-initialised by the caller) and, for the query rows, as a sparse row for stage 5.
// k2 == 1: the reference skips the stage, i.e. row i is V[i] itself.
__global__ void __launch_bounds__(RR_THREADS) rr_qe_kernel(const int64_t* __restrict__ rank, int ldr, int N, int nq, int k2, int cap,
                                                           int qcap, int npad, const int32_t* __restrict__ v_idx,
                                                           const float* __restrict__ v_w, const int32_t* __restrict__ v_cnt,
                                                           float* __restrict__ Vt, int32_t* __restrict__ q_idx,
                                                           float* __restrict__ q_w, int32_t* __restrict__ q_cnt) {
    extern __shared__ uint32_t rr_smem[];
    uint32_t* uni = rr_smem;                                  // [npad]
    __shared__ int rows[64];
    __shared__ int scratch[RR_THREADS / 32 + 1];
    const int i = blockIdx.x;
    const int kk = min(k2, N);
    if (threadIdx.x < kk) rows[threadIdx.x] = (k2 == 1) ? i : (int)rank[(long long)i * ldr + threadIdx.x];
    for (int t = threadIdx.x; t < npad; t += blockDim.x) uni[t] = 0xFFFFFFFFu;
    __syncthreads();
    int total = 0;
    for (int r = 0; r < kk; ++r) {
        const int row = rows[r], n = v_cnt[row];
        for (int t = threadIdx.x; t < n; t += blockDim.x) uni[total + t] = (uint32_t)v_idx[(long long)row * cap + t];
        total += n;
    }
    __syncthreads();
    int n = total;
    if (kk > 1) {
        block_sort_u32(uni, npad);
        n = block_unique_u32(uni, min(total, npad), scratch);
    }
    const float denom = (float)kk;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int c = (int)uni[t];
        float acc = 0.f;
        for (int r = 0; r < kk; ++r) {
            const int row = rows[r];
            const int32_t* idx = v_idx + (long long)row * cap;
            int lo = 0, hi = v_cnt[row];
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (idx[mid] < c) lo = mid + 1; else hi = mid;
            }
            if (lo < v_cnt[row] && idx[lo] == c) acc = __fadd_rn(acc, v_w[(long long)row * cap + lo]);
        }
        const float v = (k2 == 1) ? acc : __fdiv_rn(acc, denom);
        Vt[(long long)c * N + i] = v;
        if (i < nq) { q_idx[(long long)i * qcap + t] = c; q_w[(long long)i * qcap + t] = v; }
    }
    if (i < nq && threadIdx.x == 0) q_cnt[i] = n;
}

// ------------------------------------------------------------------ stage 5 + 6: Jaccard distance and blend (:86-104)
constexpr int RRJ_THREADS = 256;
constexpr int RRJ_CHUNK = 512;
__global__ void __launch_bounds__(RRJ_THREADS) rr_jaccard_kernel(const float* __restrict__ Vt, const float* __restrict__ O, int N, int nq,
                                                                 int qcap, const int32_t* __restrict__ q_idx,
                                                                 const float* __restrict__ q_w, const int32_t* __restrict__ q_cnt,
                                                                 float w_jac, float w_orig, float* __restrict__ out, long long ld_out) {
    __shared__ int s_c[RRJ_CHUNK];
    __shared__ float s_a[RRJ_CHUNK];
    const int i = blockIdx.y;
    const int j = nq + blockIdx.x * RRJ_THREADS + threadIdx.x;
    const bool live = j < N;
    const int n = q_cnt[i];
    float t = 0.f;
    for (int c0 = 0; c0 < n; c0 += RRJ_CHUNK) {
        const int m = min(RRJ_CHUNK, n - c0);
        __syncthreads();
        for (int u = threadIdx.x; u < m; u += blockDim.x) {
            s_c[u] = q_idx[(long long)i * qcap + c0 + u];
            s_a[u] = q_w[(long long)i * qcap + c0 + u];
        }
        __syncthreads();
        if (live) {
            int u = 0;
            for (; u + 8 <= m; u += 8) {                      // 8 independent row loads in flight, adds stay in order
                float b[8];
#pragma unroll
                for (int v = 0; v < 8; ++v) b[v] = __ldg(Vt + (long long)s_c[u + v] * N + j);
#pragma unroll
                for (int v = 0; v < 8; ++v) t = __fadd_rn(t, fminf(s_a[u + v], b[v]));
            }
            for (; u < m; ++u) t = __fadd_rn(t, fminf(s_a[u], __ldg(Vt + (long long)s_c[u] * N + j)));
        }
    }
    if (live) {
        const float jac = __fsub_rn(1.f, __fdiv_rn(t, __fsub_rn(2.f, t)));                         // :98
        out[(long long)i * ld_out + (j - nq)] =
            __fadd_rn(__fmul_rn(jac, w_jac), __fmul_rn(O[(long long)i * N + j], w_orig));          // :100, :104
    }
}

static int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

struct RerankPlan {
    int N, K, half, cap, qcap, npad_v, npad_q;
    size_t off_O, off_Vt, off_cmax, off_topd, off_topi, off_vidx, off_vw, off_vcnt, off_qidx, off_qw, off_qcnt, total;
};

static int rerank_plan(int nq, int ng, int k1, int k2, RerankPlan* p) {
    if (nq <= 0 || ng <= 0 || k1 < 1 || k1 > 31 || k2 < 1 || k2 > 32) return GRL_EINVAL;
    const long long N = (long long)nq + ng;
    if (N > 65536) return GRL_EINVAL;
    p->N = (int)N;
    p->K = (int)(N < (k1 + 1 > k2 ? k1 + 1 : k2) ? N : (k1 + 1 > k2 ? k1 + 1 : k2));
    // int(np.around(k1 / 2.)): round half to even
    const int fl = k1 / 2;
    p->half = (k1 % 2 == 0) ? fl : ((fl % 2 == 0) ? fl : fl + 1);
    p->cap = (k1 + 1) * (p->half + 2);
    p->npad_v = next_pow2(p->cap);
    p->qcap = k2 * p->cap;
    p->npad_q = next_pow2(p->qcap);
    if (p->npad_q > 16384) return GRL_EINVAL;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    p->off_O = take((size_t)N * N * 4);
    p->off_Vt = take((size_t)N * N * 4);
    p->off_cmax = take((size_t)N * 4);
    p->off_topd = take((size_t)N * p->K * 4);
    p->off_topi = take((size_t)N * p->K * 8);
    p->off_vidx = take((size_t)N * p->cap * 4);
    p->off_vw = take((size_t)N * p->cap * 4);
    p->off_vcnt = take((size_t)N * 4);
    p->off_qidx = take((size_t)nq * p->qcap * 4);
    p->off_qw = take((size_t)nq * p->qcap * 4);
    p->off_qcnt = take((size_t)nq * 4);
    p->total = off;
    return GRL_OK;
}

}  // namespace grl

using namespace grl;

extern "C" size_t grl_rerank_workspace_bytes(int nq, int ng, int k1, int k2) {
    RerankPlan p;
    if (rerank_plan(nq, ng, k1, k2, &p) != GRL_OK) return 0;
    return p.total;
}

extern "C" int grl_rerank(grl_handle* h, const float* q_g, const float* q_q, const float* g_g, int nq, int ng, int k1, int k2,
                          double lambda_value, float* final_dist, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return GRL_EINVAL;
    if (!q_g || !q_q || !g_g || !final_dist || !workspace) return set_error(h, GRL_EINVAL, "grl_rerank: null pointer");
    RerankPlan p;
    if (rerank_plan(nq, ng, k1, k2, &p) != GRL_OK)
        return set_error(h, GRL_EINVAL, "grl_rerank: need nq, ng > 0, nq + ng <= 65536, 1 <= k1 <= 31, 1 <= k2 <= 32 and k2*(k1+1)*(round(k1/2)+2) <= 16384");
    if (workspace_bytes < p.total) return set_error(h, GRL_ENOMEM, "grl_rerank: workspace %zu < %zu bytes", workspace_bytes, p.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    const int N = p.N;
    float* O = (float*)(ws + p.off_O);
    float* Vt = (float*)(ws + p.off_Vt);
    float* cmax = (float*)(ws + p.off_cmax);
    float* top_d = (float*)(ws + p.off_topd);
    int64_t* top_i = (int64_t*)(ws + p.off_topi);
    int32_t* v_idx = (int32_t*)(ws + p.off_vidx);
    float* v_w = (float*)(ws + p.off_vw);
    int32_t* v_cnt = (int32_t*)(ws + p.off_vcnt);
    int32_t* q_idx = (int32_t*)(ws + p.off_qidx);
    float* q_w = (float*)(ws + p.off_qw);
    int32_t* q_cnt = (int32_t*)(ws + p.off_qcnt);

    // ---- stage 1: column maxima of [[qq, qg], [qg^T, gg]]^2, then O = (D^2 / colmax)^T
    GRL_CUDA(h, cudaMemsetAsync(cmax, 0, (size_t)N * 4, st));
    GRL_CUDA(h, cudaMemsetAsync(Vt, 0, (size_t)N * N * 4, st));
    const int RPB = 256;
    auto colmax_cols = [&](const float* src, long long ld, int rows, int cols, float* dstmax) -> int {
        dim3 grid((cols + 255) / 256, (rows + RPB - 1) / RPB);
        rr_colmax_cols_kernel<<<grid, 256, 0, st>>>(src, ld, rows, cols, RPB, (unsigned int*)dstmax);
        GRL_LAUNCH_CHECK(h);
        return GRL_OK;
    };
    GRL_TRY(colmax_cols(q_q, nq, nq, nq, cmax));              // columns < nq, rows < nq
    rr_colmax_rows_kernel<<<(nq + 7) / 8, 256, 0, st>>>(q_g, ng, nq, ng, (unsigned int*)cmax);   // columns < nq, rows >= nq (qg^T)
    GRL_LAUNCH_CHECK(h);
    GRL_TRY(colmax_cols(q_g, ng, nq, ng, cmax + nq));         // columns >= nq, rows < nq
    GRL_TRY(colmax_cols(g_g, ng, ng, ng, cmax + nq));         // columns >= nq, rows >= nq
    const dim3 tb(32, 8);
    // O[i < nq][j < nq] = qq[j][i]^2 / cmax[i]
    rr_scale_transposed_kernel<<<dim3((nq + 31) / 32, (nq + 31) / 32), tb, 0, st>>>(q_q, nq, nq, nq, cmax, O, N);
    GRL_LAUNCH_CHECK(h);
    // O[i < nq][nq + j] = D[nq + j][i] = qg[i][j]
    rr_scale_direct_kernel<<<dim3(std::min((ng + 255) / 256, 64), nq), 256, 0, st>>>(q_g, ng, nq, ng, cmax, O + nq, N);
    GRL_LAUNCH_CHECK(h);
    // O[nq + i][j < nq] = D[j][nq + i] = qg[j][i]
    rr_scale_transposed_kernel<<<dim3((ng + 31) / 32, (nq + 31) / 32), tb, 0, st>>>(q_g, ng, ng, nq, cmax + nq, O + (size_t)nq * N, N);
    GRL_LAUNCH_CHECK(h);
    // O[nq + i][nq + j] = gg[j][i]
    rr_scale_transposed_kernel<<<dim3((ng + 31) / 32, (ng + 31) / 32), tb, 0, st>>>(g_g, ng, ng, ng, cmax + nq, O + (size_t)nq * N + nq, N);
    GRL_LAUNCH_CHECK(h);

    // ---- stage 2: the only part of initial_rank the reference reads
    GRL_TRY(grl_topk_init(h, top_d, top_i, N, p.K, stream));
    GRL_TRY(grl_topk_rows(h, O, N, N, N, p.K, 0, top_d, top_i, stream));

    // ---- stage 3
    const size_t smem_v = (size_t)p.npad_v * 4 + (size_t)p.cap * 4;
    rr_vrow_kernel<<<N, RR_THREADS, smem_v, st>>>(top_i, p.K, O, N, k1, p.half, p.cap, p.npad_v, v_idx, v_w, v_cnt);
    GRL_LAUNCH_CHECK(h);
    // ---- stage 4
    const size_t smem_q = (size_t)p.npad_q * 4;
    GRL_TRY(ensure_dyn_smem(h, (const void*)rr_qe_kernel, (int)smem_q));
    rr_qe_kernel<<<N, RR_THREADS, smem_q, st>>>(top_i, p.K, N, nq, k2, p.cap, p.qcap, p.npad_q, v_idx, v_w, v_cnt, Vt, q_idx, q_w, q_cnt);
    GRL_LAUNCH_CHECK(h);
    // ---- stage 5 + 6
    const float w_jac = (float)(1 - lambda_value), w_orig = (float)lambda_value;
    rr_jaccard_kernel<<<dim3((ng + RRJ_THREADS - 1) / RRJ_THREADS, nq), RRJ_THREADS, 0, st>>>(Vt, O, N, nq, p.qcap, q_idx, q_w, q_cnt,
                                                                                                w_jac, w_orig, final_dist, ng);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}
