// grl_b200 — host side of the split-bf16 tcgen05 GEMM: TMA tensor maps, tile choice, launch,
// the fp32 -> bf16 hi/lo split kernels, handle lifecycle and the grl_gemm_bf16x3 C entry point.
#include "api.h"
#include "coarse_gemm.cuh"
#include "gemm_pair.cuh"

#include <utility>
#include <vector>

struct grl_prof { std::vector<grl_prof_rec> recs; };
typedef std::vector<std::pair<const void*, int>> grl_func_attrs;

namespace grl {

int set_error(grl_handle* h, int code, const char* fmt, ...) {
    static thread_local char scratch[512];
    char* dst = h ? h->err : scratch;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

int ensure_dyn_smem(grl_handle* h, const void* func, int bytes) {
    if (bytes <= 48 * 1024) return GRL_OK;
    grl_func_attrs* fa = static_cast<grl_func_attrs*>(h->func_attrs);
    for (auto& e : *fa)
        if (e.first == func) {
            if (e.second >= bytes) return GRL_OK;
            GRL_CUDA(h, cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            e.second = bytes;
            return GRL_OK;
        }
    GRL_CUDA(h, cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    fa->push_back(std::make_pair(func, bytes));
    return GRL_OK;
}

// ------------------------------------------------------------------ fork / join helpers
cudaEvent_t pool_event(grl_handle* h, int k) {
    if (k >= h->n_events) {
        int n = h->n_events ? h->n_events : 64;
        while (n <= k) n *= 2;
        cudaEvent_t* ev = new cudaEvent_t[n];
        for (int i = 0; i < n; ++i) {
            if (i < h->n_events) ev[i] = h->events[i];
            else cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
        }
        delete[] h->events;
        h->events = ev;
        h->n_events = n;
    }
    return h->events[k];
}
int ev_record(grl_handle* h, int k, cudaStream_t s) {
    GRL_CUDA(h, cudaEventRecord(pool_event(h, k), s));
    return GRL_OK;
}
int ev_wait(grl_handle* h, int k, cudaStream_t s) {
    GRL_CUDA(h, cudaStreamWaitEvent(s, pool_event(h, k), 0));
    return GRL_OK;
}
int stream_wait(grl_handle* h, cudaStream_t signaler, cudaStream_t waiter, int k) {
    GRL_TRY(ev_record(h, k, signaler));
    return ev_wait(h, k, waiter);
}

// ------------------------------------------------------------------ split kernels
__global__ void split_planes_kernel(const float* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, long long ld_dst, long long rows, int cols) {
    const int vec_per_row = cols >> 2;
    const long long total = rows * vec_per_row;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / vec_per_row;
        const int c = int(i - r * vec_per_row) << 2;
        const float4 v = *reinterpret_cast<const float4*>(src + r * ld_src + c);
        __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
        split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
        *reinterpret_cast<uint2*>(hi + r * ld_dst + c) = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
        *reinterpret_cast<uint2*>(lo + r * ld_dst + c) = make_uint2(pack_bf16(l0, l1), pack_bf16(l2, l3));
    }
}

int split_planes(grl_handle* h, cudaStream_t st, const float* src, long long ld_src, __nv_bfloat16* hi,
                 __nv_bfloat16* lo, long long ld_dst, long long rows, int cols) {
    if ((cols & 3) || (ld_src & 3) || (ld_dst & 3)) return set_error(h, GRL_EINVAL, "split_planes: cols/ld must be multiples of 4");
    const long long total = rows * (cols >> 2);
    if (total == 0) return GRL_OK;
    int blocks = (int)((total + 255) / 256);
    const int cap = h->num_sms * 16;
    if (blocks > cap) blocks = cap;
    split_planes_kernel<<<blocks, 256, 0, st>>>(src, ld_src, hi, lo, ld_dst, rows, cols);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

// 32x32 smem-tiled transpose + split: src [rows][cols] -> planes [cols][rows]
__global__ void split_planes_t_kernel(const float* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ hi,
                                      __nv_bfloat16* __restrict__ lo, long long ld_dst, int rows, int cols) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (r < rows && c < cols) ? src[(long long)r * ld_src + c] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (c < cols && r < rows) {
            __nv_bfloat16 hh, ll;
            split_bf16(tile[threadIdx.x][j], hh, ll);
            hi[(long long)c * ld_dst + r] = hh;
            lo[(long long)c * ld_dst + r] = ll;
        }
    }
}

int split_planes_transposed(grl_handle* h, cudaStream_t st, const float* src, long long ld_src, __nv_bfloat16* hi,
                            __nv_bfloat16* lo, long long ld_dst, int rows, int cols) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    split_planes_t_kernel<<<grid, block, 0, st>>>(src, ld_src, hi, lo, ld_dst, rows, cols);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

// ------------------------------------------------------------------ tensor maps
static int make_tmap(grl_handle* h, CUtensorMap* map, const void* base, long long ld, long long bstride,
                     int mn_major, long long rows, long long K, int batch, int tile_rows,
                     CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 7) || (batch > 1 && (bstride & 7)))
        return set_error(h, GRL_EINVAL, "gemm operand must be 16-byte aligned with ld %% 8 == 0 (ld=%lld)", ld);
    cuuint64_t dims[3];
    cuuint64_t strides[2];
    cuuint32_t box[3];
    cuuint32_t estr[3] = {1, 1, 1};
    const long long outer = mn_major ? K : rows;
    if (mn_major) {
        dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K;
        box[0] = 64; box[1] = GEMM_BK;
    } else {
        dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows;
        box[0] = GEMM_BK; box[1] = (cuuint32_t)tile_rows;
    }
    dims[2] = (cuuint64_t)batch;
    box[2] = 1;
    strides[0] = (cuuint64_t)ld * 2;
    strides[1] = (cuuint64_t)((batch > 1) ? bstride : outer * ld) * 2;
    CUresult r = h->encode(map, dtype, 3, const_cast<void*>(base), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(h, GRL_ECUDA, "cuTensorMapEncodeTiled failed (%d): rows=%lld K=%lld ld=%lld batch=%d mn=%d", (int)r,
                         rows, K, ld, batch, mn_major);
    return GRL_OK;
}

template <int BN, bool A_MN, bool B_MN, int PLANES = 3, int EW = 4>
static int launch_variant(grl_handle* h, cudaStream_t st, const GemmParams& p, int grid) {
    auto kern = gemm_bf16x3_kernel<BN, A_MN, B_MN, PLANES, EW>;
    GRL_TRY(ensure_dyn_smem(h, (const void*)kern, GemmCfg<BN, PLANES>::SMEM_BYTES));
    grl_prof_rec rec;
    if (h->prof_on) {    // bench.py roofline: bracket the launch with events on the caller's stream
        GRL_CUDA(h, cudaEventCreate(&rec.e0));
        GRL_CUDA(h, cudaEventCreate(&rec.e1));
        rec.flops = 2.0 * p.M * (double)p.N * p.K * p.batch;
        GRL_CUDA(h, cudaEventRecord(rec.e0, st));
    }
    kern<<<grid, 64 + 32 * EW, GemmCfg<BN, PLANES>::SMEM_BYTES, st>>>(p);
    GRL_LAUNCH_CHECK(h);
    if (h->prof_on) {
        GRL_CUDA(h, cudaEventRecord(rec.e1, st));
        h->prof->recs.push_back(rec);
    }
    return GRL_OK;
}

template <bool A_MN, bool B_MN, int PLANES = 3>
static int launch_pair_variant(grl_handle* h, cudaStream_t st, const GemmParams& p, int grid) {
    auto kern = gemm_pair_bf16x3_kernel<A_MN, B_MN, PLANES>;
    constexpr int GP_SMEM_BYTES = GpCfg<PLANES>::SMEM_BYTES;
    GRL_TRY(ensure_dyn_smem(h, (const void*)kern, GP_SMEM_BYTES));
    grl_prof_rec rec;
    if (h->prof_on) {
        GRL_CUDA(h, cudaEventCreate(&rec.e0));
        GRL_CUDA(h, cudaEventCreate(&rec.e1));
        rec.flops = 2.0 * p.M * (double)p.N * p.K * p.batch;
        GRL_CUDA(h, cudaEventRecord(rec.e0, st));
    }
    kern<<<grid, GEMM_THREADS, GP_SMEM_BYTES, st>>>(p);
    GRL_LAUNCH_CHECK(h);
    if (h->prof_on) {
        GRL_CUDA(h, cudaEventRecord(rec.e1, st));
        h->prof->recs.push_back(rec);
    }
    return GRL_OK;
}

// The CTA-pair kernel (gemm_pair.cuh): 256 x 256 tiles, each CTA loads 128-row boxes of A and of B.  a_lo == NULL: one fp16
// plane per operand (single MMA per k-step), else bf16 hi/lo planes.
static int gemm_launch_pair(grl_handle* h, cudaStream_t st, int M, int N, int K, int batch, const void* a_hi, const void* a_lo, long long lda,
                            long long a_bstride, int a_mn, const void* b_hi, const void* b_lo, long long ldb, long long b_bstride, int b_mn,
                            GemmEpi epi) {
    const bool x1 = a_lo == nullptr;
    const CUtensorMapDataType dt = x1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K; p.batch = batch;
    p.num_m_tiles = (M + GP_BM - 1) / GP_BM;
    p.num_n_tiles = (N + GP_BN - 1) / GP_BN;
    p.group_m = p.num_m_tiles > 16 ? 8 : p.num_m_tiles;
    p.epi = epi;
    GRL_TRY(make_tmap(h, &p.ta_hi, a_hi, lda, a_bstride, a_mn, M, K, batch, 128, dt));
    GRL_TRY(make_tmap(h, &p.tb_hi, b_hi, ldb, b_bstride, b_mn, N, K, batch, 128, dt));
    if (x1) { p.ta_lo = p.ta_hi; p.tb_lo = p.tb_hi; }
    else {
        GRL_TRY(make_tmap(h, &p.ta_lo, a_lo, lda, a_bstride, a_mn, M, K, batch, 128, dt));
        GRL_TRY(make_tmap(h, &p.tb_lo, b_lo, ldb, b_bstride, b_mn, N, K, batch, 128, dt));
    }
    const long long tiles = (long long)p.num_m_tiles * p.num_n_tiles * batch;
    const long long pairs = tiles < (long long)(h->num_sms / 2) ? tiles : (long long)(h->num_sms / 2);
    const int grid = (int)(2 * pairs);
    if (x1) {
        if (a_mn) return launch_pair_variant<true, true, 1>(h, st, p, grid);
        if (b_mn) return launch_pair_variant<false, true, 1>(h, st, p, grid);
        return launch_pair_variant<false, false, 1>(h, st, p, grid);
    }
    if (a_mn) return launch_pair_variant<true, true>(h, st, p, grid);
    if (b_mn) return launch_pair_variant<false, true>(h, st, p, grid);
    return launch_pair_variant<false, false>(h, st, p, grid);
}

// Pairs pay off when the k-loop is long enough to amortise the cross-CTA handshakes and the 256 x 256 tiles keep most SMs busy as
// pairs (measured: the K = 512 GEMMs of the memory block run a few percent faster on the single-CTA kernel; its K = 2048 ones --
// conv1 forward, conv3 dgrad: 64 tiles -- 0.1 ms per step faster on pairs, which halve the B-operand traffic per CTA).
static bool use_pair_kernel(const grl_handle* h, int M, int N, int K, int batch) {
    if ((h->overlap & 16) != 0 || M <= GEMM_BM) return false;
    if ((h->overlap & 64) != 0) return true;                                              // debug bit 6: pairs for every 256-wide tile
    const long long tiles = (long long)((M + GP_BM - 1) / GP_BM) * ((N + GP_BN - 1) / GP_BN) * batch;
    return K >= 1024 && 2 * tiles >= (long long)(h->num_sms - h->num_sms / 7);            // >= 64 pairs on 148 SMs
}

int gemm_launch(grl_handle* h, cudaStream_t st, int M, int N, int K, int batch, const Operand& A, const Operand& B,
                GemmEpi epi, int bn) {
    if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) return set_error(h, GRL_EINVAL, "gemm: empty problem %dx%dx%d", M, N, K);
    if (A.mn_major && !B.mn_major) return set_error(h, GRL_EINVAL, "gemm: MN-major A with K-major B is not instantiated");
    if ((A.mn_major || B.mn_major) && (K % GEMM_BK)) return set_error(h, GRL_EINVAL, "gemm: MN-major operands need K %% 64 == 0");
    if (epi.bnb_mask && ((N % 32) || (M % GEMM_BM) || !epi.col_sum || !epi.col_sq || !epi.bnb_hraw || !epi.bnb_stat))
        return set_error(h, GRL_EINVAL, "gemm: the BatchNorm-backward epilogue needs N %% 32 == 0, M %% 128 == 0 and both partial-sum buffers");
    const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    if (bn == 0) {
        const long long t256 = (long long)m_tiles * ((N + 255) / 256) * batch;
        bn = (N > 128 && t256 >= (long long)h->num_sms * 3 / 4) ? 256 : 128;
    }
    if (bn != 128 && bn != 256) return set_error(h, GRL_EINVAL, "gemm: bn must be 0, 128 or 256");
    // 256-wide tiles go to the CTA-pair kernel (two SMs per 256 x 256 tile) unless grl_set_overlap bit 4 asks for the single-CTA one
    if (bn == 256 && use_pair_kernel(h, M, N, K, batch))
        return gemm_launch_pair(h, st, M, N, K, batch, A.hi, A.lo, A.ld, A.bstride, A.mn_major, B.hi, B.lo, B.ld, B.bstride, B.mn_major, epi);
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K; p.batch = batch;
    p.num_m_tiles = m_tiles;
    p.num_n_tiles = (N + bn - 1) / bn;
    p.group_m = p.num_m_tiles > 32 ? 16 : p.num_m_tiles;
    p.epi = epi;
    GRL_TRY(make_tmap(h, &p.ta_hi, A.hi, A.ld, A.bstride, A.mn_major, M, K, batch, GEMM_BM));
    GRL_TRY(make_tmap(h, &p.ta_lo, A.lo, A.ld, A.bstride, A.mn_major, M, K, batch, GEMM_BM));
    GRL_TRY(make_tmap(h, &p.tb_hi, B.hi, B.ld, B.bstride, B.mn_major, N, K, batch, bn));
    GRL_TRY(make_tmap(h, &p.tb_lo, B.lo, B.ld, B.bstride, B.mn_major, N, K, batch, bn));
    const long long tiles = (long long)p.num_m_tiles * p.num_n_tiles * batch;
    const int grid = (int)(tiles < h->num_sms ? tiles : h->num_sms);
    // K-major/K-major: fprop and the distance GEMM; MN/MN: wgrad; K-major A with MN-major B: dgrad
    // (weights [Cout][Cin] consumed as B[n = Cin][k = Cout] without a transposed copy).
    if (bn == 256) {
        if (A.mn_major) return launch_variant<256, true, true>(h, st, p, grid);
        // short k-loops (the memory block's K = 512 convolutions): the epilogue of a tile is as long as its main loop, so eight
        // epilogue warps instead of four (ncu: 28.1 -> 26.5 us conv2, 71 -> 68 us conv3 forward; results bit-identical)
        if (K <= 512) {
            if (B.mn_major) return launch_variant<256, false, true, 3, 8>(h, st, p, grid);
            return launch_variant<256, false, false, 3, 8>(h, st, p, grid);
        }
        if (B.mn_major) return launch_variant<256, false, true>(h, st, p, grid);
        return launch_variant<256, false, false>(h, st, p, grid);
    }
    if (A.mn_major) return launch_variant<128, true, true>(h, st, p, grid);
    if (B.mn_major) return launch_variant<128, false, true>(h, st, p, grid);
    return launch_variant<128, false, false>(h, st, p, grid);
}

// D[z] = A[z] * B[z]^T with ONE fp16 plane per operand and any operand majors (one MMA per k-step): the weight / input
// gradients of the TRL attention convs f1 / f2, whose single-pass fp16 form stays inside every gradient gate.
int gemm_launch_x1(grl_handle* h, cudaStream_t st, int M, int N, int K, int batch, const __half* A, long long lda, long long a_bstride,
                   int a_mn, const __half* B, long long ldb, long long b_bstride, int b_mn, GemmEpi epi) {
    if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) return set_error(h, GRL_EINVAL, "gemm_x1: empty problem %dx%dx%d", M, N, K);
    if (a_mn && !b_mn) return set_error(h, GRL_EINVAL, "gemm_x1: MN-major A with K-major B is not instantiated");
    if ((a_mn || b_mn) && (K % GEMM_BK)) return set_error(h, GRL_EINVAL, "gemm_x1: MN-major operands need K %% 64 == 0");
    const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    const long long t256 = (long long)m_tiles * ((N + 255) / 256) * batch;
    const int bn = (N > 128 && t256 >= (long long)h->num_sms * 3 / 4) ? 256 : 128;
    if (bn == 256 && (h->overlap & 32) == 0 && use_pair_kernel(h, M, N, K, batch))      // debug bit 5: single-CTA kernel for the fp16 GEMMs
        return gemm_launch_pair(h, st, M, N, K, batch, A, nullptr, lda, a_bstride, a_mn, B, nullptr, ldb, b_bstride, b_mn, epi);
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K; p.batch = batch;
    p.num_m_tiles = m_tiles;
    p.num_n_tiles = (N + bn - 1) / bn;
    p.group_m = p.num_m_tiles > 32 ? 16 : p.num_m_tiles;
    p.epi = epi;
    GRL_TRY(make_tmap(h, &p.ta_hi, A, lda, a_bstride, a_mn, M, K, batch, GEMM_BM, CU_TENSOR_MAP_DATA_TYPE_FLOAT16));
    GRL_TRY(make_tmap(h, &p.tb_hi, B, ldb, b_bstride, b_mn, N, K, batch, bn, CU_TENSOR_MAP_DATA_TYPE_FLOAT16));
    p.ta_lo = p.ta_hi; p.tb_lo = p.tb_hi;
    const long long tiles = (long long)p.num_m_tiles * p.num_n_tiles * batch;
    const int grid = (int)(tiles < h->num_sms ? tiles : h->num_sms);
    if (bn == 256) {
        if (a_mn) return launch_variant<256, true, true, 1>(h, st, p, grid);
        if (b_mn) return launch_variant<256, false, true, 1>(h, st, p, grid);
        return launch_variant<256, false, false, 1>(h, st, p, grid);
    }
    if (a_mn) return launch_variant<128, true, true, 1>(h, st, p, grid);
    if (b_mn) return launch_variant<128, false, true, 1>(h, st, p, grid);
    return launch_variant<128, false, false, 1>(h, st, p, grid);
}

// D = A * B^T with ONE fp16 plane per operand (K-major both, no batch): the coarse pass of the retrieval search.
int gemm_launch_f16(grl_handle* h, cudaStream_t st, int M, int N, int K, const __half* A, long long lda, const __half* B,
                    long long ldb, GemmEpi epi, int bn) {
    if (M <= 0 || N <= 0 || K <= 0) return set_error(h, GRL_EINVAL, "gemm_f16: empty problem %dx%dx%d", M, N, K);
    const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    if (bn == 0) {
        const long long t256 = (long long)m_tiles * ((N + 255) / 256);
        bn = (N > 128 && t256 >= (long long)h->num_sms * 3 / 4) ? 256 : 128;
    }
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K; p.batch = 1;
    p.num_m_tiles = m_tiles;
    p.num_n_tiles = (N + bn - 1) / bn;
    p.group_m = p.num_m_tiles > 32 ? 16 : p.num_m_tiles;
    p.epi = epi;
    GRL_TRY(make_tmap(h, &p.ta_hi, A, lda, 0, 0, M, K, 1, GEMM_BM, CU_TENSOR_MAP_DATA_TYPE_FLOAT16));
    GRL_TRY(make_tmap(h, &p.tb_hi, B, ldb, 0, 0, N, K, 1, bn, CU_TENSOR_MAP_DATA_TYPE_FLOAT16));
    p.ta_lo = p.ta_hi; p.tb_lo = p.tb_hi;
    const long long tiles = (long long)p.num_m_tiles * p.num_n_tiles;
    const int grid = (int)(tiles < h->num_sms ? tiles : h->num_sms);
    if (bn == 256) return launch_variant<256, false, false, 1>(h, st, p, grid);
    return launch_variant<128, false, false, 1>(h, st, p, grid);
}

// The 256 x 256-tile coarse GEMM (coarse_gemm.cuh); falls back to the 128-row single-plane variant for small problems.
int coarse_gemm_launch(grl_handle* h, cudaStream_t st, int M, int N, int K, const __half* A, long long lda, const __half* B,
                       long long ldb, GemmEpi epi) {
    const bool aligned = !epi.C || ((reinterpret_cast<uintptr_t>(epi.C) & 15) == 0 && (epi.ldc & 3) == 0);
    const bool aligned_cols = (!epi.col_scale || (reinterpret_cast<uintptr_t>(epi.col_scale) & 15) == 0) &&
                              (!epi.col_norm || (reinterpret_cast<uintptr_t>(epi.col_norm) & 15) == 0);
    // (the first chunks of a search are 512 columns wide: still 2 column tiles x M/256 row tiles for the 256 x 256 kernel)
    if (M < 1024 || N < 512 || !aligned || !aligned_cols) return gemm_launch_f16(h, st, M, N, K, A, lda, B, ldb, epi, 0);
    CoarseGemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K;
    p.num_m_tiles = (M + CG_BM - 1) / CG_BM;
    p.num_n_tiles = (N + CG_BN - 1) / CG_BN;
    p.group_m = p.num_m_tiles > 16 ? 8 : p.num_m_tiles;
    p.epi = epi;
    const bool pair_boxes = (h->overlap & 8) == 0;   // the CTA-pair kernel loads 128-row boxes per CTA
    GRL_TRY(make_tmap(h, &p.ta, A, lda, 0, 0, M, K, 1, pair_boxes ? 128 : CG_BM, CU_TENSOR_MAP_DATA_TYPE_FLOAT16));
    GRL_TRY(make_tmap(h, &p.tb, B, ldb, 0, 0, N, K, 1, pair_boxes ? 128 : CG_BN, CU_TENSOR_MAP_DATA_TYPE_FLOAT16));
    const bool pairs = (h->overlap & 8) == 0;        // CTA-pair kernel (cta_group::2) unless grl_set_overlap bit 3 asks for the single-CTA one
    if (pairs) GRL_TRY(ensure_dyn_smem(h, (const void*)coarse_gemm2_kernel, C2_SMEM_BYTES));
    else GRL_TRY(ensure_dyn_smem(h, (const void*)coarse_gemm_kernel, CG_SMEM_BYTES));
    const long long tiles = (long long)p.num_m_tiles * p.num_n_tiles;
    int grid = (int)(tiles < h->num_sms ? tiles : h->num_sms);
    if (pairs) {
        long long g2 = 2 * tiles < (long long)(h->num_sms & ~1) ? 2 * tiles : (long long)(h->num_sms & ~1);
        grid = (int)g2;
    }
    grl_prof_rec rec;
    if (h->prof_on) {
        GRL_CUDA(h, cudaEventCreate(&rec.e0));
        GRL_CUDA(h, cudaEventCreate(&rec.e1));
        rec.flops = 2.0 * M * (double)N * K;
        GRL_CUDA(h, cudaEventRecord(rec.e0, st));
    }
    if (pairs) coarse_gemm2_kernel<<<grid, C2_THREADS, C2_SMEM_BYTES, st>>>(p);
    else coarse_gemm_kernel<<<grid, CG_THREADS, CG_SMEM_BYTES, st>>>(p);
    GRL_LAUNCH_CHECK(h);
    if (h->prof_on) {
        GRL_CUDA(h, cudaEventRecord(rec.e1, st));
        h->prof->recs.push_back(rec);
    }
    return GRL_OK;
}

}  // namespace grl

// ------------------------------------------------------------------ C ABI: lifecycle
using namespace grl;

extern "C" const char* grl_version(void) { return "grl_b200 0.1 (sm_100a)"; }

extern "C" int grl_create(int device, grl_handle** out) {
    if (!out) return set_error(nullptr, GRL_EINVAL, "grl_create: out is NULL");
    *out = nullptr;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return set_error(nullptr, GRL_ECUDA, "cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
    if (prop.major != 10) return set_error(nullptr, GRL_EARCH, "device %d is sm_%d%d; grl_b200 needs sm_100", device, prop.major, prop.minor);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error(nullptr, GRL_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    grl_handle* h = new grl_handle();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        delete h;
        return set_error(nullptr, GRL_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    }
    h->encode = reinterpret_cast<grl_encode_tiled_fn>(fn);
    h->prof = new grl_prof();
    h->func_attrs = new grl_func_attrs();
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);          // numerically: lo >= hi; `lo` is the least urgent
    e = cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_lo);
    if (e != cudaSuccess) {
        delete h->prof; delete static_cast<grl_func_attrs*>(h->func_attrs); delete h;
        return set_error(nullptr, GRL_ECUDA, "cudaStreamCreateWithPriority: %s", cudaGetErrorString(e));
    }
    h->overlap = 3;
    *out = h;
    return GRL_OK;
}

extern "C" void grl_destroy(grl_handle* h) {
    if (!h) return;
    if (h->prof) {
        for (auto& r : h->prof->recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        delete h->prof;
    }
    for (int i = 0; i < h->n_events; ++i) cudaEventDestroy(h->events[i]);
    delete[] h->events;
    if (h->side) cudaStreamDestroy(h->side);
    grl::comm_release(h);
    for (int i = 0; i < h->n_stage_ev; ++i) cudaEventDestroy(h->stage_ev[i]);
    delete[] h->stage_ev;
    delete static_cast<grl_func_attrs*>(h->func_attrs);
    delete h;
}

extern "C" int grl_set_overlap(grl_handle* h, int on) {
    if (!h) return GRL_EINVAL;
    h->overlap = on & 255;
    return GRL_OK;
}

extern "C" int grl_profile_enable(grl_handle* h, int on) {
    if (!h) return GRL_EINVAL;
    h->prof_on = on ? 1 : 0;
    return GRL_OK;
}

extern "C" int grl_profile_read(grl_handle* h, double* gemm_ms, double* gemm_flops, long long* gemm_launches) {
    if (!h || !gemm_ms || !gemm_flops || !gemm_launches) return set_error(h, GRL_EINVAL, "grl_profile_read: NULL argument");
    double ms = 0.0, fl = 0.0;
    long long n = 0;
    for (auto& r : h->prof->recs) {
        GRL_CUDA(h, cudaEventSynchronize(r.e1));
        float t = 0.f;
        GRL_CUDA(h, cudaEventElapsedTime(&t, r.e0, r.e1));
        ms += t; fl += r.flops; ++n;
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    h->prof->recs.clear();
    *gemm_ms = ms; *gemm_flops = fl; *gemm_launches = n;
    return GRL_OK;
}

extern "C" const char* grl_last_error(const grl_handle* h) {
    if (h) return h->err;
    static thread_local char none[8] = "";
    return none;
}
extern "C" int grl_num_sms(const grl_handle* h) { return h ? h->num_sms : 0; }
extern "C" long long grl_launch_count(const grl_handle* h) { return h ? h->launches : 0; }

// ------------------------------------------------------------------ C ABI: GEMM primitive
static void gemm_ws_layout(const grl_gemm_desc* d, size_t* a_elems, size_t* b_elems) {
    const size_t a_rows = d->a_mn_major ? d->K : d->M, b_rows = d->b_mn_major ? d->K : d->N;
    *a_elems = (size_t)d->batch * a_rows * (size_t)d->lda;
    *b_elems = (size_t)d->batch * b_rows * (size_t)d->ldb;
}

extern "C" size_t grl_gemm_workspace_bytes(const grl_gemm_desc* d) {
    size_t a, b;
    gemm_ws_layout(d, &a, &b);
    return 2 * (align_up(a * 2, 1024) + align_up(b * 2, 1024));
}

extern "C" int grl_gemm_bf16x3(grl_handle* h, const grl_gemm_desc* d, const float* A, const float* B, float* C,
                               void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !d || !A || !B || !workspace) return set_error(h, GRL_EINVAL, "grl_gemm_bf16x3: NULL argument");
    if (workspace_bytes < grl_gemm_workspace_bytes(d)) return set_error(h, GRL_ENOMEM, "grl_gemm_bf16x3: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    size_t a_el, b_el;
    gemm_ws_layout(d, &a_el, &b_el);
    uint8_t* w = (uint8_t*)workspace;
    __nv_bfloat16* a_hi = (__nv_bfloat16*)w; w += align_up(a_el * 2, 1024);
    __nv_bfloat16* a_lo = (__nv_bfloat16*)w; w += align_up(a_el * 2, 1024);
    __nv_bfloat16* b_hi = (__nv_bfloat16*)w; w += align_up(b_el * 2, 1024);
    __nv_bfloat16* b_lo = (__nv_bfloat16*)w;
    // batches are assumed densely stacked (bstride == rows*ld) in this debug entry point
    const long long a_rows = d->a_mn_major ? d->K : d->M, b_rows = d->b_mn_major ? d->K : d->N;
    const int a_cols = d->a_mn_major ? d->M : d->K, b_cols = d->b_mn_major ? d->N : d->K;
    GRL_TRY(split_planes(h, st, A, d->lda, a_hi, a_lo, d->lda, a_rows * d->batch, a_cols));
    GRL_TRY(split_planes(h, st, B, d->ldb, b_hi, b_lo, d->ldb, b_rows * d->batch, b_cols));
    Operand oa{a_hi, a_lo, d->lda, a_rows * d->lda, d->a_mn_major};
    Operand ob{b_hi, b_lo, d->ldb, b_rows * d->ldb, d->b_mn_major};
    GemmEpi e = epi_default();
    e.C = C; e.ldc = d->ldc; e.c_bstride = d->c_bstride;
    e.alpha = d->alpha;
    e.row_scale = d->row_scale; e.rs_bstride = d->M;
    e.col_bias = d->col_bias; e.cb_bstride = d->N;
    e.relu = d->relu; e.accumulate = d->accumulate;
    e.col_sum = d->col_sum; e.col_sq = d->col_sq;
    e.stat_bstride = (long long)4 * ((d->M + GEMM_BM - 1) / GEMM_BM) * d->N;
    e.Phi = (__nv_bfloat16*)d->planes_hi; e.Plo = (__nv_bfloat16*)d->planes_lo;
    e.ldp = d->N; e.p_bstride = (long long)d->M * d->N;
    return gemm_launch(h, st, d->M, d->N, d->K, d->batch, oa, ob, e, d->bn);
}
