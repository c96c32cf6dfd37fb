// grl_b200 — split-bf16 ("bf16x3") tcgen05 GEMM for sm_100a.
//
//   D[z][m][n] = sum_k A[z][m][k] * B[z][n][k]        (fp32 accumulate in TMEM)
//
// A and B are given as two bf16 planes each (hi, lo) with x ~= hi + lo (16 mantissa
// bits).  Three MMAs per k-step (hi*hi + lo*hi + hi*lo) give ~2e-5 relative error
// per contraction, which is what GRL's 1e-3 parity bar on a >=30-deep GEMM chain
// needs; single-pass bf16/tf32 does not.
//
// Structure (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0   : TMA producer   — cp.async.bulk.tensor into a ring of smem stages,
//              hi planes and lo planes complete on separate mbarriers so the
//              hi*hi MMAs start as soon as half a stage has landed
//   warp 1   : MMA issuer     — one lane issues tcgen05.mma (UMMA 128 x BN x 16),
//              tcgen05.commit frees smem stages / publishes the accumulator
//   warps 2-5: epilogue       — tcgen05.ld 32x32b.x32 (thread == output row),
//              fused bias / row-scale / ReLU / L2-distance / (v - sub)^2 column
//              reductions / BatchNorm partial statistics / bf16 hi-lo re-split,
//              overlapped with the next tile's MMAs via 2 TMEM accumulator stages
//
// Operand majors: K-major (rows x K, K contiguous) for fprop/dgrad; MN-major
// (K x rows, rows contiguous) for wgrad, so activations/gradients are consumed
// in their pixel-major layout without a transposed copy.  Both use SWIZZLE_128B.
#pragma once
#include "common.cuh"

namespace grl {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

struct GemmEpi {
    // ---- outputs (all optional) ----
    float* C;                 // fp32 [z][M][ldc]
    long long ldc, c_bstride;
    __nv_bfloat16* Phi;       // bf16 hi/lo planes of the stored value [z][M][ldp]
    __nv_bfloat16* Plo;
    long long ldp, p_bstride;
    float* col_sum;           // [z][4*num_m_tiles][N] per-(tile,warp) column sums of the "stat value"
    float* col_sq;            // same, squares
    long long stat_bstride;
    // ---- transforms: v = alpha*acc; v *= row_scale[m]; v += col_bias[n] + grp_bias[m/grp_rows][n]; relu ----
    float alpha;
    const float* dscale_a;    // device scalars (or NULL) multiplied into alpha: the de-scaling factors of fp16 operands whose
    const float* dscale_b;    // power-of-two scale was chosen on the device (gemm_launch_x1)
    const float* row_scale;
    long long rs_bstride;
    const float* col_bias;
    long long cb_bstride;
    const float* col_scale;   // [N] or NULL: v *= col_scale[n] before everything else (per-row power-of-two scales of fp16 operands)
    const float* grp_bias;
    int grp_rows;
    long long ld_gb;
    int relu;
    int accumulate;           // C += v (read-modify-write)
    int mode;                 // 0 plain, 1 L2: v = sqrt(max(row_norm[m] + col_norm[n] - 2*acc, 1e-12)), 2: the same without the sqrt
    const float* row_norm;
    const float* col_norm;
    // ---- top-k candidate filter (retrieval): every value v <= tk_thresh[m] is appended, as a (distance, global column) key, to
    //      row m's candidate list (capacity tk_cap; tk_cnt keeps counting past it so the consumer sees the overflow) ----
    unsigned long long* tk_cand;
    int* tk_cnt;
    const float* tk_thresh;
    int tk_cap;
    long long tk_idx_base;
    // ---- "sub" mode (TRL f1): e = v - sub[sub_row][sub_col]; stat value = e (else v) ----
    const float* sub;
    long long ld_sub;
    long long sub_tile_rows;  // sub_row = m_tile*sub_tile_rows + sub_row_off[z] + row_in_tile
    long long sub_row_off[2];
    long long sub_col_off[2];
    // ---- "BatchNorm backward" statistics of the stored tile (the dgrad GEMM that PRODUCES the gradient of a BN + ReLU output):
    //      g = v where the ReLU output bnb_mask > 0 else 0, xhat = (bnb_hraw - mean[n]) * rstd[n] (bnb_stat: [z][4][N], mean in
    //      row 2, rstd in row 3); col_sum receives the column sums of g, col_sq those of g * xhat (same [z][4*num_m_tiles][N]
    //      partial layout), bnb_scal[0] / [1] the launch-wide max |g| / max |xhat| as float bits (atomicMax; may be NULL) ----
    const __nv_bfloat16* bnb_mask;    // [z][M][bnb_ld]
    const float* bnb_hraw;            // [z][M][bnb_ld]
    const float* bnb_stat;
    long long bnb_ld, bnb_bstride;
    unsigned int* bnb_scal;
};

struct GemmParams {
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    int M, N, K, batch;
    int num_m_tiles, num_n_tiles;
    int group_m;              // tile rasterisation: groups of `group_m` m-tiles sweep all n-tiles before the next group
    GemmEpi epi;
};

// Tile index -> (batch z, m tile, n tile).  Within a batch, groups of group_m m-tiles are swept across all n-tiles
// (m fastest inside a group) before moving on, so the CTAs running at any time share a slab of A rows and a band of B
// rows that fit L2: each operand is read from HBM about once even when M*K does not fit L2 (P = B*T*128 pixel rows).
__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int tiles_per_batch, int& z, int& m_tile, int& n_tile) {
    z = tile / tiles_per_batch;
    const int r = tile - z * tiles_per_batch;
    const int per_group = p.group_m * p.num_n_tiles;
    const int mg = r / per_group;
    const int rr = r - mg * per_group;
    const int gsize = min(p.group_m, p.num_m_tiles - mg * p.group_m);
    n_tile = rr / gsize;
    m_tile = mg * p.group_m + (rr - n_tile * gsize);
}

// Per-row state of one epilogue thread for one output tile (thread == output row): loaded BEFORE the accumulator is waited for.
struct EpiRow {
    int grow;
    bool row_ok;
    float rscale, rnorm, rthresh;
    const float* gb_row;
    const float* sub_row;
    float* c_row;
};
// m_tile counts 128-row tiles (the unit of the statistic / `sub` layouts), row_in_tile = TMEM lane of this thread.
__device__ __forceinline__ EpiRow epi_row_setup(const GemmParams& p, int z, int m_tile, int row_in_tile) {
    const GemmEpi& e = p.epi;
    EpiRow r;
    r.grow = m_tile * GEMM_BM + row_in_tile;
    r.row_ok = r.grow < p.M;
    r.rscale = 1.f; r.rnorm = 0.f; r.rthresh = 0.f;
    if (r.row_ok) {
        if (e.row_scale) r.rscale = e.row_scale[z * e.rs_bstride + r.grow];
        if (e.mode != 0) r.rnorm = e.row_norm[r.grow];
        if (e.tk_cand)      // a row whose list already overflowed is rescanned by the consumer anyway: stop appending
            r.rthresh = (e.tk_cnt[r.grow] > e.tk_cap) ? -__int_as_float(0x7f800000) : e.tk_thresh[r.grow];
    }
    if (e.dscale_a) r.rscale *= __ldg(e.dscale_a);
    if (e.dscale_b) r.rscale *= __ldg(e.dscale_b);
    r.gb_row = e.grp_bias ? e.grp_bias + (long long)(r.grow / e.grp_rows) * e.ld_gb : nullptr;
    r.sub_row = nullptr;
    if (e.sub) r.sub_row = e.sub + ((long long)m_tile * e.sub_tile_rows + e.sub_row_off[z] + row_in_tile) * e.ld_sub + e.sub_col_off[z];
    r.c_row = e.C ? e.C + z * e.c_bstride + (long long)r.grow * e.ldc : nullptr;
    return r;
}

// Drains one 128 x BN accumulator (TMEM address t_acc: lane quadrant and first column included) through the fused epilogue.
template <int BN>
__device__ __forceinline__ void epi_tile(const GemmParams& p, const EpiRow& R, int z, int m_tile, int n0, uint32_t t_acc, int quad, int lane,
                                         int c_begin = 0, int c_end = BN / 32) {
    const GemmEpi& e = p.epi;
    const int grow = R.grow;
    const bool row_ok = R.row_ok;
    const float rscale = R.rscale, rnorm = R.rnorm, rthresh = R.rthresh;
    const float* gb_row = R.gb_row;
    const float* sub_row = R.sub_row;
    float* c_row = R.c_row;
    float bnb_gmax = 0.f, bnb_xmax = 0.f;
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {       // (the 32-column chunks of the tile this warp drains: all of them, or one half)
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;                   // warp-uniform
        // BatchNorm-backward statistics: the ReLU outputs and raw activations of this chunk are requested first (their latency
        // hides behind the accumulator read and the stores of the tile)
        uint4 bnb_mk[4];
        float4 bnb_hv[8];
        if (e.bnb_mask && row_ok) {
            const long long off = z * e.bnb_bstride + (long long)grow * e.bnb_ld + col0;
            const uint4* m4 = reinterpret_cast<const uint4*>(e.bnb_mask + off);
            const float4* h4 = reinterpret_cast<const float4*>(e.bnb_hraw + off);
#pragma unroll
            for (int j = 0; j < 4; ++j) bnb_mk[j] = __ldg(m4 + j);
#pragma unroll
            for (int j = 0; j < 8; ++j) bnb_hv[j] = __ldg(h4 + j);
        }
        float v[32];
        tmem_ld_32x32(t_acc + uint32_t(c * 32), v);
        tmem_ld_wait();
        const bool full_chunk = col0 + 32 <= p.N;
        if (e.col_scale) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= (full_chunk || col0 + j < p.N) ? __ldg(e.col_scale + col0 + j) : 0.f;
        }
        if (e.mode != 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float cn = (col0 + j < p.N) ? e.col_norm[col0 + j] : 0.f;
                const float sq = fmaxf(rnorm + cn - 2.f * (v[j] * rscale), 1e-12f);
                v[j] = (e.mode == 1) ? sqrtf(sq) : sq;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= e.alpha * rscale;
            if (e.col_bias) {
                const float* cb = e.col_bias + z * e.cb_bstride + col0;
#pragma unroll
                for (int j = 0; j < 32; ++j) if (full_chunk || col0 + j < p.N) v[j] += __ldg(cb + j);
            }
            if (gb_row && row_ok) {
#pragma unroll
                for (int j = 0; j < 32; ++j) if (full_chunk || col0 + j < p.N) v[j] += __ldg(gb_row + col0 + j);
            }
            if (e.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
        }
        // ---- top-k candidates (small problems only; the large-problem kernel is coarse_gemm.cuh) ----
        if (e.tk_cand && row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (v[j] <= rthresh && (full_chunk || col0 + j < p.N)) {
                    const int pos = atomicAdd(e.tk_cnt + grow, 1);
                    if (pos < e.tk_cap)
                        e.tk_cand[(long long)grow * e.tk_cap + pos] = make_key(v[j], (uint32_t)(e.tk_idx_base + col0 + j));
                }
            }
        }
        // ---- stores of v ----
        if (row_ok) {
            if (c_row) {
                float* dst = c_row + col0;
                if (full_chunk && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                    float4* d4 = reinterpret_cast<float4*>(dst);
                    if (e.accumulate) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 o = d4[j];
                            v[4 * j] += o.x; v[4 * j + 1] += o.y; v[4 * j + 2] += o.z; v[4 * j + 3] += o.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < p.N) {
                            if (e.accumulate) v[j] += dst[j];
                            dst[j] = v[j];
                        }
                }
            }
            if (e.Phi && full_chunk) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(v[2 * j], h0, l0); split_bf16(v[2 * j + 1], h1, l1);
                    hi[j] = pack_bf16(h0, h1); lo[j] = pack_bf16(l0, l1);
                }
                const long long off = z * e.p_bstride + (long long)grow * e.ldp + col0;
                uint4* ph = reinterpret_cast<uint4*>(e.Phi + off);
                uint4* pl = reinterpret_cast<uint4*>(e.Plo + off);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    ph[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                    pl[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                }
            }
        }
        // ---- BatchNorm-backward partial sums of the stored gradient tile ----
        if (e.bnb_mask) {
            float w[32];
            if (row_ok) {
                const float* mean = e.bnb_stat + (long long)z * 4 * p.N + 2 * p.N + col0;
                const float* rstd = mean + p.N;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 mk = bnb_mk[j];
                    const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float4 hv = bnb_hv[2 * j + q];
                        const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int jj = 8 * j + 4 * q + i;
                            const uint32_t bits = (mw[2 * q + (i >> 1)] >> ((i & 1) * 16)) & 0xFFFFu;     // bf16 > 0: 0x0001 .. 0x7F80
                            const float gv = (bits - 1u < 0x7F80u) ? v[jj] : 0.f;
                            const float xh = (hh[i] - __ldg(mean + jj)) * __ldg(rstd + jj);
                            v[jj] = gv; w[jj] = gv * xh;
                            bnb_gmax = fmaxf(bnb_gmax, fabsf(gv)); bnb_xmax = fmaxf(bnb_xmax, fabsf(xh));
                        }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) { v[j] = 0.f; w[j] = 0.f; }
            }
            const long long sidx = z * e.stat_bstride + (long long)(m_tile * 4 + quad) * p.N + col0 + lane;
            const float s2 = warp_transpose_sum32(w);
            e.col_sq[sidx] = s2;
            const float s1 = warp_transpose_sum32(v);
            e.col_sum[sidx] = s1;
            continue;
        }
        // ---- column reductions (BN statistics / squared-difference pooling) ----
        if (e.col_sum || e.col_sq) {
            if (sub_row) {
                const float4* s4 = reinterpret_cast<const float4*>(sub_row + col0);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 s = __ldg(s4 + j);
                    v[4 * j] -= s.x; v[4 * j + 1] -= s.y; v[4 * j + 2] -= s.z; v[4 * j + 3] -= s.w;
                }
            }
            if (!row_ok) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            const long long sidx = z * e.stat_bstride + (long long)(m_tile * 4 + quad) * p.N + col0 + lane;
            if (e.col_sq) {
                float w[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) w[j] = v[j] * v[j];
                const float s2 = warp_transpose_sum32(w);
                if (col0 + lane < p.N) e.col_sq[sidx] = s2;
            }
            if (e.col_sum) {
                const float s1 = warp_transpose_sum32(v);
                if (col0 + lane < p.N) e.col_sum[sidx] = s1;
            }
        }
    }
    if (e.bnb_mask && e.bnb_scal) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            bnb_gmax = fmaxf(bnb_gmax, __shfl_xor_sync(0xffffffffu, bnb_gmax, off));
            bnb_xmax = fmaxf(bnb_xmax, __shfl_xor_sync(0xffffffffu, bnb_xmax, off));
        }
        if (lane == 0) {
            if (bnb_gmax == bnb_gmax && bnb_gmax > 0.f) atomicMax(e.bnb_scal + 0, __float_as_uint(bnb_gmax));
            if (bnb_xmax == bnb_xmax && bnb_xmax > 0.f) atomicMax(e.bnb_scal + 1, __float_as_uint(bnb_xmax));
        }
    }
}

// PLANES = 3: split-bf16 (hi and lo planes, three MMAs per k-step).  PLANES = 1: one fp16 plane per operand, one MMA per
// k-step -- the coarse pass of the retrieval search (eval.cu), whose candidates are re-scored exactly afterwards.
template <int BN, int PLANES = 3>
struct GemmCfg {
    static constexpr int STAGES = (PLANES == 1) ? ((BN == 256) ? 4 : 6) : ((BN == 256) ? 2 : 3);
    static constexpr int A_PLANE = GEMM_BM * GEMM_BK * 2;   // bytes per A plane per stage (16 KB)
    static constexpr int B_PLANE = BN * GEMM_BK * 2;        // 16 / 32 KB
    static constexpr int STAGE_BYTES = (PLANES == 1 ? 1 : 2) * (A_PLANE + B_PLANE);
    static constexpr int TMEM_COLS = 2 * BN;                // two accumulator stages
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// EW = 4 or 8 epilogue warps.  A warp may only read its own TMEM lane quadrant (warp % 4), so with eight the two warps of a
// quadrant split the tile's columns: two warps per scheduler instead of one for the latency-bound epilogue of the short-K GEMMs.
template <int BN, bool A_MN, bool B_MN, int PLANES = 3, int EW = 4>
__global__ void __launch_bounds__(64 + 32 * EW, 1) gemm_bf16x3_kernel(const __grid_constant__ GemmParams p) {
    static_assert(EW == 4 || EW == 8, "four or eight epilogue warps");
    using Cfg = GemmCfg<BN, PLANES>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_hi = bars;                    // [STAGES]
    uint64_t* full_lo = bars + STAGES;           // [STAGES]
    uint64_t* empty = bars + 2 * STAGES;         // [STAGES]
    uint64_t* tmem_full = bars + 3 * STAGES;     // [2]
    uint64_t* tmem_empty = bars + 3 * STAGES + 2;  // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.ta_hi); tma_prefetch_desc(&p.ta_lo);
        tma_prefetch_desc(&p.tb_hi); tma_prefetch_desc(&p.tb_lo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_hi[s], 1); mbar_init(&full_lo[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], EW); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
    const int tiles_per_batch = p.num_m_tiles * p.num_n_tiles;
    const int num_tiles = tiles_per_batch * p.batch;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int z, m_tile, n_tile;
                tile_coords(p, tile, tiles_per_batch, z, m_tile, n_tile);
                const int m0 = m_tile * GEMM_BM;
                const int n0 = n_tile * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int k0 = kb * GEMM_BK;
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sA_hi = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sB_hi = sA_hi + Cfg::A_PLANE;
                    uint8_t* sA_lo = sB_hi + Cfg::B_PLANE;
                    uint8_t* sB_lo = sA_lo + Cfg::A_PLANE;
                    mbar_arrive_expect_tx(&full_hi[stage], Cfg::A_PLANE + Cfg::B_PLANE);
                    if (A_MN) {
#pragma unroll
                        for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_3d(sA_hi + j * 8192, &p.ta_hi, &full_hi[stage], m0 + 64 * j, k0, z);
                    } else {
                        tma_load_3d(sA_hi, &p.ta_hi, &full_hi[stage], k0, m0, z);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) tma_load_3d(sB_hi + j * 8192, &p.tb_hi, &full_hi[stage], n0 + 64 * j, k0, z);
                    } else {
                        tma_load_3d(sB_hi, &p.tb_hi, &full_hi[stage], k0, n0, z);
                    }
                    if (PLANES == 1) {
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full_lo[stage], Cfg::A_PLANE + Cfg::B_PLANE);
                    if (A_MN) {
#pragma unroll
                        for (int j = 0; j < GEMM_BM / 64; ++j) tma_load_3d(sA_lo + j * 8192, &p.ta_lo, &full_lo[stage], m0 + 64 * j, k0, z);
                    } else {
                        tma_load_3d(sA_lo, &p.ta_lo, &full_lo[stage], k0, m0, z);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) tma_load_3d(sB_lo + j * 8192, &p.tb_lo, &full_lo[stage], n0 + 64 * j, k0, z);
                    } else {
                        tma_load_3d(sB_lo, &p.tb_lo, &full_lo[stage], k0, n0, z);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = (PLANES == 1) ? make_idesc_f16(GEMM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0)
                                                     : make_idesc_bf16(GEMM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            // K-major  SW128: 8-row groups 1024 B apart (SBO); LBO unused.   k-step (16 elems) = +32 B
            // MN-major SW128: 64-element MN atoms BK*128 B apart (LBO), 8-row k groups 1024 B apart (SBO); k-step = 16 rows = +2048 B
            constexpr uint32_t a_lbo = A_MN ? GEMM_BK * 128 : 16, b_lbo = B_MN ? GEMM_BK * 128 : 16;
            constexpr uint32_t a_kstep = A_MN ? (2048 >> 4) : (32 >> 4), b_kstep = B_MN ? (2048 >> 4) : (32 >> 4);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const uint32_t sA_hi = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sB_hi = sA_hi + Cfg::A_PLANE;
                    const uint32_t sA_lo = sB_hi + Cfg::B_PLANE;
                    const uint32_t sB_lo = sA_lo + Cfg::A_PLANE;
                    const uint64_t dA_hi = make_smem_desc(sA_hi, a_lbo, 1024), dA_lo = make_smem_desc(sA_lo, a_lbo, 1024);
                    const uint64_t dB_hi = make_smem_desc(sB_hi, b_lbo, 1024), dB_lo = make_smem_desc(sB_lo, b_lbo, 1024);
                    mbar_wait(&full_hi[stage], phase);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k)
                        umma_bf16(d_tmem, dA_hi + k * a_kstep, dB_hi + k * b_kstep, idesc, (kb | k) ? 1u : 0u);
                    if (PLANES == 3) {
                        mbar_wait(&full_lo[stage], phase);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) {
                            umma_bf16(d_tmem, dA_lo + k * a_kstep, dB_hi + k * b_kstep, idesc, 1u);
                            umma_bf16(d_tmem, dA_hi + k * a_kstep, dB_lo + k * b_kstep, idesc, 1u);
                        }
                    }
                    umma_commit(&empty[stage]);                 // smem stage reusable once these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const GemmEpi& e = p.epi;
        const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
        const int row_in_tile = quad * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            int z, m_tile, n_tile;
            tile_coords(p, tile, tiles_per_batch, z, m_tile, n_tile);
            const int m0 = m_tile * GEMM_BM;
            const int n0 = n_tile * BN;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const EpiRow R = epi_row_setup(p, z, m_tile, row_in_tile);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            constexpr int CH = BN / 32 / (EW / 4);            // chunks per epilogue warp
            const int c_begin = (EW == 8) ? ((warp - 2) >> 2) * CH : 0;
            epi_tile<BN>(p, R, z, m_tile, n0, tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * BN), quad, lane, c_begin, c_begin + CH);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace grl
