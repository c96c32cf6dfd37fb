// grl_b200 — coarse-pass GEMM of the retrieval search (sm_100a): D = A * B^T, one fp16 plane per operand.
//
// With a single MMA per k-step the 128 x 256 tile of gemm.cuh is bound by L2 -> shared-memory traffic (48 KB per 4.2 MFLOP,
// 85 FLOP/B against ~6300 B/clk of L2), not by the tensor pipe.  This kernel computes a 256 x 256 tile per CTA instead: one
// 256-row A box and one 256-row B box per k-block (64 KB per 8.4 MFLOP, 128 FLOP/B), two UMMA 128 x 256 x 16 per k-step into
// the two halves of TMEM (512 columns = the whole tensor memory, so the accumulator is single-buffered and the epilogue is
// spread over EIGHT warps, two per TMEM lane quadrant, to keep its exposed time short).
//   warp 0    TMA producer (3 stages of 64 KB)          warp 1    MMA issuer, TMEM owner
//   warps 2-9 epilogue: scale by the per-row/column power-of-two factors, optional squared-L2 form, top-k candidate filter
//             (the same contract as GemmEpi::tk_* in gemm.cuh), fp32 tile store
#pragma once
#include "gemm.cuh"

namespace grl {

constexpr int CG_BM = 256, CG_BN = 256, CG_BK = 64, CG_STAGES = 3;
constexpr int CG_THREADS = 320;
constexpr int CG_A_BYTES = CG_BM * CG_BK * 2, CG_B_BYTES = CG_BN * CG_BK * 2;
constexpr int CG_STAGE_BYTES = CG_A_BYTES + CG_B_BYTES;
constexpr int CG_COLBUF_BYTES = 2 * 2 * CG_BN * 4;   // double-buffered column scales and norms of the current tile
constexpr int CG_SMEM_BYTES = CG_STAGES * CG_STAGE_BYTES + 1024 + 256 + CG_COLBUF_BYTES;

struct CoarseGemmParams {
    CUtensorMap ta, tb;
    int M, N, K;
    int num_m_tiles, num_n_tiles, group_m;
    GemmEpi epi;            // uses C/ldc, alpha, row_scale, col_scale, mode (0 | 2), row_norm, col_norm, tk_*
};

// v[j] for a run-time j without indexing the register array dynamically (31 selects)
__device__ __forceinline__ float select32(const float (&v)[32], int j) {
    float a[16], b[8], c[4], d[2];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
    for (int i = 0; i < 2; ++i) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
    return (j & 16) ? d[1] : d[0];
}

// Epilogue of one warp's 32 rows x (<= 256) columns of a tile whose accumulator sits at TMEM address `t_row` (lane base and first
// column included): de-scale, optional squared-L2 form, top-k candidate filter, optional tile store.  `on_drained()` runs as
// soon as the accumulator has been read out completely (before the last candidate appends).
template <class OnDrained>
__device__ __forceinline__ void coarse_tile_epilogue(const GemmEpi& e, int N, int n0, int grow, bool row_ok, float rscale, float rnorm,
                                                     float rthresh, float* c_row, uint32_t t_row, const float* cs_s, const float* cn_s,
                                                     OnDrained on_drained) {
    const int nchunks = min(CG_BN / 32, (N - n0 + 31) / 32);
    float nxt[32];
    tmem_ld_32x32(t_row, nxt);
    tmem_ld_wait();
    // candidate reservation in flight: the slot index comes back from an atomic issued one chunk earlier, so its
    // latency hides behind the next chunk; up to two (value, column) pairs wait in registers
    int pend_n = 0, ppos = 0;
    float pv0 = 0.f, pv1 = 0.f;
    uint32_t pc0 = 0, pc1 = 0;
    auto flush_pending = [&]() {
        if (pend_n) {
            if (ppos < e.tk_cap) e.tk_cand[(long long)grow * e.tk_cap + ppos] = make_key(pv0, pc0);
            if (pend_n > 1 && ppos + 1 < e.tk_cap) e.tk_cand[(long long)grow * e.tk_cap + ppos + 1] = make_key(pv1, pc1);
            pend_n = 0;
        }
    };
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        const int col0 = n0 + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = nxt[j];
        if (c + 1 < nchunks) tmem_ld_32x32(t_row + uint32_t((c + 1) * 32), nxt);   // in flight while this chunk is processed
        if (col0 + 32 <= N) {
            // ---- full chunk: vectorised uniform loads, no per-element bounds checks
            {
                const float4* cs4 = reinterpret_cast<const float4*>(cs_s + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 s4 = cs4[j];
                    v[4 * j] *= s4.x; v[4 * j + 1] *= s4.y; v[4 * j + 2] *= s4.z; v[4 * j + 3] *= s4.w;
                }
            }
            if (e.mode != 0) {
                const float4* cn4 = reinterpret_cast<const float4*>(cn_s + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 n4 = cn4[j];
                    v[4 * j] = fmaxf(rnorm + n4.x - 2.f * (v[4 * j] * rscale), 1e-12f);
                    v[4 * j + 1] = fmaxf(rnorm + n4.y - 2.f * (v[4 * j + 1] * rscale), 1e-12f);
                    v[4 * j + 2] = fmaxf(rnorm + n4.z - 2.f * (v[4 * j + 2] * rscale), 1e-12f);
                    v[4 * j + 3] = fmaxf(rnorm + n4.w - 2.f * (v[4 * j + 3] * rscale), 1e-12f);
                }
            } else {
                const float ar = e.alpha * rscale;
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= ar;
            }
            flush_pending();
            // Branch-free pass mask; once the lists have warmed up a lane sees a candidate in a few percent of its chunks.
            if (e.tk_cand) {
                uint32_t mask = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) mask |= (v[j] <= rthresh) ? (1u << j) : 0u;
                if (!row_ok) mask = 0;
                if (mask) {
                    ppos = atomicAdd(e.tk_cnt + grow, __popc(mask));       // one atomic reserves all slots of this row and chunk
                    const uint32_t gcol = (uint32_t)(e.tk_idx_base + col0);
                    int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    pv0 = select32(v, j); pc0 = gcol + j; pend_n = 1;
                    if (mask) {
                        j = __ffs(mask) - 1;
                        mask &= mask - 1;
                        pv1 = select32(v, j); pc1 = gcol + j; pend_n = 2;
                    }
                    int extra = 2;
                    while (mask) {                                          // > 2 candidates in one 32-column chunk: rare
                        j = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const float val = select32(v, j);
                        if (ppos + extra < e.tk_cap) e.tk_cand[(long long)grow * e.tk_cap + ppos + extra] = make_key(val, gcol + j);
                        ++extra;
                    }
                }
            }
            if (row_ok && c_row) {
                float4* d4 = reinterpret_cast<float4*>(c_row + col0);     // ldc % 4 == 0 and 16-byte aligned C (checked on the host)
#pragma unroll
                for (int j = 0; j < 8; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        } else {
            flush_pending();
            // ---- ragged last chunk of the matrix
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const bool ok = col0 + j < N;
                float x = v[j] * (ok ? cs_s[c * 32 + j] : 1.f);
                if (e.mode != 0) x = fmaxf(rnorm + cn_s[c * 32 + j] - 2.f * (x * rscale), 1e-12f);
                else x *= e.alpha * rscale;
                v[j] = x;
            }
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {        // fully unrolled: v[] must stay in registers
                    if (col0 + j < N) {
                        if (e.tk_cand && v[j] <= rthresh) {
                            const int pos1 = atomicAdd(e.tk_cnt + grow, 1);
                            if (pos1 < e.tk_cap)
                                e.tk_cand[(long long)grow * e.tk_cap + pos1] = make_key(v[j], (uint32_t)(e.tk_idx_base + col0 + j));
                        }
                        if (c_row) c_row[col0 + j] = v[j];
                    }
                }
            }
        }
        tmem_ld_wait();
    }
    tc_fence_before();
    __syncwarp();
    on_drained();                                    // TMEM is drained: the next tile's MMAs may start
    flush_pending();
}

__global__ void __launch_bounds__(CG_THREADS, 1) coarse_gemm_kernel(const __grid_constant__ CoarseGemmParams p) {
    extern __shared__ uint8_t cg_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(cg_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CG_STAGES * CG_STAGE_BYTES);
    uint64_t* full = bars;                       // [STAGES]
    uint64_t* empty = bars + CG_STAGES;          // [STAGES]
    uint64_t* tmem_full = bars + 2 * CG_STAGES;  // [1]
    uint64_t* tmem_empty = tmem_full + 1;        // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 1);
    float* colbuf = reinterpret_cast<float*>(smem + CG_STAGES * CG_STAGE_BYTES + 256);   // [2][2][CG_BN]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.ta); tma_prefetch_desc(&p.tb);
        for (int s = 0; s < CG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1); mbar_init(tmem_empty, 8);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int num_kb = (p.K + CG_BK - 1) / CG_BK;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    // same rasterisation as gemm.cuh: groups of group_m row tiles sweep all column tiles
    auto coords = [&](int tile, int& m_tile, int& n_tile) {
        const int per_group = p.group_m * p.num_n_tiles;
        const int mg = tile / per_group;
        const int rr = tile - mg * per_group;
        const int gsize = min(p.group_m, p.num_m_tiles - mg * p.group_m);
        n_tile = rr / gsize;
        m_tile = mg * p.group_m + (rr - n_tile * gsize);
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int m_tile, n_tile;
                coords(tile, m_tile, n_tile);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * CG_STAGE_BYTES;
                    uint8_t* sB = sA + CG_A_BYTES;
                    mbar_arrive_expect_tx(&full[stage], CG_STAGE_BYTES);
                    tma_load_3d(sA, &p.ta, &full[stage], kb * CG_BK, m_tile * CG_BM, 0);
                    tma_load_3d(sB, &p.tb, &full[stage], kb * CG_BK, n_tile * CG_BN, 0);
                    if (++stage == CG_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(128, CG_BN, 0, 0);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                mbar_wait(tmem_empty, (it & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < num_kb; ++kb) {
                    const uint32_t sA = smem_u32(smem + stage * CG_STAGE_BYTES);
                    const uint32_t sB = sA + CG_A_BYTES;
                    const uint64_t dA0 = make_smem_desc(sA, 16, 1024), dA1 = make_smem_desc(sA + 128 * 128, 16, 1024);
                    const uint64_t dB = make_smem_desc(sB, 16, 1024);
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < CG_BK / 16; ++k) {
                        umma_bf16(tmem_base, dA0 + k * 2, dB + k * 2, idesc, (kb | k) ? 1u : 0u);
                        umma_bf16(tmem_base + 256, dA1 + k * 2, dB + k * 2, idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);
                    if (kb == num_kb - 1) umma_commit(tmem_full);
                    if (++stage == CG_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        const GemmEpi& e = p.epi;
        const int quad = warp & 3;                       // TMEM lane quadrant of this warp
        const int half = (warp - 2) >> 2;                // 0: rows 0-127 of the tile (TMEM columns 0-255), 1: rows 128-255
        const int row_in_tile = half * 128 + quad * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            int m_tile, n_tile;
            coords(tile, m_tile, n_tile);
            const int n0 = n_tile * CG_BN;
            const int grow = m_tile * CG_BM + row_in_tile;
            const bool row_ok = grow < p.M;
            float rscale = 1.f, rnorm = 0.f, rthresh = 0.f;
            if (row_ok) {
                if (e.row_scale) rscale = e.row_scale[grow];
                if (e.mode != 0) rnorm = e.row_norm[grow];
                if (e.tk_cand) rthresh = (e.tk_cnt[grow] > e.tk_cap) ? -__int_as_float(0x7f800000) : e.tk_thresh[grow];
            }
            float* c_row = e.C ? e.C + (long long)grow * e.ldc : nullptr;
            // column scales / norms of this tile -> shared memory while the MMAs of the tile are still running
            float* cs_s = colbuf + (it & 1) * 2 * CG_BN;
            float* cn_s = cs_s + CG_BN;
            {
                const int t = threadIdx.x - 64, col = n0 + t;     // 256 epilogue threads, one column each
                cs_s[t] = (e.col_scale && col < p.N) ? __ldg(e.col_scale + col) : 1.f;
                cn_s[t] = (e.mode != 0 && col < p.N) ? __ldg(e.col_norm + col) : 0.f;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(tmem_full, it & 1);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(half * 256);
            coarse_tile_epilogue(e, p.N, n0, grow, row_ok, rscale, rnorm, rthresh, c_row, t_row, cs_s, cn_s,
                                 [&]() { if (lane == 0) mbar_arrive(tmem_empty); });
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------ the same contraction on CTA PAIRS (cta_group::2)
// Two CTAs of a cluster (one TPC) compute one 256 x 256 tile: tcgen05.mma.cta_group::2 with M = 256 is issued by the even CTA
// only and reads A (this CTA's 128 rows) and B (this CTA's 128 of the 256 columns) from BOTH CTAs' shared memory, so each CTA
// moves 32 KB per k-block for its 4.2 MFLOP share (128 FLOP/B) and keeps its 128 x 256 accumulator half in its own TMEM --
// 256 columns, which leaves room for TWO accumulator stages: the epilogue (4 warps per CTA) overlaps the next tile's MMAs.
//   both CTAs: warp 0 = TMA producer (own A rows, own B rows; transaction bytes complete on the EVEN CTA's barrier),
//              warps 2-5 = epilogue of the own 128 rows, arriving on the even CTA's accumulator-free barrier
//   even CTA : warp 1 = MMA issuer; tcgen05.commit multicasts the stage-free / accumulator-full arrivals to both CTAs
constexpr int C2_STAGES = 6, C2_THREADS = 192;
constexpr int C2_A_BYTES = 128 * CG_BK * 2, C2_B_BYTES = 128 * CG_BK * 2;
constexpr int C2_STAGE_BYTES = C2_A_BYTES + C2_B_BYTES;
constexpr int C2_SMEM_BYTES = C2_STAGES * C2_STAGE_BYTES + 1024 + 256 + CG_COLBUF_BYTES;
constexpr uint32_t C2_PEER_MASK = 0xFEFFFFFFu;        // clears the CTA-pair bit of a shared::cluster address: the even CTA's copy

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2cta(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & C2_PEER_MASK), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {        // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta0(uint64_t* bar) {         // arrive on the even CTA's copy of `bar`
    asm volatile(
        "{\n\t.reg .b32 rem;\n\t"
        "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [rem];\n\t}"
        ::"r"(smem_u32(bar)), "r"(0u) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C2_THREADS, 1) coarse_gemm2_kernel(const __grid_constant__ CoarseGemmParams p) {
    extern __shared__ uint8_t c2_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(c2_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C2_STAGES * C2_STAGE_BYTES);
    uint64_t* full = bars;                        // [STAGES]  (used on the even CTA)
    uint64_t* empty = bars + C2_STAGES;           // [STAGES]  (one per CTA, multicast arrivals)
    uint64_t* tmem_full = bars + 2 * C2_STAGES;   // [2]       (one per CTA, multicast arrivals)
    uint64_t* tmem_empty = tmem_full + 2;         // [2]       (used on the even CTA: 4 warps x 2 CTAs arrive)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* colbuf = reinterpret_cast<float*>(smem + C2_STAGES * C2_STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.ta); tma_prefetch_desc(&p.tb);
        for (int s = 0; s < C2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2cta(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                           // both CTAs' barriers and TMEM exist before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int num_kb = (p.K + CG_BK - 1) / CG_BK;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    auto coords = [&](int tile, int& m_tile, int& n_tile) {
        const int per_group = p.group_m * p.num_n_tiles;
        const int mg = tile / per_group;
        const int rr = tile - mg * per_group;
        const int gsize = min(p.group_m, p.num_m_tiles - mg * p.group_m);
        n_tile = rr / gsize;
        m_tile = mg * p.group_m + (rr - n_tile * gsize);
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                int m_tile, n_tile;
                coords(tile, m_tile, n_tile);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * C2_STAGE_BYTES;
                    uint8_t* sB = sA + C2_A_BYTES;
                    if (leader) mbar_arrive_expect_tx(&full[stage], 2 * C2_STAGE_BYTES);      // both CTAs' boxes land on this barrier
                    tma_load_3d_2cta(sA, &p.ta, &full[stage], kb * CG_BK, m_tile * 256 + (int)rank * 128, 0);
                    tma_load_3d_2cta(sB, &p.tb, &full[stage], kb * CG_BK, n_tile * 256 + (int)rank * 128, 0);
                    if (++stage == C2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = make_idesc_f16(256, 256, 0, 0);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const uint32_t sA = smem_u32(smem + stage * C2_STAGE_BYTES);
                    const uint32_t sB = sA + C2_A_BYTES;
                    const uint64_t dA = make_smem_desc(sA, 16, 1024), dB = make_smem_desc(sB, 16, 1024);
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < CG_BK / 16; ++k) umma_f16_2cta(d_tmem, dA + k * 2, dB + k * 2, idesc, (kb | k) ? 1u : 0u);
                    umma_commit_2cta(&empty[stage]);
                    if (kb == num_kb - 1) umma_commit_2cta(&tmem_full[acc]);
                    if (++stage == C2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        const GemmEpi& e = p.epi;
        const int quad = warp & 3;
        const int row_in_tile = (int)rank * 128 + quad * 32 + lane;
        int it = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
            int m_tile, n_tile;
            coords(tile, m_tile, n_tile);
            const int n0 = n_tile * 256;
            const int grow = m_tile * 256 + row_in_tile;
            const bool row_ok = grow < p.M;
            float rscale = 1.f, rnorm = 0.f, rthresh = 0.f;
            if (row_ok) {
                if (e.row_scale) rscale = e.row_scale[grow];
                if (e.mode != 0) rnorm = e.row_norm[grow];
                if (e.tk_cand) rthresh = (e.tk_cnt[grow] > e.tk_cap) ? -__int_as_float(0x7f800000) : e.tk_thresh[grow];
            }
            float* c_row = e.C ? e.C + (long long)grow * e.ldc : nullptr;
            float* cs_s = colbuf + (it & 1) * 2 * CG_BN;
            float* cn_s = cs_s + CG_BN;
            for (int t = threadIdx.x - 64; t < CG_BN; t += 128) {          // 128 epilogue threads stage 256 column scales / norms
                const int col = n0 + t;
                cs_s[t] = (e.col_scale && col < p.N) ? __ldg(e.col_scale + col) : 1.f;
                cn_s[t] = (e.mode != 0 && col < p.N) ? __ldg(e.col_norm + col) : 0.f;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int acc = it & 1;
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * 256);
            coarse_tile_epilogue(e, p.N, n0, grow, row_ok, rscale, rnorm, rthresh, c_row, t_row, cs_s, cn_s, [&]() {
                if (lane == 0) { if (leader) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_cta0(&tmem_empty[acc]); }
            });
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                           // the peer may still be multicasting into this CTA's barriers / reading its smem
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

}  // namespace grl
