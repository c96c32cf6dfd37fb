// grl_b200 — the split-bf16 GEMM of gemm.cuh on CTA PAIRS (sm_100a, tcgen05.mma.cta_group::2).
//
// Same contraction, operands, majors and fused epilogue as gemm_bf16x3_kernel<256, ...>; two CTAs of a cluster (one TPC)
// compute one 256 x 256 tile.  The even CTA issues UMMA 256 x 256 x 16 (hi*hi + lo*hi + hi*lo) that reads A (this CTA's
// 128 rows) and B (this CTA's 128 of the 256 columns) from BOTH CTAs' shared memory, so every CTA moves 64 KB per k-block
// (16 KB per plane and operand) instead of 96 KB for its 12.6 MFLOP share -- 197 instead of 131 FLOP per byte of
// L2 -> shared-memory traffic -- and THREE stages fit in shared memory instead of two (two stages in flight while one is
// consumed: the single-CTA kernel's k-loop stalls on TMA latency with one).  Each CTA keeps its 128 x 256 accumulator
// half in its own TMEM (two accumulator stages = 512 columns), so its four epilogue warps run the epilogue of gemm.cuh
// unchanged (thread == output row) while the next tile's MMAs are issued.
//   both CTAs: warp 0 = TMA producer (own A rows, own B columns; hi and lo planes complete on separate barriers of the EVEN CTA)
//              warps 2-5 = epilogue of the own 128 rows, arriving on the even CTA's accumulator-free barrier
//   even CTA : warp 1 = MMA issuer; tcgen05.commit multicasts the stage-free / accumulator-full arrivals to both CTAs
#pragma once
#include "coarse_gemm.cuh"

namespace grl {

constexpr int GP_PLANE = 128 * GEMM_BK * 2;                 // one plane of one operand half: 128 rows x 64 k x 2 bytes = 16 KB
constexpr int GP_BM = 256, GP_BN = 256;
// PLANES = 3: split-bf16 (A_hi | B_hi | A_lo | B_lo = 64 KB per stage, 3 stages); PLANES = 1: one fp16 plane per operand, one MMA
// per k-step (32 KB per stage, 6 stages) -- the single-pass gradient GEMMs of the attention convs
template <int PLANES>
struct GpCfg {
    static constexpr int STAGES = PLANES == 1 ? 6 : 3;
    static constexpr int STAGE_BYTES = (PLANES == 1 ? 2 : 4) * GP_PLANE;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

// tile index -> (batch, 256-row tile, 256-column tile); same grouped rasterisation as gemm.cuh with p.num_m_tiles counted in
// 256-row tiles (the host passes pair params: num_m_tiles = ceil(M / 256))
template <bool A_MN, bool B_MN, int PLANES = 3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1) gemm_pair_bf16x3_kernel(const __grid_constant__ GemmParams p) {
    constexpr int GP_STAGES = GpCfg<PLANES>::STAGES;
    constexpr int GP_STAGE_BYTES = GpCfg<PLANES>::STAGE_BYTES;
    extern __shared__ uint8_t gp_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gp_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GP_STAGES * GP_STAGE_BYTES);
    uint64_t* full_hi = bars;                        // [STAGES]  (used on the even CTA)
    uint64_t* full_lo = bars + GP_STAGES;            // [STAGES]  (used on the even CTA)
    uint64_t* empty = bars + 2 * GP_STAGES;          // [STAGES]  (one per CTA, multicast arrivals)
    uint64_t* tmem_full = bars + 3 * GP_STAGES;      // [2]       (one per CTA, multicast arrivals)
    uint64_t* tmem_empty = tmem_full + 2;            // [2]       (used on the even CTA: 4 warps x 2 CTAs arrive)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.ta_hi); tma_prefetch_desc(&p.ta_lo);
        tma_prefetch_desc(&p.tb_hi); tma_prefetch_desc(&p.tb_lo);
        for (int s = 0; s < GP_STAGES; ++s) { mbar_init(&full_hi[s], 1); mbar_init(&full_lo[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2cta(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                              // both CTAs' barriers and TMEM exist before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
    const int tiles_per_batch = p.num_m_tiles * p.num_n_tiles;
    const int num_tiles = tiles_per_batch * p.batch;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                int z, m_tile, n_tile;
                tile_coords(p, tile, tiles_per_batch, z, m_tile, n_tile);
                const int m0 = m_tile * GP_BM + (int)rank * 128;       // this CTA's A rows
                const int n0 = n_tile * GP_BN + (int)rank * 128;       // this CTA's B rows (output columns)
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int k0 = kb * GEMM_BK;
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sA_hi = smem + stage * GP_STAGE_BYTES;
                    uint8_t* sB_hi = sA_hi + GP_PLANE;
                    uint8_t* sA_lo = sB_hi + GP_PLANE;
                    uint8_t* sB_lo = sA_lo + GP_PLANE;
                    if (leader) mbar_arrive_expect_tx(&full_hi[stage], 4 * GP_PLANE);          // both CTAs' hi boxes (A and B) land on this barrier
                    if (A_MN) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) tma_load_3d_2cta(sA_hi + j * 8192, &p.ta_hi, &full_hi[stage], m0 + 64 * j, k0, z);
                    } else {
                        tma_load_3d_2cta(sA_hi, &p.ta_hi, &full_hi[stage], k0, m0, z);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) tma_load_3d_2cta(sB_hi + j * 8192, &p.tb_hi, &full_hi[stage], n0 + 64 * j, k0, z);
                    } else {
                        tma_load_3d_2cta(sB_hi, &p.tb_hi, &full_hi[stage], k0, n0, z);
                    }
                    if (PLANES == 1) {
                        if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    if (leader) mbar_arrive_expect_tx(&full_lo[stage], 4 * GP_PLANE);
                    if (A_MN) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) tma_load_3d_2cta(sA_lo + j * 8192, &p.ta_lo, &full_lo[stage], m0 + 64 * j, k0, z);
                    } else {
                        tma_load_3d_2cta(sA_lo, &p.ta_lo, &full_lo[stage], k0, m0, z);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) tma_load_3d_2cta(sB_lo + j * 8192, &p.tb_lo, &full_lo[stage], n0 + 64 * j, k0, z);
                    } else {
                        tma_load_3d_2cta(sB_lo, &p.tb_lo, &full_lo[stage], k0, n0, z);
                    }
                    if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (even CTA only) =====================
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = (PLANES == 1) ? make_idesc_f16(GP_BM, GP_BN, A_MN ? 1 : 0, B_MN ? 1 : 0)
                                                     : make_idesc_bf16(GP_BM, GP_BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            constexpr uint32_t a_lbo = A_MN ? GEMM_BK * 128 : 16, b_lbo = B_MN ? GEMM_BK * 128 : 16;
            constexpr uint32_t a_kstep = A_MN ? (2048 >> 4) : (32 >> 4), b_kstep = B_MN ? (2048 >> 4) : (32 >> 4);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * GP_BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const uint32_t sA_hi = smem_u32(smem + stage * GP_STAGE_BYTES);
                    const uint32_t sB_hi = sA_hi + GP_PLANE;
                    const uint32_t sA_lo = sB_hi + GP_PLANE;
                    const uint32_t sB_lo = sA_lo + GP_PLANE;
                    const uint64_t dA_hi = make_smem_desc(sA_hi, a_lbo, 1024), dA_lo = make_smem_desc(sA_lo, a_lbo, 1024);
                    const uint64_t dB_hi = make_smem_desc(sB_hi, b_lbo, 1024), dB_lo = make_smem_desc(sB_lo, b_lbo, 1024);
                    mbar_wait(&full_hi[stage], phase);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k)
                        umma_f16_2cta(d_tmem, dA_hi + k * a_kstep, dB_hi + k * b_kstep, idesc, (kb | k) ? 1u : 0u);
                    if (PLANES == 3) {
                        mbar_wait(&full_lo[stage], phase);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) {
                            umma_f16_2cta(d_tmem, dA_lo + k * a_kstep, dB_hi + k * b_kstep, idesc, 1u);
                            umma_f16_2cta(d_tmem, dA_hi + k * a_kstep, dB_lo + k * b_kstep, idesc, 1u);
                        }
                    }
                    umma_commit_2cta(&empty[stage]);               // frees the stage in BOTH CTAs once these MMAs retire
                    if (kb == num_kb - 1) umma_commit_2cta(&tmem_full[acc]);
                    if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5 of both CTAs): the own 128 rows x 256 columns =====================
        const int quad = warp & 3;
        const int row_in_tile = quad * 32 + lane;
        int it = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
            int z, m_tile, n_tile;
            tile_coords(p, tile, tiles_per_batch, z, m_tile, n_tile);
            const int m_tile128 = m_tile * 2 + (int)rank;      // the unit of the statistic / `sub` layouts
            const int n0 = n_tile * GP_BN;
            const int acc = it & 1;
            const bool live = m_tile128 * GEMM_BM < p.M;       // M / 128 odd: the odd CTA of the last row tile has no rows
            EpiRow R;
            if (live) R = epi_row_setup(p, z, m_tile128, row_in_tile);
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            tc_fence_after();
            if (live) epi_tile<GP_BN>(p, R, z, m_tile128, n0, tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * GP_BN), quad, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (leader) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_cta0(&tmem_empty[acc]); }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                              // the peer may still be multicasting into this CTA's barriers / reading its smem
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

}  // namespace grl
