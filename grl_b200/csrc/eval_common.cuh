// grl_b200 — device helpers shared by the matching kernels (eval.cu) and the retrieval search (search.cu).
#pragma once
#include <math_constants.h>

#include "common.cuh"
#include "../../include/grl_b200.h"

namespace grl {

constexpr int TOPK_THREADS = 256;
constexpr int TOPK_WAVE = TOPK_THREADS * 4;
constexpr int TOPK_BUF = 4096;                 // total sort size (running list + staged candidates)
constexpr int TOPK_MAXK = 1024;
constexpr uint64_t KEY_EMPTY = ~0ull;          // sorts after every (distance, index) key, NaN distances included

__device__ __forceinline__ float key_value(uint64_t key) { return from_orderable((uint32_t)(key >> 32)); }
__device__ __forceinline__ uint32_t key_index(uint64_t key) { return (uint32_t)(key & 0xFFFFFFFFu); }

// In-place ascending bitonic sort of n (power of two) keys in shared memory by the whole block.
__device__ __forceinline__ void block_bitonic_sort(uint64_t* keys, int n) {
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

// Fixed-order fp32 inner product of a shared-memory query row with a global gallery row, by one warp: lane l accumulates the
// float4 chunks l, l+32, ... component by component (separate multiply and add, no FMA), then an xor-shuffle tree.  Also returns
// |g|^2 in the same order.  The parity tests restate this order in numpy (exact_distance_fixed of the evaluator restatement).
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

}  // namespace grl
