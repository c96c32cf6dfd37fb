// grl_b200 — GCE + TRL head, backward (sm_100a).
//
// Hand-derived backward of grl_head_forward, i.e. the autograd graph of
//   Backbone.forward after self.base   /root/reference/reid/models/basebranch.py:56-68
//   TRLBlock.forward (+ BasicBlock)    /root/reference/reid/models/grl_model.py:131-180, 67-85
// in train mode (batch-statistics BatchNorm).  Back-propagation through time over the T memory
// updates, both directions batched (z = 0 forward, 1 backward) in every launch.
//   dgrad  = split-bf16 tcgen05 GEMM, A K-major (gradient planes), B MN-major (the forward's
//            weight planes [Cout][Cin] read as B[n=Cin][k=Cout]: no transposed copies)
//   wgrad  = same GEMM with both operands MN-major (K = pixels), accumulated over the T steps
//   BN     = reduce (sum g, sum g*xhat) -> finalize -> apply, HBM-bound, 128-bit accesses
// Buffers live in the caller's workspace (head_common.cuh); `dxu` is used as dZ[T][2][R][C].
#include <stddef.h>

#include <algorithm>

#include "head_common.cuh"

namespace grl {

struct BnPtrs2 { const float* gamma[2]; };
struct OutPtrs2 { float* p[2]; };

// ------------------------------------------------------------------ generic small helpers
// out[j*ldo + k] (+)= sum_{o<no, n<ni} A[o*a_os + n*a_is + j] * Bm[o*b_os + n*b_is + k]
// (outer products over a few hundred rows: SE weights, glo_fc, the glo half of corr_atte.0)
__global__ void __launch_bounds__(256) small_outer_kernel(const float* __restrict__ A, long long a_os, long long a_is,
                                                          const float* __restrict__ Bm, long long b_os, long long b_is,
                                                          int no, int ni, float* __restrict__ out, long long ldo, float scale) {
    __shared__ float sA[16][32];
    __shared__ float sB[16][64];
    const int k0 = blockIdx.x * 64, j0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int rows = no * ni;
    for (int r0 = 0; r0 < rows; r0 += 16) {
        __syncthreads();
        for (int e = threadIdx.x; e < 16 * 32; e += 256) {
            const int r = r0 + (e >> 5);
            float v = 0.f;
            if (r < rows) v = A[(long long)(r / ni) * a_os + (long long)(r % ni) * a_is + j0 + (e & 31)];
            sA[e >> 5][e & 31] = v;
        }
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int r = r0 + (e >> 6);
            float v = 0.f;
            if (r < rows) v = Bm[(long long)(r / ni) * b_os + (long long)(r % ni) * b_is + k0 + (e & 63)];
            sB[e >> 6][e & 63] = v;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const float b = sB[r][tx];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) acc[jj] += sA[r][ty * 8 + jj] * b;
        }
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) out[(long long)(j0 + ty * 8 + jj) * ldo + k0 + tx] = acc[jj] * scale;
}

// out[y*out_stride + c] = sum_{o<no, n<ni} src[y*grp_stride + o*os + n*is + c]
__global__ void small_colsum_kernel(const float* __restrict__ src, long long grp_stride, long long os, long long is, int no, int ni,
                                    float* __restrict__ out, long long out_stride, int Cn) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cn) return;
    const float* s = src + (long long)blockIdx.y * grp_stride + c;
    float acc = 0.f;
    for (int o = 0; o < no; ++o)
        for (int n = 0; n < ni; ++n) acc += s[(long long)o * os + (long long)n * is];
    out[(long long)blockIdx.y * out_stride + c] = acc;
}

// dst[b][s][c] = src[b][c] * scale  for the 128 pixels of every clip row block (both directions): grid (C/64, B, 2)
__global__ void __launch_bounds__(256) bcast_rows_kernel(const float* __restrict__ src, float scale, int R, float* __restrict__ dst) {
    const int z = blockIdx.z, b = blockIdx.y, c0 = blockIdx.x * 64;
    const Tile t;
    float v[8];
    load8(src + (size_t)b * HC + c0 + t.cg * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= scale;
#pragma unroll
    for (int k = 0; k < 4; ++k) store8(dst + ((size_t)z * R + (size_t)b * HS + t.r0 + 32 * k) * HC + c0 + t.cg * 8, v);
}

// ------------------------------------------------------------------ squeeze-excite backward, all steps at once
// grid (B, 2, T).  da = d_f_corr[b,tau] * gc[b,tau];  ds = da a (1-a);  dh = (ds L2) [h>0];  dq = dh L1
struct SePtrsB { const float* l1[2]; const float* l2[2]; };
__global__ void __launch_bounds__(256) se_bwd_kernel(const float* __restrict__ dfc, const float* __restrict__ gc, const float* __restrict__ se_a,
                                                     const float* __restrict__ se_h, SePtrsB sp, int B, int T, float* __restrict__ se_ds,
                                                     float* __restrict__ se_dh, float* __restrict__ se_dq, const float* __restrict__ se_q,
                                                     unsigned int* __restrict__ df_bound_bits) {
    __shared__ float ds[HC];
    __shared__ float wmax[8];
    __shared__ float part[2][HSE];
    __shared__ float dh[HSE];
    const int b = blockIdx.x, d = blockIdx.y, i = blockIdx.z;
    const int tau = d ? T - 1 - i : i;
    const size_t n = (size_t)b * T + tau;
    const size_t slot = ((size_t)i * 2 + d) * B + b;
    for (int c = threadIdx.x; c < HC; c += 256) {
        const float a = se_a[slot * HC + c];
        const float v = dfc[n * HC + c] * gc[n * HC + c] * a * (1.f - a);
        ds[c] = v;
        se_ds[slot * HC + c] = v;
    }
    __syncthreads();
    {
        const int j = threadIdx.x & (HSE - 1), half = threadIdx.x >> 7;
        const float* w = sp.l2[d] + j;
        float acc = 0.f;
        const int c_begin = half * (HC / 2);
#pragma unroll 8
        for (int c = c_begin; c < c_begin + HC / 2; ++c) acc += ds[c] * __ldg(w + (size_t)c * HSE);
        part[half][j] = acc;
    }
    __syncthreads();
    if (threadIdx.x < HSE) {
        const float hv = se_h[slot * HSE + threadIdx.x];
        const float v = hv > 0.f ? part[0][threadIdx.x] + part[1][threadIdx.x] : 0.f;
        dh[threadIdx.x] = v;
        se_dh[slot * HSE + threadIdx.x] = v;
    }
    __syncthreads();
    float bound = 0.f;
    for (int c = threadIdx.x; c < HC; c += 256) {
        const float* w = sp.l1[d] + c;
        float acc = 0.f;
#pragma unroll 8
        for (int j = 0; j < HSE; ++j) acc += dh[j] * __ldg(w + (size_t)j * HC);
        se_dq[slot * HC + c] = acc;
        // |dF1|, |dF2| = (2/S) |dq| |E| <= 2 |dq| sqrt(q / S)  (sum_s E^2 = S q): an upper bound of every entry trl_bwd_f1_kernel
        // writes, from which it picks the power-of-two scale of its fp16 outputs -- no pass over the outputs themselves
        bound = fmaxf(bound, 2.02f * fabsf(acc) * sqrtf(fmaxf(se_q[slot * HC + c], 0.f) * (1.f / HS)));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, off));
    if (lane_id() == 0) wmax[threadIdx.x >> 5] = bound;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int i = 0; i < 8; ++i) m = fmaxf(m, wmax[i]);
        if (m == m && m > 0.f) atomicMax(df_bound_bits, __float_as_uint(m));      // non-negative floats order like their bits
    }
}

// dgc[n][c] = d_f_corr[n][c] * (2 + a_fwd[step t] + a_bwd[step T-1-t])      (F4: out = (1+a) * GAP(x_corr))
__global__ void dgc_kernel(const float* __restrict__ dfc, const float* __restrict__ se_a, int B, int T, float* __restrict__ dgc) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * T * HC) return;
    const int c = (int)(idx % HC);
    const size_t n = idx / HC;
    const int b = (int)(n / T), t = (int)(n % T);
    const float a0 = se_a[(((size_t)t * 2 + 0) * B + b) * HC + c];
    const float a1 = se_a[(((size_t)(T - 1 - t) * 2 + 1) * B + b) * HC + c];
    dgc[idx] = dfc[idx] * (2.f + a0 + a1);
}

// ------------------------------------------------------------------ bf16 hi/lo planes -> ONE fp16 plane with a device-chosen scale
// The weight / input gradients of the attention convs f1 / f2 run as single-pass fp16 GEMMs (gemm_launch_x1): emulated on the
// fp64 plan, rounding the operands of exactly those contractions to fp16 leaves every head output and gradient where the
// three-MMA split-bf16 form puts them (tools/exp_f1f2_precision.py; the FORWARD f1 / f2 must stay split-bf16).  fp16 needs a
// scale: each tensor gets one power of two that maps its largest magnitude into [2^14, 2^15) -- 28 binades of full precision
// below the maximum -- found by a pass over the hi plane; the reciprocal is handed to the GEMM epilogue as a device scalar.
__global__ void __launch_bounds__(256) absmax_bf16_kernel(const __nv_bfloat16* __restrict__ x, size_t n8, unsigned int* __restrict__ out_bits) {
    unsigned int m = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
        const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { m = max(m, w[j] & 0x7FFFu); m = max(m, (w[j] >> 16) & 0x7FFFu); }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (lane_id() == 0 && m) atomicMax(out_bits, m << 16);       // bf16 bits << 16 == the float's bits: orders like the magnitude
}
// scal[0] = amax bits (in), scal[1] = 1 / scale (out).  NaN / inf maxima fall back to scale 1.
__global__ void __launch_bounds__(256) planes_to_f16_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, size_t n8,
                                                            float* __restrict__ scal, __half* __restrict__ out) {
    const unsigned int mb = __float_as_uint(scal[0]);
    int e = (int)((mb >> 23) & 0xff) - 127;
    if (mb == 0 || ((mb >> 23) & 0xff) == 0xff) e = 14;
    const float s = ldexpf(1.f, 14 - e);
    if (blockIdx.x == 0 && threadIdx.x == 0) scal[1] = ldexpf(1.f, e - 14);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        float v[8];
        load8_planes(hi + i * 8, lo + i * 8, v);
        __half2 h2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h2[j] = __floats2half2_rn(v[2 * j] * s, v[2 * j + 1] * s);
        *reinterpret_cast<uint4*>(out + i * 8) = *reinterpret_cast<const uint4*>(h2);
    }
}
// planes [n] -> fp16 plane `out` + scal[slot] (amax, 1/scale); everything on `st`
static int planes_to_f16(grl_handle* h, cudaStream_t st, const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t n, float* scal, __half* out) {
    const size_t n8 = n / 8;
    const int blocks = (int)std::min<size_t>((n8 + 255) / 256, (size_t)h->num_sms * 16);
    GRL_CUDA(h, cudaMemsetAsync(scal, 0, 8, st));
    absmax_bf16_kernel<<<blocks, 256, 0, st>>>(hi, n8, reinterpret_cast<unsigned int*>(scal));
    GRL_LAUNCH_CHECK(h);
    planes_to_f16_kernel<<<blocks, 256, 0, st>>>(hi, lo, n8, scal, out);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

// ------------------------------------------------------------------ f1/f2 squared-difference backward for one step
// grid (C/64, B, 2).  E = F1 - F2[tau];  dE = (2/S) dq E;  dF1 = dE [F1>0];  dF2[tau] = -dE [F2>0]
__global__ void __launch_bounds__(256) trl_bwd_f1_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ dq,
                                                         int B, int T, int R, int tau0, int tau1, float* __restrict__ scal,
                                                         __half* __restrict__ df1_16, __half* __restrict__ df2_16, float* __restrict__ dbf1_part,
                                                         float* __restrict__ dbf2_part) {
    __shared__ float red[32 * 65];
    const int z = blockIdx.z, b = blockIdx.y, c0 = blockIdx.x * 64;
    const int tau = z ? tau1 : tau0;
    const Tile t;
    // the outputs are ONE fp16 plane each (single-pass gradient GEMMs): power-of-two scale from se_bwd_kernel's bound on |dF|,
    // scal[0] = bound bits (in), scal[1] = 1 / scale (out, the same value from every launch of one backward)
    const unsigned int mb = __float_as_uint(scal[0]);
    int ex = (int)((mb >> 23) & 0xff) - 127;
    if (mb == 0 || ((mb >> 23) & 0xff) == 0xff) ex = 14;
    const float sc = ldexpf(1.f, 14 - ex);
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) scal[1] = ldexpf(1.f, ex - 14);
    float q[8], s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    load8(dq + ((size_t)z * B + b) * HC + c0 + t.cg * 8, q);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] *= (2.f / HS);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int s = t.r0 + 32 * k;
        const size_t off1 = ((size_t)z * R + (size_t)b * HS + s) * HC + c0 + t.cg * 8;
        const size_t off2 = (((size_t)b * T + tau) * HS + s) * (2 * HC) + (size_t)z * HC + c0 + t.cg * 8;
        float a[8], f[8], g1[8], g2[8];
        load8(f1 + off1, a);
        load8(f2 + off2, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float de = q[i] * (a[i] - f[i]);
            g1[i] = a[i] > 0.f ? de : 0.f;
            g2[i] = f[i] > 0.f ? -de : 0.f;
            s1[i] += g1[i]; s2[i] += g2[i];
        }
        store8_f16(df1_16 + off1, g1, sc);
        store8_f16(df2_16 + off2, g2, sc);
    }
    tile_colsum(s1, red, dbf1_part + ((size_t)z * B + b) * HC + c0, t);
    tile_colsum(s2, red, dbf2_part + ((size_t)z * B + b) * HC + c0, t);
}

// ------------------------------------------------------------------ BatchNorm backward (train mode), 3 phases
// g = (srcA [+ srcB]) * [relu output > 0];  xhat = (H - mean) * rstd
//   reduce  : per 128-row tile column sums of g and g*xhat            grid (Cn/64, R/128, 2)
//   finalize: k0 = gamma rstd, k1 = sum g / n, k2 = sum g xhat / n;  dgamma, dbeta (+)=
//   apply   : dH = k0 (g - k1 - xhat k2) as bf16 planes (+ optional fp32 copy of g)
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ srcA, const float* __restrict__ srcB,
                                                            const __nv_bfloat16* __restrict__ mask_hi, const float* __restrict__ hraw,
                                                            const float* __restrict__ stat, int Cn, int R, float* __restrict__ psum,
                                                            float* __restrict__ pxh, unsigned int* __restrict__ scal) {
    __shared__ float red[32 * 65];
    __shared__ float wmx[2][8];
    const int z = blockIdx.z, c0 = blockIdx.x * 64;
    const Tile t;
    const float* st = stat + (size_t)z * 4 * Cn;
    float mean[8], rstd[8], as[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ax[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float gmax = 0.f, xmax = 0.f;
    load8(st + 2 * Cn + c0 + t.cg * 8, mean); load8(st + 3 * Cn + c0 + t.cg * 8, rstd);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t off = ((size_t)z * R + blockIdx.y * 128 + t.r0 + 32 * k) * Cn + c0 + t.cg * 8;
        float g[8], mk[8], h[8];
        load8(srcA + off, g);
        if (srcB) {
            float g2[8];
            load8(srcB + off, g2);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] += g2[i];
        }
        load8_hi(mask_hi + off, mk);
        load8(hraw + off, h);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float gv = mk[i] > 0.f ? g[i] : 0.f;
            const float xh = (h[i] - mean[i]) * rstd[i];
            as[i] += gv;
            ax[i] += gv * xh;
            gmax = fmaxf(gmax, fabsf(gv)); xmax = fmaxf(xmax, fabsf(xh));
        }
    }
    const size_t pidx = ((size_t)z * gridDim.y + blockIdx.y) * Cn + c0;
    tile_colsum(as, red, psum + pidx, t);
    tile_colsum(ax, red, pxh + pidx, t);
    if (scal) {     // scal[0] = max |g|, scal[1] = max |xhat| over the launch: bn_bwd_finalize turns them into a bound on |dH|
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) { gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, off)); xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, off)); }
        if (lane_id() == 0) { wmx[0][threadIdx.x >> 5] = gmax; wmx[1][threadIdx.x >> 5] = xmax; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f, b = 0.f;
            for (int i = 0; i < 8; ++i) { a = fmaxf(a, wmx[0][i]); b = fmaxf(b, wmx[1][i]); }
            if (a == a && a > 0.f) atomicMax(scal + 0, __float_as_uint(a));
            if (b == b && b > 0.f) atomicMax(scal + 1, __float_as_uint(b));
        }
    }
}

__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ psum, const float* __restrict__ pxh, int nparts, int Cn,
                                                              double count, BnPtrs2 bp, const float* __restrict__ stat,
                                                              float* __restrict__ kcoef, OutPtrs2 dgamma, OutPtrs2 dbeta, int accumulate,
                                                              unsigned int* __restrict__ scal) {
    __shared__ double sh[2][8][33];
    const int z = blockIdx.y;
    const int c = blockIdx.x * 32 + threadIdx.x;          // block (32, 8): 8 lanes share the partials of a channel
    double s = 0.0, x = 0.0;
    if (c < Cn)
        for (int i = threadIdx.y; i < nparts; i += 8) {
            s += psum[((size_t)z * nparts + i) * Cn + c];
            x += pxh[((size_t)z * nparts + i) * Cn + c];
        }
    sh[0][threadIdx.y][threadIdx.x] = s; sh[1][threadIdx.y][threadIdx.x] = x;
    __syncthreads();
    if (threadIdx.y != 0 || c >= Cn) return;
    s = 0.0; x = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += sh[0][i][threadIdx.x]; x += sh[1][i][threadIdx.x]; }
    const float rstd = stat[(size_t)z * 4 * Cn + 3 * Cn + c];
    float* kc = kcoef + (size_t)z * 3 * Cn;
    kc[c] = bp.gamma[z][c] * rstd;
    kc[Cn + c] = (float)(s / count);
    kc[2 * Cn + c] = (float)(x / count);
    if (scal) {     // |dH| = |k0 (g - k1 - xhat k2)| <= |k0| (max|g| + |k1| + max|xhat| |k2|): scal[2] = the largest such bound
        const float bnd = 1.01f * fabsf(kc[c]) * (__uint_as_float(scal[0]) + fabsf(kc[Cn + c]) + __uint_as_float(scal[1]) * fabsf(kc[2 * Cn + c]));
        if (bnd == bnd && bnd > 0.f) atomicMax(scal + 2, __float_as_uint(bnd));
    }
    if (accumulate) { dgamma.p[z][c] += (float)x; dbeta.p[z][c] += (float)s; }
    else { dgamma.p[z][c] = (float)x; dbeta.p[z][c] = (float)s; }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ srcA, const float* __restrict__ srcB,
                                                           const __nv_bfloat16* __restrict__ mask_hi, const float* __restrict__ hraw,
                                                           const float* __restrict__ stat, const float* __restrict__ kcoef, int Cn, int R,
                                                           __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                                           float* __restrict__ g_out, float* __restrict__ scal, __half* __restrict__ out16) {
    const int z = blockIdx.z, c0 = blockIdx.x * 64;
    const Tile t;
    const float* st = stat + (size_t)z * 4 * Cn;
    const float* kc = kcoef + (size_t)z * 3 * Cn;
    float sc16 = 1.f;
    if (out16) {    // one more copy of dH as ONE fp16 plane (the operand of the single-pass weight-gradient GEMM): power-of-two scale
                    // from the bound in scal[2]; scal[3] = 1 / scale for that GEMM's epilogue
        const unsigned int mb = __float_as_uint(scal[2]);
        int ex = (int)((mb >> 23) & 0xff) - 127;
        if (mb == 0 || ((mb >> 23) & 0xff) == 0xff) ex = 14;
        sc16 = ldexpf(1.f, 14 - ex);
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) scal[3] = ldexpf(1.f, ex - 14);
    }
    float mean[8], rstd[8], k0[8], k1[8], k2[8];
    load8(st + 2 * Cn + c0 + t.cg * 8, mean); load8(st + 3 * Cn + c0 + t.cg * 8, rstd);
    load8(kc + c0 + t.cg * 8, k0); load8(kc + Cn + c0 + t.cg * 8, k1); load8(kc + 2 * Cn + c0 + t.cg * 8, k2);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t off = ((size_t)z * R + blockIdx.y * 128 + t.r0 + 32 * k) * Cn + c0 + t.cg * 8;
        float g[8], mk[8], h[8], o[8];
        load8(srcA + off, g);
        if (srcB) {
            float g2[8];
            load8(srcB + off, g2);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] += g2[i];
        }
        load8_hi(mask_hi + off, mk);
        load8(hraw + off, h);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            g[i] = mk[i] > 0.f ? g[i] : 0.f;
            o[i] = k0[i] * (g[i] - k1[i] - (h[i] - mean[i]) * rstd[i] * k2[i]);
        }
        store8_planes(out_hi + off, out_lo + off, o);
        if (out16) store8_f16(out16 + off, o, sc16);
        if (g_out) store8(g_out + off, g);
    }
}

// ------------------------------------------------------------------ optional upstream grads of the stand-alone maps
// dst_pm[n*S + s][c] (+)= src_nchw[n][c][s]     grid (C/64, N)
__global__ void __launch_bounds__(256) nchw_to_pm_kernel(const float* __restrict__ src, float* __restrict__ dst, int add) {
    __shared__ float tile[64][129];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = warp * 8 + i;
        const float4 v = *reinterpret_cast<const float4*>(src + ((size_t)n * HC + c0 + c) * HS + lane * 4);
        tile[c][lane * 4 + 0] = v.x; tile[c][lane * 4 + 1] = v.y; tile[c][lane * 4 + 2] = v.z; tile[c][lane * 4 + 3] = v.w;
    }
    __syncthreads();
    const Tile t;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int s = t.r0 + 32 * k;
        float* p = dst + ((size_t)n * HS + s) * HC + c0 + t.cg * 8;
        float v[8];
        if (add) load8(p, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (add ? v[i] : 0.f) + tile[t.cg * 8 + i][s];
        store8(p, v);
    }
}

// ------------------------------------------------------------------ GCE backward: gate
// grid (C/64, N).  Collects d x_uncorr / d x_corr from the TRL backward, applies the gate:
//   dXu = dZ[t][fwd] + dZ[T-1-t][bwd] + (dM0_fwd + dM0_bwd)/T,  dM0_d = dmem_d + dZ[0][d]
//   dXc = dxc + dgc/S
//   dX1 = dXc m + dXu (1-m)   (in place over dxc);   dm_part[cblock][p] = sum_c (dXc - dXu) x
__global__ void __launch_bounds__(256) gce_bwd_gate_kernel(const float* __restrict__ dz, const float* __restrict__ dmem, float* __restrict__ dxc,
                                                           const float* __restrict__ dgc, const __nv_bfloat16* __restrict__ xh,
                                                           const __nv_bfloat16* __restrict__ xl, const float* __restrict__ m, int B, int T,
                                                           const float* __restrict__ dxu_extra, int use_trl, float* __restrict__ dm_part) {
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const int b = n / T, tt = n - b * T;
    const int R = B * HS;
    const size_t P = (size_t)B * T * HS;
    const Tile t;
    float gcv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (use_trl) load8(dgc + (size_t)n * HC + c0 + t.cg * 8, gcv);
    const float invT = 1.f / (float)T;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int s = t.r0 + 32 * k;
        const size_t col = c0 + t.cg * 8;
        const size_t rrow = (size_t)b * HS + s;
        const size_t p = (size_t)n * HS + s;
        float u[8] = {0, 0, 0, 0, 0, 0, 0, 0}, v[8], a[8];
        if (use_trl) {
            load8(dz + (((size_t)tt * 2 + 0) * R + rrow) * HC + col, u);
            load8(dz + (((size_t)(T - 1 - tt) * 2 + 1) * R + rrow) * HC + col, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] += v[i];
            load8(dz + ((size_t)0 * R + rrow) * HC + col, v);
            load8(dz + ((size_t)1 * R + rrow) * HC + col, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += a[i];
            load8(dmem + ((size_t)0 * R + rrow) * HC + col, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += a[i];
            load8(dmem + ((size_t)1 * R + rrow) * HC + col, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] += (v[i] + a[i]) * invT;   // u = dXu
        }
        if (dxu_extra) {                                                 // upstream gradient on the stand-alone x_uncorr
            load8(dxu_extra + p * HC + col, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] += a[i];
        }
        float c[8], x[8], o[8];
        load8(dxc + p * HC + col, c);
        load8_planes(xh + p * HC + col, xl + p * HC + col, x);
        const float mp = m[p];
        float dmv = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            c[i] += gcv[i] * (1.f / HS);                                  // c = dXc
            o[i] = c[i] * mp + u[i] * (1.f - mp);
            dmv += (c[i] - u[i]) * x[i];
        }
        store8(dxc + p * HC + col, o);
        dmv += __shfl_xor_sync(0xffffffffu, dmv, 1);
        dmv += __shfl_xor_sync(0xffffffffu, dmv, 2);
        dmv += __shfl_xor_sync(0xffffffffu, dmv, 4);
        if (t.cg == 0) dm_part[(size_t)blockIdx.x * P + p] = dmv;
    }
}

// dm -> dz3 = dm m (1-m);  block partial sums for the scalar BN (corr_atte.6)       grid (P/256)
__global__ void __launch_bounds__(256) gce_bwd_dm_kernel(const float* __restrict__ dm_part, int nparts, const float* __restrict__ d_corr_map,
                                                         const float* __restrict__ m, const float* __restrict__ y3, const float* __restrict__ stat3,
                                                         int P, float* __restrict__ dz3, float* __restrict__ psum, float* __restrict__ pxh) {
    __shared__ float sh[2][8];
    const int p = blockIdx.x * 256 + threadIdx.x;
    float g = 0.f, gx = 0.f;
    if (p < P) {
        float d = d_corr_map ? d_corr_map[p] : 0.f;
        for (int i = 0; i < nparts; ++i) d += dm_part[(size_t)i * P + p];
        const float mp = m[p];
        g = d * mp * (1.f - mp);
        dz3[p] = g;
        gx = g * (y3[p] - stat3[2]) * stat3[3];
    }
    g = warp_sum(g); gx = warp_sum(gx);
    const int warp = threadIdx.x >> 5;
    if (lane_id() == 0) { sh[0][warp] = g; sh[1][warp] = gx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) { a += sh[0][i]; b += sh[1][i]; }
        psum[blockIdx.x] = a; pxh[blockIdx.x] = b;
    }
}

// scalar BN backward -> dy3; block 0 writes d gamma / d beta of corr_atte.6            grid (P/256)
__global__ void __launch_bounds__(256) gce_bwd_dy3_kernel(const float* __restrict__ dz3, const float* __restrict__ psum, const float* __restrict__ pxh,
                                                          int nparts, const float* __restrict__ y3, const float* __restrict__ stat3,
                                                          const float* __restrict__ gamma, int P, float* __restrict__ dy3,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ double sh[2][256];
    double s = 0.0, x = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256) { s += psum[i]; x += pxh[i]; }
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = x;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
        __syncthreads();
    }
    const float k1 = (float)(sh[0][0] / P), k2 = (float)(sh[1][0] / P);
    if (blockIdx.x == 0 && threadIdx.x == 0) { dgamma[0] = (float)sh[1][0]; dbeta[0] = (float)sh[0][0]; }
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p < P) dy3[p] = gamma[0] * stat3[3] * (dz3[p] - k1 - (y3[p] - stat3[2]) * stat3[3] * k2);
}

// corr_atte.5 (256 -> 1) and corr_atte.3/.4 (BN 256 + ReLU) backward, reduce phase.   grid (P/64); warp per pixel row
//   Z2 = relu(a2 Y2 + c2);  dw3 += dy3 Z2;  dA2 = dy3 w3 [Z2>0];  sums of dA2 and dA2 * xhat2
__global__ void __launch_bounds__(256) gce_bwd_y2_reduce_kernel(const float* __restrict__ y2, const float* __restrict__ stat2, const float* __restrict__ w3,
                                                                const float* __restrict__ dy3, int P, float* __restrict__ pw3,
                                                                float* __restrict__ psum, float* __restrict__ pxh) {
    __shared__ float red[3][8][HMID];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float a[8], c[8], w[8], mean[8], rstd[8];
    load8(stat2 + lane * 8, a); load8(stat2 + HMID + lane * 8, c); load8(stat2 + 2 * HMID + lane * 8, mean);
    load8(stat2 + 3 * HMID + lane * 8, rstd); load8(w3 + lane * 8, w);
    float aw[8] = {0, 0, 0, 0, 0, 0, 0, 0}, as[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ax[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 8; ++k) {
        const int p = blockIdx.x * 64 + warp * 8 + k;
        if (p >= P) break;
        float v[8];
        load8(y2 + (size_t)p * HMID + lane * 8, v);
        const float d = dy3[p];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float zv = fmaxf(a[i] * v[i] + c[i], 0.f);
            aw[i] += d * zv;
            const float g = zv > 0.f ? d * w[i] : 0.f;
            as[i] += g;
            ax[i] += g * (v[i] - mean[i]) * rstd[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[0][warp][lane * 8 + i] = aw[i]; red[1][warp][lane * 8 + i] = as[i]; red[2][warp][lane * 8 + i] = ax[i]; }
    __syncthreads();
    {
        const int cidx = threadIdx.x;   // 256 threads == HMID columns
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) { s0 += red[0][wv][cidx]; s1 += red[1][wv][cidx]; s2 += red[2][wv][cidx]; }
        pw3[(size_t)blockIdx.x * HMID + cidx] = s0;
        psum[(size_t)blockIdx.x * HMID + cidx] = s1;
        pxh[(size_t)blockIdx.x * HMID + cidx] = s2;
    }
}

// finalize d w3, BN(256) grads and the apply coefficients.   grid (HMID/32), block (32, 8)
__global__ void __launch_bounds__(256) gce_bwd_y2_finalize_kernel(const float* __restrict__ pw3, const float* __restrict__ psum,
                                                                  const float* __restrict__ pxh, int nparts, double count,
                                                                  const float* __restrict__ gamma, const float* __restrict__ stat2,
                                                                  float* __restrict__ kcoef, float* __restrict__ dw3,
                                                                  float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ double sh[3][8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    double w = 0.0, s = 0.0, x = 0.0;
    for (int i = threadIdx.y; i < nparts; i += 8) { w += pw3[(size_t)i * HMID + c]; s += psum[(size_t)i * HMID + c]; x += pxh[(size_t)i * HMID + c]; }
    sh[0][threadIdx.y][threadIdx.x] = w; sh[1][threadIdx.y][threadIdx.x] = s; sh[2][threadIdx.y][threadIdx.x] = x;
    __syncthreads();
    if (threadIdx.y != 0) return;
    w = s = x = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { w += sh[0][i][threadIdx.x]; s += sh[1][i][threadIdx.x]; x += sh[2][i][threadIdx.x]; }
    dw3[c] = (float)w; dgamma[c] = (float)x; dbeta[c] = (float)s;
    kcoef[c] = gamma[c] * stat2[3 * HMID + c];
    kcoef[HMID + c] = (float)(s / count);
    kcoef[2 * HMID + c] = (float)(x / count);
}

// dY2 planes = k0 (dA2 - k1 - xhat2 k2)        grid (P/64); warp per pixel row
__global__ void __launch_bounds__(256) gce_bwd_y2_apply_kernel(const float* __restrict__ y2, const float* __restrict__ stat2, const float* __restrict__ w3,
                                                               const float* __restrict__ dy3, const float* __restrict__ kcoef, int P,
                                                               __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float a[8], c[8], w[8], mean[8], rstd[8], k0[8], k1[8], k2[8];
    load8(stat2 + lane * 8, a); load8(stat2 + HMID + lane * 8, c); load8(stat2 + 2 * HMID + lane * 8, mean);
    load8(stat2 + 3 * HMID + lane * 8, rstd); load8(w3 + lane * 8, w);
    load8(kcoef + lane * 8, k0); load8(kcoef + HMID + lane * 8, k1); load8(kcoef + 2 * HMID + lane * 8, k2);
    for (int k = 0; k < 8; ++k) {
        const int p = blockIdx.x * 64 + warp * 8 + k;
        if (p >= P) break;
        float v[8], o[8];
        load8(y2 + (size_t)p * HMID + lane * 8, v);
        const float d = dy3[p];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float g = (a[i] * v[i] + c[i] > 0.f) ? d * w[i] : 0.f;
            o[i] = k0[i] * (g - k1[i] - (v[i] - mean[i]) * rstd[i] * k2[i]);
        }
        const size_t off = (size_t)p * HMID + lane * 8;
        store8_planes(out_hi + off, out_lo + off, o);
    }
}

// d corr_atte.2.weight[j][k] = (sum_z g2[z][j][k]) * a1[k]     (BN folded into the weights in the forward;
// the colsum(dY2) (x) c1 term vanishes because train-mode BN backward output sums to zero over the batch)
__global__ void gce_bwd_w2_kernel(const float* __restrict__ g2, int nsplit, const float* __restrict__ stat1, float* __restrict__ dw2) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= HMID * HG) return;
    float s = 0.f;
    for (int z = 0; z < nsplit; ++z) s += g2[(size_t)z * HMID * HG + idx];
    dw2[idx] = s * stat1[idx % HG];
}

// BN(1024) of corr_atte.1 backward on dZ1 (fp32) with xhat from the Y1 planes: reduce over one frame.  grid (HG/64, N)
__global__ void __launch_bounds__(256) gce_bwd_bn1_reduce_kernel(const float* __restrict__ dz1, const __nv_bfloat16* __restrict__ yh,
                                                                 const __nv_bfloat16* __restrict__ yl, const float* __restrict__ stat1,
                                                                 float* __restrict__ psum, float* __restrict__ pxh) {
    __shared__ float red[32 * 65];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const Tile t;
    float mean[8], rstd[8], as[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ax[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    load8(stat1 + 2 * HG + c0 + t.cg * 8, mean); load8(stat1 + 3 * HG + c0 + t.cg * 8, rstd);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t off = ((size_t)n * HS + t.r0 + 32 * k) * HG + c0 + t.cg * 8;
        float g[8], y[8];
        load8(dz1 + off, g);
        load8_planes(yh + off, yl + off, y);
#pragma unroll
        for (int i = 0; i < 8; ++i) { as[i] += g[i]; ax[i] += g[i] * (y[i] - mean[i]) * rstd[i]; }
    }
    tile_colsum(as, red, psum + (size_t)n * HG + c0, t);
    tile_colsum(ax, red, pxh + (size_t)n * HG + c0, t);
}

__global__ void __launch_bounds__(256) gce_bwd_bn1_finalize_kernel(const float* __restrict__ psum, const float* __restrict__ pxh, int nparts,
                                                                   double count, const float* __restrict__ gamma,
                                                                   const float* __restrict__ stat1, float* __restrict__ kcoef,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ double sh[2][8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;          // grid (HG/32), block (32, 8)
    double s = 0.0, x = 0.0;
    for (int i = threadIdx.y; i < nparts; i += 8) { s += psum[(size_t)i * HG + c]; x += pxh[(size_t)i * HG + c]; }
    sh[0][threadIdx.y][threadIdx.x] = s; sh[1][threadIdx.y][threadIdx.x] = x;
    __syncthreads();
    if (threadIdx.y != 0) return;
    s = x = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += sh[0][i][threadIdx.x]; x += sh[1][i][threadIdx.x]; }
    dgamma[c] = (float)x; dbeta[c] = (float)s;
    kcoef[c] = gamma[c] * stat1[3 * HG + c];
    kcoef[HG + c] = (float)(s / count);
    kcoef[2 * HG + c] = (float)(x / count);
}

// dY1 planes + per-frame column sums (for the per-clip bias gradient).  grid (HG/64, N)
__global__ void __launch_bounds__(256) gce_bwd_bn1_apply_kernel(const float* __restrict__ dz1, const __nv_bfloat16* __restrict__ yh,
                                                                const __nv_bfloat16* __restrict__ yl, const float* __restrict__ stat1,
                                                                const float* __restrict__ kcoef, __nv_bfloat16* __restrict__ out_hi,
                                                                __nv_bfloat16* __restrict__ out_lo, float* __restrict__ frame_sum) {
    __shared__ float red[32 * 65];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const Tile t;
    float mean[8], rstd[8], k0[8], k1[8], k2[8], as[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    load8(stat1 + 2 * HG + c0 + t.cg * 8, mean); load8(stat1 + 3 * HG + c0 + t.cg * 8, rstd);
    load8(kcoef + c0 + t.cg * 8, k0); load8(kcoef + HG + c0 + t.cg * 8, k1); load8(kcoef + 2 * HG + c0 + t.cg * 8, k2);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t off = ((size_t)n * HS + t.r0 + 32 * k) * HG + c0 + t.cg * 8;
        float g[8], y[8], o[8];
        load8(dz1 + off, g);
        load8_planes(yh + off, yl + off, y);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            o[i] = k0[i] * (g[i] - k1[i] - (y[i] - mean[i]) * rstd[i] * k2[i]);
            as[i] += o[i];
        }
        store8_planes(out_hi + off, out_lo + off, o);
    }
    tile_colsum(as, red, frame_sum + (size_t)n * HG + c0, t);
}

// glo_fc.1 (BatchNorm1d over B samples) + ReLU backward; thread per channel
__global__ void gce_bwd_glo_bn_kernel(const float* __restrict__ dglo, const float* __restrict__ glo, const float* __restrict__ u,
                                      const float* __restrict__ stat, const float* __restrict__ gamma, int B, float* __restrict__ du,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dfc_bias) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= HG) return;
    const float mean = stat[2 * HG + k], rstd = stat[3 * HG + k];
    double s = 0.0, x = 0.0;
    for (int b = 0; b < B; ++b) {
        const float g = glo[(size_t)b * HG + k] > 0.f ? dglo[(size_t)b * HG + k] : 0.f;
        s += g;
        x += (double)g * ((u[(size_t)b * HG + k] - mean) * rstd);
    }
    dgamma[k] = (float)x; dbeta[k] = (float)s;
    const float k0 = gamma[k] * rstd, k1 = (float)(s / B), k2 = (float)(x / B);
    float bsum = 0.f;
    for (int b = 0; b < B; ++b) {
        const float g = glo[(size_t)b * HG + k] > 0.f ? dglo[(size_t)b * HG + k] : 0.f;
        const float v = k0 * (g - k1 - (u[(size_t)b * HG + k] - mean) * rstd * k2);
        du[(size_t)b * HG + k] = v;
        bsum += v;
    }
    dfc_bias[k] = bsum;
}

// Stand-alone TRL backward: d x_uncorr / d x_corr in pixel-major form (same sums as the gate kernel).  grid (C/64, N)
//   dxu_out = dZ[t][fwd] + dZ[T-1-t][bwd] + (dM0_fwd + dM0_bwd)/T ;   dxc (in place) += dgc/S
__global__ void __launch_bounds__(256) trl_bwd_collect_kernel(const float* __restrict__ dz, const float* __restrict__ dmem, float* __restrict__ dxc,
                                                              const float* __restrict__ dgc, int B, int T, float* __restrict__ dxu_out) {
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const int b = n / T, tt = n - b * T;
    const int R = B * HS;
    const Tile t;
    float gcv[8];
    load8(dgc + (size_t)n * HC + c0 + t.cg * 8, gcv);
    const float invT = 1.f / (float)T;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int s = t.r0 + 32 * k;
        const size_t col = c0 + t.cg * 8;
        const size_t rrow = (size_t)b * HS + s;
        const size_t p = (size_t)n * HS + s;
        float u[8], v[8], a[8];
        load8(dz + (((size_t)tt * 2 + 0) * R + rrow) * HC + col, u);
        load8(dz + (((size_t)(T - 1 - tt) * 2 + 1) * R + rrow) * HC + col, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] += v[i];
        load8(dz + ((size_t)0 * R + rrow) * HC + col, v);
        load8(dz + ((size_t)1 * R + rrow) * HC + col, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += a[i];
        load8(dmem + ((size_t)0 * R + rrow) * HC + col, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += a[i];
        load8(dmem + ((size_t)1 * R + rrow) * HC + col, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] += (v[i] + a[i]) * invT;
        store8(dxu_out + p * HC + col, u);
        load8(dxc + p * HC + col, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += gcv[i] * (1.f / HS);
        store8(dxc + p * HC + col, v);
    }
}

// dx (NCHW) = transpose(dX1 pixel-major) + dg[b][c] / (T*S)      grid (C/64, N)
__global__ void __launch_bounds__(256) pm_to_nchw_bias_kernel(const float* __restrict__ src, const float* __restrict__ dg, int T, float scale,
                                                              float* __restrict__ out) {
    __shared__ float tile[64][129];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const int b = n / T;
    const Tile t;
    float add[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (dg) load8(dg + (size_t)b * HC + c0 + t.cg * 8, add);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int s = t.r0 + 32 * k;
        float v[8];
        load8(src + ((size_t)n * HS + s) * HC + c0 + t.cg * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) tile[t.cg * 8 + i][s] = v[i] + add[i] * scale;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = warp * 8 + i;
        const float4 v = make_float4(tile[c][lane * 4], tile[c][lane * 4 + 1], tile[c][lane * 4 + 2], tile[c][lane * 4 + 3]);
        *reinterpret_cast<float4*>(out + ((size_t)n * HC + c0 + c) * HS + lane * 4) = v;
    }
}

// ------------------------------------------------------------------ host helpers
static int bn_backward(grl_handle* h, cudaStream_t st, const HeadWs& w, const float* srcA, const float* srcB, const __nv_bfloat16* mask_hi,
                       const float* hraw, const float* stat, int Cn, const float* gamma0, const float* gamma1, float* dgamma0,
                       float* dgamma1, float* dbeta0, float* dbeta1, int accumulate, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
                       float* g_out, float* scal = nullptr, __half* out16 = nullptr, bool reduced = false) {
    // scal (4 floats, zeroed by the caller): max |g|, max |xhat|, bound of |dH|, 1 / scale of the fp16 copy `out16`
    // reduced: the GEMM that produced srcA already left the partial sums (4 per 128-row tile) in part_a / part_b and the two
    // maxima in scal (bn_reduce_in_epilogue below) -- the reduce pass over srcA, the mask and the raw activations is skipped
    const int R = w.R, B = w.B;
    dim3 grid(Cn / 64, R / 128, 2);
    unsigned int* su = reinterpret_cast<unsigned int*>(scal);
    if (!reduced) {
        bn_bwd_reduce_kernel<<<grid, 256, 0, st>>>(srcA, srcB, mask_hi, hraw, stat, Cn, R, WS_F32(w, part_a), WS_F32(w, part_b), su);
        GRL_LAUNCH_CHECK(h);
    }
    BnPtrs2 bp; bp.gamma[0] = gamma0; bp.gamma[1] = gamma1;
    OutPtrs2 dg, db; dg.p[0] = dgamma0; dg.p[1] = dgamma1; db.p[0] = dbeta0; db.p[1] = dbeta1;
    bn_bwd_finalize_kernel<<<dim3((Cn + 31) / 32, 2), dim3(32, 8), 0, st>>>(WS_F32(w, part_a), WS_F32(w, part_b), reduced ? 4 * (R / 128) : B, Cn, (double)R, bp, stat,
                                                                      WS_F32(w, kcoef), dg, db, accumulate, su);
    GRL_LAUNCH_CHECK(h);
    bn_bwd_apply_kernel<<<grid, 256, 0, st>>>(srcA, srcB, mask_hi, hraw, stat, WS_F32(w, kcoef), Cn, R, out_hi, out_lo, g_out, scal, out16);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

// The dgrad GEMM that produces the gradient of a BatchNorm + ReLU output [2][R][Cn] also leaves the BatchNorm-backward partial
// sums (sum g, sum g xhat per 32-row slab and channel) and the two maxima: bn_backward(..., reduced = true) follows.
static void bn_reduce_in_epilogue(GemmEpi& e, const HeadWs& w, const __nv_bfloat16* mask_hi, const float* hraw, const float* stat, int Cn, float* scal) {
    e.col_sum = WS_F32(w, part_a); e.col_sq = WS_F32(w, part_b);
    e.stat_bstride = (long long)4 * (w.R / 128) * Cn;
    e.bnb_mask = mask_hi; e.bnb_hraw = hraw; e.bnb_stat = stat;
    e.bnb_ld = Cn; e.bnb_bstride = (long long)w.R * Cn;
    e.bnb_scal = reinterpret_cast<unsigned int*>(scal);
}

static int outer(grl_handle* h, cudaStream_t st, const float* A, long long a_os, long long a_is, const float* Bm, long long b_os,
                 long long b_is, int no, int ni, float* out, long long ldo, int J, int K) {
    small_outer_kernel<<<dim3(K / 64, J / 32), 256, 0, st>>>(A, a_os, a_is, Bm, b_os, b_is, no, ni, out, ldo, 1.f);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

}  // namespace grl

using namespace grl;

namespace grl {

// TRL backward: leaves dZ[T][2][R][C] (in `dxu`), dmem (= dF1_0 Wf1), dxc (= dF2cat Wf2cat) and dgc in the workspace
static int trl_backward_part(grl_handle* h, cudaStream_t st, const grl_head_params* p, const HeadWs& w, const float* d_f_uncorr,
                             const float* d_f_corr, const grl_head_grads* g) {
    const int B = w.B, T = w.T, N = w.N, P = w.P, R = w.R;
    const size_t slotM = (size_t)2 * R * HC, slotB = (size_t)2 * R * HB;
    float* dz_all = WS_F32(w, dxu);                 // dZ[T][2][R][HC]
    float* dmem_all = WS_F32(w, dmem);              // [T+1][2][R][HC]: slot i = dF1_i Wf1, slot T = d f_uncorr / S broadcast
    // Two streams.  The caller's stream `st` carries only the recurrence's critical path: per step BN3' -> conv3 dgrad -> BN2' ->
    // conv2 dgrad -> BN1' -> conv1 dgrad (HBM-bound BN passes + three small GEMMs).  Everything else -- the f1 / f2
    // reciprocal-attention path (its gradients depend on saved activations only) and every weight gradient -- runs on the
    // low-priority side stream `sd` and fills the tensor pipe while the chain's elementwise kernels stream through HBM.
    //   EV_DMEM(i): dF1_i Wf1 ready (side -> main);  EV_DH3/2/1(i): dH planes of step i ready (main -> side, for the wgrads)
    cudaStream_t sd = (h->overlap & 2) ? h->side : st;
    const bool two = sd != st;
    auto EV_DMEM = [&](int i) { return 8 + i; };
    auto EV_DH3 = [&](int i) { return 8 + (T + 1) + i; };
    auto EV_DH2 = [&](int i) { return 8 + 2 * (T + 1) + i; };
    auto EV_DH1 = [&](int i) { return 8 + 3 * (T + 1) + i; };

    // ---------------- squeeze-excite backward for every step (independent of the recurrence) ----------------
    {
        SePtrsB sp; sp.l1[0] = p->se1_w[0]; sp.l1[1] = p->se1_w[1]; sp.l2[0] = p->se2_w[0]; sp.l2[1] = p->se2_w[1];
        GRL_CUDA(h, cudaMemsetAsync(WS_F32(w, f16_scal), 0, (size_t)(64 + 12 * T) * 4, st));   // [0] bound of |dF1|, |dF2| (bits), [1] 1 / their scale; ...
        se_bwd_kernel<<<dim3(B, 2, T), 256, 0, st>>>(d_f_corr, WS_F32(w, gc), WS_F32(w, se_a), WS_F32(w, se_h), sp, B, T, WS_F32(w, se_ds),
                                                     WS_F32(w, se_dh), WS_F32(w, se_dq), WS_F32(w, se_q),
                                                     reinterpret_cast<unsigned int*>(WS_F32(w, f16_scal)));
        GRL_LAUNCH_CHECK(h);
        dgc_kernel<<<(unsigned)(((size_t)N * HC + 255) / 256), 256, 0, st>>>(d_f_corr, WS_F32(w, se_a), B, T, WS_F32(w, dgc));
        GRL_LAUNCH_CHECK(h);
    }
    if (two) GRL_TRY(stream_wait(h, st, sd, 0));
    for (int d = 0; d < 2; ++d) {
        // d L2[c][j] = sum_{i,b} ds[c] h[j];   d L1[j][c] = sum_{i,b} dh[j] q[c]
        GRL_TRY(outer(h, sd, WS_F32(w, se_ds) + (size_t)d * B * HC, (long long)2 * B * HC, HC, WS_F32(w, se_h) + (size_t)d * B * HSE,
                      (long long)2 * B * HSE, HSE, T, B, g->se2_w[d], HSE, HC, HSE));
        GRL_TRY(outer(h, sd, WS_F32(w, se_dh) + (size_t)d * B * HSE, (long long)2 * B * HSE, HSE, WS_F32(w, se_q) + (size_t)d * B * HC,
                      (long long)2 * B * HC, HC, T, B, g->se1_w[d], HC, HSE, HC));
    }

    // reciprocal-attention path of step i: dF1 (planes), dF2[tau] (planes), bias partials, f1 wgrad, f1 dgrad -> dmem[i]
    auto f1_path = [&](int i) -> int {
        const int acc = (i == T - 1) ? 0 : 1;
        const int tau0 = i, tau1 = T - 1 - i;
        // dF1 / dF2[tau] come out as ONE fp16 plane each (scale from se_bwd_kernel's bound); the memory slot M_i was written as an
        // fp16 plane by the forward: the single-pass operands of this step's gradient GEMMs
        float* scal = WS_F32(w, f16_scal);
        __half* df1_16 = reinterpret_cast<__half*>(WS_BF(w, df1_16));
        const __half* mem_16 = reinterpret_cast<const __half*>(WS_BF(w, mem_16)) + (size_t)i * slotM;
        trl_bwd_f1_kernel<<<dim3(HC / 64, B, 2), 256, 0, sd>>>(WS_F32(w, f1) + (size_t)i * slotM, WS_F32(w, f2),
                                                               WS_F32(w, se_dq) + (size_t)i * 2 * B * HC, B, T, R, tau0, tau1, scal, df1_16,
                                                               reinterpret_cast<__half*>(WS_BF(w, df2_16)),
                                                               WS_F32(w, dbf1_part) + (size_t)i * 2 * B * HC,
                                                               WS_F32(w, dbf2_part) + (size_t)i * 2 * B * HC);
        GRL_LAUNCH_CHECK(h);
        {   // f1 wgrad: gw_f1[z] (+)= dF1^T M
            GemmEpi e = epi_default();
            e.C = WS_F32(w, gw_f1); e.ldc = HC; e.c_bstride = (long long)HC * HC; e.accumulate = acc;
            e.dscale_a = scal + 1;                       // the memory planes are unscaled fp16
            GRL_TRY(gemm_launch_x1(h, sd, HC, HC, R, 2, df1_16, HC, (long long)R * HC, 1, mem_16, HC, (long long)R * HC, 1, e));
        }
        {   // f1 dgrad: dmem[i] = dF1 Wf1   (the other half of dM for step i-1 is dZ of step i)
            GemmEpi e = epi_default();
            e.C = dmem_all + (size_t)i * slotM; e.ldc = HC; e.c_bstride = (long long)R * HC;
            e.dscale_a = scal + 1; e.dscale_b = scal + 5;
            GRL_TRY(gemm_launch_x1(h, sd, R, HC, HC, 2, df1_16, HC, (long long)R * HC, 0, reinterpret_cast<const __half*>(WS_BF(w, wf1_16)), HC,
                                   (long long)HC * HC, 1, e));
        }
        if (two) GRL_TRY(ev_record(h, EV_DMEM(i), sd));
        return GRL_OK;
    };

    // ---------------- BPTT over the memory updates ----------------
    bcast_rows_kernel<<<dim3(HC / 64, B, 2), 256, 0, st>>>(d_f_uncorr, 1.f / HS, R, dmem_all + (size_t)T * slotM);   // d f_uncorr = mean_s(M_fwd) + mean_s(M_bwd)
    GRL_LAUNCH_CHECK(h);
    // the f1 weights as one fp16 plane (input gradient of f1: dM += dF1 Wf1), converted once per backward on the side stream
    GRL_TRY(planes_to_f16(h, sd, WS_BF(w, wf1_hi), WS_BF(w, wf1_lo), (size_t)2 * HC * HC, WS_F32(w, f16_scal) + 4,
                          reinterpret_cast<__half*>(WS_BF(w, wf1_16))));
    GRL_TRY(f1_path(T - 1));
    const bool fuse_bn = (h->overlap & 128) == 0;      // debug bit 7: separate bn_bwd_reduce passes for bn1 / bn2 (A/B)
    for (int i = T - 1; i >= 0; --i) {
        const int first = (i == T - 1) ? 1 : 0;
        const int acc = first ? 0 : 1;
        if (i >= 1) GRL_TRY(f1_path(i - 1));        // side stream runs one step ahead of the chain
        float* dz = dz_all + (size_t)i * slotM;
        const float* dz_prev = first ? nullptr : dz_all + (size_t)(i + 1) * slotM;
        const float* dmem_in = dmem_all + (size_t)(i + 1) * slotM;
        const __nv_bfloat16* memn_hi = WS_BF(w, mem_hi) + (size_t)(i + 1) * slotM;
        const __nv_bfloat16 *z_hi = WS_BF(w, z_hi) + (size_t)i * slotM, *z_lo = WS_BF(w, z_lo) + (size_t)i * slotM;
        const float* h1 = WS_F32(w, h1) + (size_t)i * slotB;
        const float* h2 = WS_F32(w, h2) + (size_t)i * slotB;
        const float* h3 = WS_F32(w, h3) + (size_t)i * slotM;
        const __nv_bfloat16 *h1p_hi = WS_BF(w, h1p_hi) + (size_t)i * slotB, *h1p_lo = WS_BF(w, h1p_lo) + (size_t)i * slotB;
        const __nv_bfloat16 *h2p_hi = WS_BF(w, h2p_hi) + (size_t)i * slotB, *h2p_lo = WS_BF(w, h2p_lo) + (size_t)i * slotB;
        const float* s1 = WS_F32(w, sbn1) + (size_t)i * 2 * 4 * HB;
        const float* s2 = WS_F32(w, sbn2) + (size_t)i * 2 * 4 * HB;
        const float* s3 = WS_F32(w, sbn3) + (size_t)i * 2 * 4 * HC;
        __nv_bfloat16 *dh3_hi = WS_BF(w, dh3_hi) + (size_t)i * slotM, *dh3_lo = WS_BF(w, dh3_lo) + (size_t)i * slotM;
        __nv_bfloat16 *dh2_hi = WS_BF(w, dh2_hi) + (size_t)i * slotB, *dh2_lo = WS_BF(w, dh2_lo) + (size_t)i * slotB;
        __nv_bfloat16 *dh1_hi = WS_BF(w, dh1_hi) + (size_t)i * slotB, *dh1_lo = WS_BF(w, dh1_lo) + (size_t)i * slotB;
        // single-pass fp16 operands of this step's three weight-gradient GEMMs (leaf sums, no error propagation):
        // dH as scaled fp16 from bn_bwd_apply, the activations as fp16 planes from the forward; 4 scalars per BatchNorm and step
        __half* dh3_16 = reinterpret_cast<__half*>(WS_BF(w, dh3_16)) + (size_t)i * slotM;
        __half* dh2_16 = reinterpret_cast<__half*>(WS_BF(w, dh2_16)) + (size_t)i * slotB;
        __half* dh1_16 = reinterpret_cast<__half*>(WS_BF(w, dh1_16)) + (size_t)i * slotB;
        const __half* h2p_16 = reinterpret_cast<const __half*>(WS_BF(w, h2p_16)) + (size_t)i * slotB;
        const __half* h1p_16 = reinterpret_cast<const __half*>(WS_BF(w, h1p_16)) + (size_t)i * slotB;
        const __half* z_16 = reinterpret_cast<const __half*>(WS_BF(w, z_16)) + (size_t)i * slotM;
        float* sc3 = WS_F32(w, f16_scal) + 64 + (size_t)(i * 3 + 0) * 4;
        float* sc2 = WS_F32(w, f16_scal) + 64 + (size_t)(i * 3 + 1) * 4;
        float* sc1 = WS_F32(w, f16_scal) + 64 + (size_t)(i * 3 + 2) * 4;

        // ---- critical path (caller's stream) ----
        if (two && !first) GRL_TRY(ev_wait(h, EV_DMEM(i + 1), st));
        // bn3 + residual ReLU:  dPre = dMn [Mn>0] -> dz (fp32);  dH3 planes
        GRL_TRY(bn_backward(h, st, w, dmem_in, dz_prev, memn_hi, h3, s3, HC, p->memo_bn3[0].weight, p->memo_bn3[1].weight, g->memo_bn3_w[0],
                            g->memo_bn3_w[1], g->memo_bn3_b[0], g->memo_bn3_b[1], acc, dh3_hi, dh3_lo, dz, sc3, dh3_16));
        if (two) GRL_TRY(ev_record(h, EV_DH3(i), st));
        {   // conv3 dgrad: dH2p = dH3 Wc3
            GemmEpi e = epi_default();
            e.C = WS_F32(w, dh2p); e.ldc = HB; e.c_bstride = (long long)R * HB;
            Operand a{dh3_hi, dh3_lo, HC, (long long)R * HC, 0}, b{WS_BF(w, wc3_hi), WS_BF(w, wc3_lo), HB, (long long)HC * HB, 1};
            if (fuse_bn) bn_reduce_in_epilogue(e, w, h2p_hi, h2, s2, HB, sc2);
            GRL_TRY(gemm_launch(h, st, R, HB, HC, 2, a, b, e, 0));
        }
        GRL_TRY(bn_backward(h, st, w, WS_F32(w, dh2p), nullptr, h2p_hi, h2, s2, HB, p->memo_bn2[0].weight, p->memo_bn2[1].weight,
                            g->memo_bn2_w[0], g->memo_bn2_w[1], g->memo_bn2_b[0], g->memo_bn2_b[1], acc, dh2_hi, dh2_lo, nullptr, sc2, dh2_16, fuse_bn));
        if (two) GRL_TRY(ev_record(h, EV_DH2(i), st));
        {   // conv2 dgrad
            GemmEpi e = epi_default();
            e.C = WS_F32(w, dh1p); e.ldc = HB; e.c_bstride = (long long)R * HB;
            Operand a{dh2_hi, dh2_lo, HB, (long long)R * HB, 0}, b{WS_BF(w, wc2_hi), WS_BF(w, wc2_lo), HB, (long long)HB * HB, 1};
            if (fuse_bn) bn_reduce_in_epilogue(e, w, h1p_hi, h1, s1, HB, sc1);
            GRL_TRY(gemm_launch(h, st, R, HB, HB, 2, a, b, e, 0));
        }
        GRL_TRY(bn_backward(h, st, w, WS_F32(w, dh1p), nullptr, h1p_hi, h1, s1, HB, p->memo_bn1[0].weight, p->memo_bn1[1].weight,
                            g->memo_bn1_w[0], g->memo_bn1_w[1], g->memo_bn1_b[0], g->memo_bn1_b[1], acc, dh1_hi, dh1_lo, nullptr, sc1, dh1_16, fuse_bn));
        if (two) GRL_TRY(ev_record(h, EV_DH1(i), st));
        {   // conv1 dgrad: dZ = dPre + dH1 Wc1   (accumulates onto dPre)
            GemmEpi e = epi_default();
            e.C = dz; e.ldc = HC; e.c_bstride = (long long)R * HC; e.accumulate = 1;
            Operand a{dh1_hi, dh1_lo, HB, (long long)R * HB, 0}, b{WS_BF(w, wc1_hi), WS_BF(w, wc1_lo), HC, (long long)HB * HC, 1};
            GRL_TRY(gemm_launch(h, st, R, HC, HB, 2, a, b, e, 0));
        }

        // ---- weight gradients of the memory block (side stream) ----
        if (two) GRL_TRY(ev_wait(h, EV_DH3(i), sd));
        {   // conv3 wgrad: gw_c3[z] (+)= dH3^T H2p
            GemmEpi e = epi_default();
            e.C = WS_F32(w, gw_c3); e.ldc = HB; e.c_bstride = (long long)HC * HB; e.accumulate = acc;
            e.dscale_a = sc3 + 3;
            GRL_TRY(gemm_launch_x1(h, sd, HC, HB, R, 2, dh3_16, HC, (long long)R * HC, 1, h2p_16, HB, (long long)R * HB, 1, e));
        }
        if (two) GRL_TRY(ev_wait(h, EV_DH2(i), sd));
        {   // conv2 wgrad
            GemmEpi e = epi_default();
            e.C = WS_F32(w, gw_c2); e.ldc = HB; e.c_bstride = (long long)HB * HB; e.accumulate = acc;
            e.dscale_a = sc2 + 3;
            GRL_TRY(gemm_launch_x1(h, sd, HB, HB, R, 2, dh2_16, HB, (long long)R * HB, 1, h1p_16, HB, (long long)R * HB, 1, e));
        }
        if (two) GRL_TRY(ev_wait(h, EV_DH1(i), sd));
        {   // conv1 wgrad: gw_c1[z] (+)= dH1^T Z
            GemmEpi e = epi_default();
            e.C = WS_F32(w, gw_c1); e.ldc = HC; e.c_bstride = (long long)HB * HC; e.accumulate = acc;
            e.dscale_a = sc1 + 3;
            GRL_TRY(gemm_launch_x1(h, sd, HB, HC, R, 2, dh1_16, HB, (long long)R * HB, 1, z_16, HC, (long long)R * HC, 1, e));
        }
    }
    // ---------------- parameter gradients accumulated over the steps; f2 over all frames (side stream) ----------------
    for (int d = 0; d < 2; ++d) {
        GRL_CUDA(h, cudaMemcpyAsync(g->memo_conv3_w[d], WS_F32(w, gw_c3) + (size_t)d * HC * HB, (size_t)HC * HB * 4, cudaMemcpyDeviceToDevice, sd));
        GRL_CUDA(h, cudaMemcpyAsync(g->memo_conv2_w[d], WS_F32(w, gw_c2) + (size_t)d * HB * HB, (size_t)HB * HB * 4, cudaMemcpyDeviceToDevice, sd));
        GRL_CUDA(h, cudaMemcpyAsync(g->memo_conv1_w[d], WS_F32(w, gw_c1) + (size_t)d * HB * HC, (size_t)HB * HC * 4, cudaMemcpyDeviceToDevice, sd));
        GRL_CUDA(h, cudaMemcpyAsync(g->f1_w[d], WS_F32(w, gw_f1) + (size_t)d * HC * HC, (size_t)HC * HC * 4, cudaMemcpyDeviceToDevice, sd));
        small_colsum_kernel<<<dim3(HC / 256, 1), 256, 0, sd>>>(WS_F32(w, dbf1_part) + (size_t)d * B * HC, 0, (long long)2 * B * HC, HC, T, B,
                                                               g->f1_b[d], 0, HC);
        GRL_LAUNCH_CHECK(h);
        small_colsum_kernel<<<dim3(HC / 256, 1), 256, 0, sd>>>(WS_F32(w, dbf2_part) + (size_t)d * B * HC, 0, (long long)2 * B * HC, HC, T, B,
                                                               g->f2_b[d], 0, HC);
        GRL_LAUNCH_CHECK(h);
        {   // f2 wgrad over all frames: d Wf2[d] = dF2[:, d]^T Xc     (K = P); single-pass fp16 operands (converted below, d == 0)
            if (d == 0) {
                float* scal = WS_F32(w, f16_scal);
                GRL_TRY(planes_to_f16(h, sd, WS_BF(w, xc_hi), WS_BF(w, xc_lo), (size_t)P * HC, scal + 8, reinterpret_cast<__half*>(WS_BF(w, xc_16))));
                GRL_TRY(planes_to_f16(h, sd, WS_BF(w, wf2_hi), WS_BF(w, wf2_lo), (size_t)2 * HC * HC, scal + 10, reinterpret_cast<__half*>(WS_BF(w, wf2_16))));
            }
            GemmEpi e = epi_default();
            e.C = g->f2_w[d]; e.ldc = HC;
            e.dscale_a = WS_F32(w, f16_scal) + 1; e.dscale_b = WS_F32(w, f16_scal) + 9;
            GRL_TRY(gemm_launch_x1(h, sd, HC, HC, P, 1, reinterpret_cast<const __half*>(WS_BF(w, df2_16)) + (size_t)d * HC, 2 * HC, 0, 1,
                                   reinterpret_cast<const __half*>(WS_BF(w, xc_16)), HC, 0, 1, e));
        }
    }
    {   // f2 dgrad, both directions in one contraction (K = 4096): dxc = dF2cat Wf2cat
        GemmEpi e = epi_default();
        e.C = WS_F32(w, dxc); e.ldc = HC;
        e.dscale_a = WS_F32(w, f16_scal) + 1; e.dscale_b = WS_F32(w, f16_scal) + 11;
        GRL_TRY(gemm_launch_x1(h, sd, P, HC, 2 * HC, 1, reinterpret_cast<const __half*>(WS_BF(w, df2_16)), 2 * HC, 0, 0,
                               reinterpret_cast<const __half*>(WS_BF(w, wf2_16)), HC, 0, 1, e));
    }
    if (two) GRL_TRY(stream_wait(h, sd, st, 1));
    return GRL_OK;
}

// GCE backward.  use_trl: take d x_uncorr / d x_corr from the TRL backward state; dxu_extra: extra [P][C] gradient on x_uncorr
static int gce_backward_part(grl_handle* h, cudaStream_t st, const grl_head_params* p, const HeadWs& w, int use_trl,
                             const float* d_corr_map, const float* dxu_extra, float* dx, const grl_head_grads* g) {
    const int B = w.B, T = w.T, N = w.N, P = w.P;
    float* dz_all = WS_F32(w, dxu);
    float* dmem = WS_F32(w, dmem);
    // ---------------- GCE backward ----------------
    float* gs = WS_F32(w, gsmall);
    float* k2coef = gs;                 // [3][HMID]
    float* k1coef = gs + 1024;          // [3][HG]
    gce_bwd_gate_kernel<<<dim3(HC / 64, N), 256, 0, st>>>(dz_all, dmem, WS_F32(w, dxc), WS_F32(w, dgc), WS_BF(w, xp_hi), WS_BF(w, xp_lo),
                                                          WS_F32(w, m), B, T, dxu_extra, use_trl, WS_F32(w, part_c));
    GRL_LAUNCH_CHECK(h);
    const int pblocks = (P + 255) / 256;
    gce_bwd_dm_kernel<<<pblocks, 256, 0, st>>>(WS_F32(w, part_c), HC / 64, d_corr_map, WS_F32(w, m), WS_F32(w, y3), WS_F32(w, bn3_stat), P,
                                               WS_F32(w, dm), WS_F32(w, part_a), WS_F32(w, part_b));
    GRL_LAUNCH_CHECK(h);
    gce_bwd_dy3_kernel<<<pblocks, 256, 0, st>>>(WS_F32(w, dm), WS_F32(w, part_a), WS_F32(w, part_b), pblocks, WS_F32(w, y3), WS_F32(w, bn3_stat),
                                                p->atte_bn6.weight, P, WS_F32(w, dy3), g->atte_bn6_w, g->atte_bn6_b);
    GRL_LAUNCH_CHECK(h);
    const int yblocks = (P + 63) / 64;
    gce_bwd_y2_reduce_kernel<<<yblocks, 256, 0, st>>>(WS_F32(w, y2), WS_F32(w, bn2_stat), p->atte5_w, WS_F32(w, dy3), P, WS_F32(w, part_a),
                                                      WS_F32(w, part_b), WS_F32(w, part_c));
    GRL_LAUNCH_CHECK(h);
    gce_bwd_y2_finalize_kernel<<<HMID / 32, dim3(32, 8), 0, st>>>(WS_F32(w, part_a), WS_F32(w, part_b), WS_F32(w, part_c), yblocks, (double)P, p->atte_bn3.weight,
                                                   WS_F32(w, bn2_stat), k2coef, g->atte5_w, g->atte_bn3_w, g->atte_bn3_b);
    GRL_LAUNCH_CHECK(h);
    gce_bwd_y2_apply_kernel<<<yblocks, 256, 0, st>>>(WS_F32(w, y2), WS_F32(w, bn2_stat), p->atte5_w, WS_F32(w, dy3), k2coef, P, WS_BF(w, dy2_hi),
                                                     WS_BF(w, dy2_lo));
    GRL_LAUNCH_CHECK(h);
    {   // corr_atte.2 wgrad (on the raw Y1; BN scale applied afterwards), split-K over frames so the 16 output tiles fill the SMs
        int nsplit = 1;
        while (nsplit < 16 && (N % (nsplit * 2)) == 0) nsplit *= 2;
        const long long krows = (long long)P / nsplit;
        GemmEpi e = epi_default();
        e.C = WS_F32(w, g2); e.ldc = HG; e.c_bstride = (long long)HMID * HG;
        Operand a{WS_BF(w, dy2_hi), WS_BF(w, dy2_lo), HMID, krows * HMID, 1}, b{WS_BF(w, y1_hi), WS_BF(w, y1_lo), HG, krows * HG, 1};
        GRL_TRY(gemm_launch(h, st, HMID, HG, (int)krows, nsplit, a, b, e, 128));
        gce_bwd_w2_kernel<<<(HMID * HG + 255) / 256, 256, 0, st>>>(WS_F32(w, g2), nsplit, WS_F32(w, bn1_stat), g->atte2_w);
        GRL_LAUNCH_CHECK(h);
    }
    {   // corr_atte.2 dgrad: dZ1 = dY2 W2
        GemmEpi e = epi_default();
        e.C = WS_F32(w, dz1); e.ldc = HG;
        Operand a{WS_BF(w, dy2_hi), WS_BF(w, dy2_lo), HMID, 0, 0}, b{WS_BF(w, w2_hi), WS_BF(w, w2_lo), HG, 0, 1};
        GRL_TRY(gemm_launch(h, st, P, HG, HMID, 1, a, b, e, 0));
    }
    gce_bwd_bn1_reduce_kernel<<<dim3(HG / 64, N), 256, 0, st>>>(WS_F32(w, dz1), WS_BF(w, y1_hi), WS_BF(w, y1_lo), WS_F32(w, bn1_stat),
                                                                WS_F32(w, part_a), WS_F32(w, part_b));
    GRL_LAUNCH_CHECK(h);
    gce_bwd_bn1_finalize_kernel<<<HG / 32, dim3(32, 8), 0, st>>>(WS_F32(w, part_a), WS_F32(w, part_b), N, (double)P, p->atte_bn1.weight,
                                                                  WS_F32(w, bn1_stat), k1coef, g->atte_bn1_w, g->atte_bn1_b);
    GRL_LAUNCH_CHECK(h);
    gce_bwd_bn1_apply_kernel<<<dim3(HG / 64, N), 256, 0, st>>>(WS_F32(w, dz1), WS_BF(w, y1_hi), WS_BF(w, y1_lo), WS_F32(w, bn1_stat), k1coef,
                                                               WS_BF(w, dy1_hi), WS_BF(w, dy1_lo), WS_F32(w, dbias1_part));
    GRL_LAUNCH_CHECK(h);
    {   // corr_atte.0 wgrad (feature half): d W1[:, :2048] = dY1^T X, written straight into the [1024][3072] gradient
        GemmEpi e = epi_default();
        e.C = g->atte0_w; e.ldc = HC + HG;
        Operand a{WS_BF(w, dy1_hi), WS_BF(w, dy1_lo), HG, 0, 1}, b{WS_BF(w, xp_hi), WS_BF(w, xp_lo), HC, 0, 1};
        GRL_TRY(gemm_launch(h, st, HG, HC, P, 1, a, b, e, 0));
    }
    {   // corr_atte.0 dgrad: dX1 += dY1 W1a
        GemmEpi e = epi_default();
        e.C = WS_F32(w, dxc); e.ldc = HC; e.accumulate = 1;
        Operand a{WS_BF(w, dy1_hi), WS_BF(w, dy1_lo), HG, 0, 0}, b{WS_BF(w, w1a_hi), WS_BF(w, w1a_lo), HC, 0, 1};
        GRL_TRY(gemm_launch(h, st, P, HC, HG, 1, a, b, e, 0));
    }
    // global-descriptor branch (B rows)
    small_colsum_kernel<<<dim3(HG / 256, B), 256, 0, st>>>(WS_F32(w, dbias1_part), (long long)T * HG, 0, HG, 1, T, WS_F32(w, dbias1), HG, HG);
    GRL_LAUNCH_CHECK(h);
    GRL_TRY(outer(h, st, WS_F32(w, dbias1), 0, HG, WS_F32(w, glo), 0, HG, 1, B, g->atte0_w + HC, HC + HG, HG, HG));      // d W1[:, 2048:]
    float* dglo = WS_F32(w, part_c);      // [B][HG] scratch (the partial buffers are idle here)
    GRL_TRY(small_gemm(h, st, w, WS_F32(w, dbias1), B, HG, WS_BF(w, w1b_hi), WS_BF(w, w1b_lo), HG, 1, nullptr, dglo, HG));
    gce_bwd_glo_bn_kernel<<<(HG + 127) / 128, 128, 0, st>>>(dglo, WS_F32(w, glo), WS_F32(w, u), WS_F32(w, glo_stat), p->glo_bn.weight, B,
                                                            WS_F32(w, du), g->glo_bn_w, g->glo_bn_b, g->glo_fc_b);
    GRL_LAUNCH_CHECK(h);
    GRL_TRY(outer(h, st, WS_F32(w, du), 0, HG, WS_F32(w, g), 0, HC, 1, B, g->glo_fc_w, HC, HG, HC));
    GRL_TRY(small_gemm(h, st, w, WS_F32(w, du), B, HG, WS_BF(w, wg_hi), WS_BF(w, wg_lo), HC, 1, nullptr, WS_F32(w, dg), HC));
    pm_to_nchw_bias_kernel<<<dim3(HC / 64, N), 256, 0, st>>>(WS_F32(w, dxc), WS_F32(w, dg), T, 1.f / (float)(T * HS), dx);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

static int check_grads(grl_handle* h, const grl_head_grads* g, int which) {
    const float* const* gp = reinterpret_cast<const float* const*>(g);
    const size_t n_gce = offsetof(grl_head_grads, f1_w) / sizeof(float*), n_all = sizeof(grl_head_grads) / sizeof(float*);
    for (size_t i = (which & 1) ? 0 : n_gce; i < ((which & 2) ? n_all : n_gce); ++i)
        if (!gp[i]) return set_error(h, GRL_EINVAL, "grl_head backward: NULL gradient pointer (slot %zu)", i);
    return GRL_OK;
}

static int bwd_setup(grl_handle* h, const char* who, const grl_head_params* p, int B, int T, const grl_head_grads* g, void* workspace,
                     size_t workspace_bytes, int which, HeadWs* w) {
    if (!h || !p || !g || !workspace) return set_error(h, GRL_EINVAL, "%s: NULL argument", who);
    if (B < ((which & 1) ? 2 : 1) || T <= 0) return set_error(h, GRL_EINVAL, "%s: bad B / T", who);
    GRL_TRY(check_grads(h, g, which));
    *w = head_ws_layout(B, T, 1);
    if (workspace_bytes < w->total) return set_error(h, GRL_ENOMEM, "%s: workspace %zu < %zu bytes", who, workspace_bytes, w->total);
    if (reinterpret_cast<uintptr_t>(workspace) & 1023) return set_error(h, GRL_EINVAL, "%s: workspace must be 1024-byte aligned", who);
    w->base = (uint8_t*)workspace;
    return GRL_OK;
}

}  // namespace grl

extern "C" int grl_head_backward(grl_handle* h, const grl_head_params* p, const float* x, int B, int T, const float* d_f_uncorr,
                                 const float* d_f_corr, const float* d_x_uncorr, const float* d_x_corr, const float* d_corr_map,
                                 float* dx, const grl_head_grads* g, void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !d_f_uncorr || !d_f_corr || !dx) return set_error(h, GRL_EINVAL, "grl_head_backward: NULL argument");
    HeadWs w;
    GRL_TRY(bwd_setup(h, "grl_head_backward", p, B, T, g, workspace, workspace_bytes, 3, &w));
    cudaStream_t st = (cudaStream_t)stream;
    GRL_TRY(trl_backward_part(h, st, p, w, d_f_uncorr, d_f_corr, g));
    if (d_x_corr) { nchw_to_pm_kernel<<<dim3(HC / 64, w.N), 256, 0, st>>>(d_x_corr, WS_F32(w, dxc), 1); GRL_LAUNCH_CHECK(h); }
    float* dxu_extra = nullptr;       // the saved H3 activations are dead after the BPTT loop: reuse as [P][C] scratch
    if (d_x_uncorr) {
        dxu_extra = WS_F32(w, h3);
        nchw_to_pm_kernel<<<dim3(HC / 64, w.N), 256, 0, st>>>(d_x_uncorr, dxu_extra, 0);
        GRL_LAUNCH_CHECK(h);
    }
    return gce_backward_part(h, st, p, w, 1, d_corr_map, dxu_extra, dx, g);
}

extern "C" int grl_trl_backward(grl_handle* h, const grl_head_params* p, int B, int T, const float* d_f_uncorr, const float* d_f_corr,
                                float* d_x_uncorr, float* d_x_corr, const grl_head_grads* g, void* workspace, size_t workspace_bytes,
                                void* stream) {
    if (!d_f_uncorr || !d_f_corr || !d_x_uncorr || !d_x_corr) return set_error(h, GRL_EINVAL, "grl_trl_backward: NULL argument");
    HeadWs w;
    GRL_TRY(bwd_setup(h, "grl_trl_backward", p, B, T, g, workspace, workspace_bytes, 2, &w));
    cudaStream_t st = (cudaStream_t)stream;
    GRL_TRY(trl_backward_part(h, st, p, w, d_f_uncorr, d_f_corr, g));
    float* dxu_pm = WS_F32(w, h3);    // dead after the BPTT loop
    trl_bwd_collect_kernel<<<dim3(HC / 64, w.N), 256, 0, st>>>(WS_F32(w, dxu), WS_F32(w, dmem), WS_F32(w, dxc), WS_F32(w, dgc), B, T, dxu_pm);
    GRL_LAUNCH_CHECK(h);
    pm_to_nchw_bias_kernel<<<dim3(HC / 64, w.N), 256, 0, st>>>(dxu_pm, nullptr, T, 0.f, d_x_uncorr);
    GRL_LAUNCH_CHECK(h);
    pm_to_nchw_bias_kernel<<<dim3(HC / 64, w.N), 256, 0, st>>>(WS_F32(w, dxc), nullptr, T, 0.f, d_x_corr);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

extern "C" int grl_gce_backward(grl_handle* h, const grl_head_params* p, int B, int T, const float* d_x_uncorr, const float* d_x_corr,
                                const float* d_corr_map, float* dx, const grl_head_grads* g, void* workspace, size_t workspace_bytes,
                                void* stream) {
    if (!dx) return set_error(h, GRL_EINVAL, "grl_gce_backward: NULL argument");
    HeadWs w;
    GRL_TRY(bwd_setup(h, "grl_gce_backward", p, B, T, g, workspace, workspace_bytes, 1, &w));
    cudaStream_t st = (cudaStream_t)stream;
    if (d_x_corr) { nchw_to_pm_kernel<<<dim3(HC / 64, w.N), 256, 0, st>>>(d_x_corr, WS_F32(w, dxc), 0); GRL_LAUNCH_CHECK(h); }
    else GRL_CUDA(h, cudaMemsetAsync(WS_F32(w, dxc), 0, (size_t)w.P * HC * 4, st));
    float* dxu_extra = nullptr;
    if (d_x_uncorr) {
        dxu_extra = WS_F32(w, h3);
        nchw_to_pm_kernel<<<dim3(HC / 64, w.N), 256, 0, st>>>(d_x_uncorr, dxu_extra, 0);
        GRL_LAUNCH_CHECK(h);
    }
    return gce_backward_part(h, st, p, w, 0, d_corr_map, dxu_extra, dx, g);
}
