// grl_b200 — GCE + TRL head, forward (sm_100a).
//
// Replaces, on layer4 maps x [B*T][2048][16][8]:
//   Backbone.forward after self.base   /root/reference/reid/models/basebranch.py:56-68
//   TRLBlock.forward (+ BasicBlock)    /root/reference/reid/models/grl_model.py:131-180, 67-85
// following the decomposed plan of SURVEY.md §7.1 (F1 glo-as-bias, F2 batched f2, F3 gating as
// planes, F4 pooled shortcut).  Every 1x1 conv is the split-bf16 tcgen05 GEMM of gemm.cuh; the
// kernels in this file are the HBM-bound glue (layout change, BN finalisation, gating, pooling,
// squeeze-excite, memory update) written for coalesced 128-bit accesses.
#include "head_common.cuh"

namespace grl {

// ------------------------------------------------------------------ K1: NCHW -> pixel-major planes + per-frame sums
__global__ void __launch_bounds__(256) nchw_to_planes_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                             __nv_bfloat16* __restrict__ lo, float* __restrict__ gx, float gx_scale,
                                                             const float* __restrict__ x2 = nullptr) {
    __shared__ float tile[64][129];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = warp * 8 + i;
        float4 v = *reinterpret_cast<const float4*>(x + ((size_t)n * HC + c0 + c) * HS + lane * 4);
        if (x2) {                                     // BasicBlock.forward(x1, x2): the block works on x1 + x2
            const float4 u = *reinterpret_cast<const float4*>(x2 + ((size_t)n * HC + c0 + c) * HS + lane * 4);
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
        tile[c][lane * 4 + 0] = v.x; tile[c][lane * 4 + 1] = v.y; tile[c][lane * 4 + 2] = v.z; tile[c][lane * 4 + 3] = v.w;
        const float s = warp_sum(v.x + v.y + v.z + v.w);
        if (lane == 0 && gx) gx[(size_t)n * HC + c0 + c] = s * gx_scale;
    }
    __syncthreads();
    const Tile t;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int s = t.r0 + 32 * k;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = tile[t.cg * 8 + i][s];
        const size_t off = ((size_t)n * HS + s) * HC + c0 + t.cg * 8;
        store8_planes(hi + off, lo + off, v);
    }
}

// pixel-major planes -> NCHW fp32 (stand-alone x_corr / x_uncorr outputs)
__global__ void __launch_bounds__(256) planes_to_nchw_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                                             float* __restrict__ out) {
    __shared__ float tile[64][129];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const Tile t;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int s = t.r0 + 32 * k;
        float v[8];
        const size_t off = ((size_t)n * HS + s) * HC + c0 + t.cg * 8;
        load8_planes(hi + off, lo + off, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) tile[t.cg * 8 + i][s] = v[i];
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

// ------------------------------------------------------------------ K2: global descriptor branch (tiny)
__global__ void glo_mean_kernel(const float* __restrict__ gx, float* __restrict__ g, int B, int T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * HC) return;
    const int b = i / HC, c = i - b * HC;
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += gx[((size_t)b * T + t) * HC + c];
    g[i] = s / (float)(T * HS);
}

// out[b*ldo + j] = bias[j] + sum_k in[b*ldi + k] * W[j*ldw + k]   in plain fp32 FMAs (rows = B clips, K % 128 == 0).
// The global descriptor feeds BatchNorm1d over only B samples, which amplifies input error by |u| / |u_b - mean|, so this
// tiny product (B x 1024 x 2048) stays in IEEE fp32 instead of going through the split-bf16 tensor-core path.
// One warp per output column, all rows of a 32-row block in registers: 33 independent 128-bit loads per k-chunk.
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ in, long long ldi, const float* __restrict__ W,
                                                           long long ldw, const float* __restrict__ bias, float* __restrict__ out,
                                                           long long ldo, int rows, int K, int J) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= J) return;
    const int lane = lane_id();
    const float* wrow = W + (size_t)j * ldw;
    for (int b0 = 0; b0 < rows; b0 += 32) {
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        for (int k = lane * 4; k < K; k += 128) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow + k));
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (b0 + i < rows) {
                    const float4 xv = __ldg(reinterpret_cast<const float4*>(in + (size_t)(b0 + i) * ldi + k));
                    acc[i] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float s = warp_sum(acc[i]);
            if (lane == 0 && b0 + i < rows) out[(size_t)(b0 + i) * ldo + j] = s + (bias ? bias[j] : 0.f);
        }
    }
}

int small_gemm(grl_handle* h, cudaStream_t st, const HeadWs& w, const float* in, int rows, int K, const __nv_bfloat16* w_hi,
               const __nv_bfloat16* w_lo, long long ldw, int w_mn, const float* bias, float* out, int N) {
    GRL_TRY(split_planes(h, st, in, K, WS_BF(w, sa_hi), WS_BF(w, sa_lo), K, rows, K));
    GemmEpi e = epi_default();
    e.C = out; e.ldc = N;
    e.col_bias = bias;
    Operand a{WS_BF(w, sa_hi), WS_BF(w, sa_lo), K, 0, 0}, b{w_hi, w_lo, ldw, 0, w_mn};
    return gemm_launch(h, st, rows, N, K, 1, a, b, e, 128);
}

// BatchNorm1d over `rows` samples + ReLU; stat = [a | c | mean | rstd] per channel
__global__ void bn1d_relu_kernel(const float* __restrict__ u, float* __restrict__ out, int rows, int J,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ rmean,
                                 float* __restrict__ rvar, float* __restrict__ stat, int train) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J) return;
    double mean, var;
    if (train) {
        double s = 0.0, sq = 0.0;
        for (int b = 0; b < rows; ++b) { const double v = u[(size_t)b * J + j]; s += v; sq += v * v; }
        mean = s / rows;
        var = sq / rows - mean * mean;
        if (var < 0) var = 0;
        const double unb = rows > 1 ? var * rows / (rows - 1) : var;
        rmean[j] = (float)((1.0 - BN_MOM) * rmean[j] + BN_MOM * mean);
        rvar[j] = (float)((1.0 - BN_MOM) * rvar[j] + BN_MOM * unb);
    } else {
        mean = rmean[j]; var = rvar[j];
    }
    const double rstd = 1.0 / sqrt(var + (double)BN_EPS);
    const float a = (float)(gamma[j] * rstd), c = (float)(beta[j] - gamma[j] * rstd * mean);
    stat[j] = a; stat[J + j] = c; stat[2 * J + j] = (float)mean; stat[3 * J + j] = (float)rstd;
    for (int b = 0; b < rows; ++b) out[(size_t)b * J + j] = fmaxf(a * u[(size_t)b * J + j] + c, 0.f);
}

// ------------------------------------------------------------------ BN finalisation from GEMM-epilogue partials
struct BnPtrs { const float* gamma[2]; const float* beta[2]; float* rmean[2]; float* rvar[2]; };

// stat layout per z: [a | c | mean | rstd] x Cn.   grid (Cn/32, nz), block (32, NY <= 32): NY lanes share the partials of a channel
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ psum, const float* __restrict__ psq, int nparts,
                                                          long long part_bstride, int Cn, double count, BnPtrs bp,
                                                          float* __restrict__ stat, int train) {
    __shared__ double sh[2][32][33];
    const int z = blockIdx.y;
    const int c = blockIdx.x * 32 + threadIdx.x;
    const bool ok = c < Cn;
    const int NY = blockDim.y;
    if (train) {
        double s = 0.0, sq = 0.0;
        if (ok) {
            const float* ps = psum + z * part_bstride + c;
            const float* pq = psq + z * part_bstride + c;
            for (int i = threadIdx.y; i < nparts; i += NY) { s += ps[(size_t)i * Cn]; sq += pq[(size_t)i * Cn]; }
        }
        sh[0][threadIdx.y][threadIdx.x] = s; sh[1][threadIdx.y][threadIdx.x] = sq;
        __syncthreads();
    }
    if (threadIdx.y != 0 || !ok) return;
    double mean, var;
    if (train) {
        double s = 0.0, sq = 0.0;
        for (int i = 0; i < NY; ++i) { s += sh[0][i][threadIdx.x]; sq += sh[1][i][threadIdx.x]; }
        mean = s / count;
        var = sq / count - mean * mean;
        if (var < 0) var = 0;
        const double unb = count > 1 ? var * count / (count - 1) : var;
        bp.rmean[z][c] = (float)((1.0 - BN_MOM) * bp.rmean[z][c] + BN_MOM * mean);
        bp.rvar[z][c] = (float)((1.0 - BN_MOM) * bp.rvar[z][c] + BN_MOM * unb);
    } else {
        mean = bp.rmean[z][c]; var = bp.rvar[z][c];
    }
    const double rstd = 1.0 / sqrt(var + (double)BN_EPS);
    float* st = stat + (size_t)z * 4 * Cn;
    st[c] = (float)(bp.gamma[z][c] * rstd);
    st[Cn + c] = (float)(bp.beta[z][c] - bp.gamma[z][c] * rstd * mean);
    st[2 * Cn + c] = (float)mean;
    st[3 * Cn + c] = (float)rstd;
}

// corr_atte.1 (BN 1024) folded into corr_atte.2: w2s = W2 * diag(a1) as planes, bias2 = W2 c1
__global__ void fold_bn_into_w2_kernel(const float* __restrict__ w2, const float* __restrict__ stat1, __nv_bfloat16* __restrict__ hi,
                                       __nv_bfloat16* __restrict__ lo, float* __restrict__ bias2) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= HMID) return;
    const int lane = lane_id();
    float acc = 0.f;
    for (int k = lane; k < HG; k += 32) {
        const float w = w2[(size_t)j * HG + k];
        __nv_bfloat16 h, l;
        split_bf16(w * stat1[k], h, l);
        hi[(size_t)j * HG + k] = h; lo[(size_t)j * HG + k] = l;
        acc += w * stat1[HG + k];
    }
    acc = warp_sum(acc);
    if (lane == 0) bias2[j] = acc;
}

// ------------------------------------------------------------------ K6: y3 = w3 . relu(bn(Y2)), partial stats of y3
__global__ void __launch_bounds__(256) gce_y3_kernel(const float* __restrict__ y2, const float* __restrict__ stat2, const float* __restrict__ w3,
                                                     float* __restrict__ y3, float* __restrict__ psum, float* __restrict__ psq, int P) {
    __shared__ float s1[8], s2[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float a[8], c[8], w[8];
    load8(stat2 + lane * 8, a); load8(stat2 + HMID + lane * 8, c); load8(w3 + lane * 8, w);
    float bs = 0.f, bq = 0.f;
    for (int k = 0; k < 8; ++k) {
        const int p = blockIdx.x * 64 + warp * 8 + k;
        if (p >= P) break;
        float v[8];
        load8(y2 + (size_t)p * HMID + lane * 8, v);
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) d += w[i] * fmaxf(a[i] * v[i] + c[i], 0.f);
        d = warp_sum(d);
        if (lane == 0) y3[p] = d;
        bs += d; bq += d * d;
    }
    if (lane == 0) { s1[warp] = bs; s2[warp] = bq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a1 = 0.f, a2 = 0.f;
        for (int i = 0; i < 8; ++i) { a1 += s1[i]; a2 += s2[i]; }
        psum[blockIdx.x] = a1; psq[blockIdx.x] = a2;
    }
}

// scalar BN (corr_atte.6) finalisation + sigmoid -> corr_map m[p]; block 0 updates the running buffers
__global__ void __launch_bounds__(256) gce_m_kernel(const float* __restrict__ y3, const float* __restrict__ psum, const float* __restrict__ psq,
                                                    int nparts, int P, grl_bn_params bn, float* __restrict__ stat, float* __restrict__ m,
                                                    float* __restrict__ corr_map_out, int train) {
    __shared__ double sh[2][256];
    __shared__ float ac[2];
    if (train) {
        double s = 0.0, q = 0.0;
        for (int i = threadIdx.x; i < nparts; i += 256) { s += psum[i]; q += psq[i]; }
        sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = q;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        double mean, var;
        if (train) {
            mean = sh[0][0] / P;
            var = sh[1][0] / P - mean * mean;
            if (var < 0) var = 0;
        } else {
            mean = bn.running_mean[0]; var = bn.running_var[0];
        }
        const double rstd = 1.0 / sqrt(var + (double)BN_EPS);
        ac[0] = (float)(bn.weight[0] * rstd);
        ac[1] = (float)(bn.bias[0] - bn.weight[0] * rstd * mean);
        if (blockIdx.x == 0) {
            stat[0] = ac[0]; stat[1] = ac[1]; stat[2] = (float)mean; stat[3] = (float)rstd;
            if (train) {
                const double unb = P > 1 ? var * P / (P - 1.0) : var;
                bn.running_mean[0] = (float)((1.0 - BN_MOM) * bn.running_mean[0] + BN_MOM * mean);
                bn.running_var[0] = (float)((1.0 - BN_MOM) * bn.running_var[0] + BN_MOM * unb);
            }
        }
    }
    __syncthreads();
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p < P) {
        const float v = 1.f / (1.f + expf(-(ac[0] * y3[p] + ac[1])));
        m[p] = v;
        if (corr_map_out) corr_map_out[p] = v;
    }
}

// ------------------------------------------------------------------ K7: gating as planes + GAP(x_corr)
// grid (C/64, N): Xc = m X, Xu = X - Xc (planes), gc[n][c] = mean_s Xc
__global__ void __launch_bounds__(256) gate_planes_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl,
                                                          const float* __restrict__ m, __nv_bfloat16* __restrict__ ch, __nv_bfloat16* __restrict__ cl,
                                                          __nv_bfloat16* __restrict__ uh, __nv_bfloat16* __restrict__ ul, float* __restrict__ gc) {
    __shared__ float red[32 * 65];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const Tile t;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t p = (size_t)n * HS + t.r0 + 32 * k;
        const size_t off = p * HC + c0 + t.cg * 8;
        const float mp = m[p];
        float x[8], xc[8], xu[8];
        load8_planes(xh + off, xl + off, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) { xc[i] = x[i] * mp; xu[i] = x[i] * (1.f - mp); acc[i] += xc[i]; }
        store8_planes(ch + off, cl + off, xc);
        store8_planes(uh + off, ul + off, xu);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= (1.f / HS);
    tile_colsum(acc, red, gc + (size_t)n * HC + c0, t);
}

// GAP of given x_corr planes only (stand-alone TRL entry; fused path gets it from gate_planes)
__global__ void __launch_bounds__(256) gap_planes_kernel(const __nv_bfloat16* __restrict__ ch, const __nv_bfloat16* __restrict__ cl,
                                                         float* __restrict__ gc) {
    __shared__ float red[32 * 65];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const Tile t;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t off = ((size_t)n * HS + t.r0 + 32 * k) * HC + c0 + t.cg * 8;
        float x[8];
        load8_planes(ch + off, cl + off, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += x[i] * (1.f / HS);
    }
    tile_colsum(acc, red, gc + (size_t)n * HC + c0, t);
}

// ------------------------------------------------------------------ K8: M0 = mean_t Xu, Z0 = M0 + Xu[tau0(d)]   grid (C/64, B)
__global__ void __launch_bounds__(256) trl_init_kernel(const __nv_bfloat16* __restrict__ uh, const __nv_bfloat16* __restrict__ ul, int T, int R,
                                                       __nv_bfloat16* __restrict__ mem_hi, __nv_bfloat16* __restrict__ mem_lo,
                                                       __nv_bfloat16* __restrict__ z_hi, __nv_bfloat16* __restrict__ z_lo,
                                                       __half* __restrict__ mem16, __half* __restrict__ z16) {
    const int b = blockIdx.y, c0 = blockIdx.x * 64;
    const Tile t;
    float acc[4][8], first[4][8], last[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
    for (int tt = 0; tt < T; ++tt) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const size_t off = (((size_t)b * T + tt) * HS + t.r0 + 32 * k) * HC + c0 + t.cg * 8;
            load8_planes(uh + off, ul + off, last[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[k][i] += last[k][i];
                if (tt == 0) first[k][i] = last[k][i];
            }
        }
    }
    const float inv = 1.f / (float)T;
    const size_t dstride = (size_t)R * HC;   // direction stride inside one slot
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t off = ((size_t)b * HS + t.r0 + 32 * k) * HC + c0 + t.cg * 8;
        float m0[8], zf[8], zb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { m0[i] = acc[k][i] * inv; zf[i] = m0[i] + first[k][i]; zb[i] = m0[i] + last[k][i]; }
        store8_planes(mem_hi + off, mem_lo + off, m0);
        store8_planes(mem_hi + dstride + off, mem_lo + dstride + off, m0);
        if (mem16) { store8_f16(mem16 + off, m0); store8_f16(mem16 + dstride + off, m0); }   // the f1 weight-gradient operand (training)
        store8_planes(z_hi + off, z_lo + off, zf);
        store8_planes(z_hi + dstride + off, z_lo + dstride + off, zb);
        if (z16) { store8_f16(z16 + off, zf); store8_f16(z16 + dstride + off, zb); }
    }
}

// ------------------------------------------------------------------ K11-13: squeeze-excite on the pooled squared difference
// Nothing in the recurrence depends on it (it only feeds f_corr), so it runs once after the T steps for every
// (clip, direction, step): grid (B, 2, T), 512 threads.
//   q = sum of the 4 per-warp partials / S;  h = relu(L1 q);  a = sigmoid(L2 h);
//   out_d[d][b*T + tau][c] = (1 + a[c]) * gc[b*T + tau][c]      (F4)
struct SePtrs { const float* l1[2]; const float* l2[2]; };
constexpr int SE_THREADS = 512;
__global__ void __launch_bounds__(SE_THREADS) se_fwd_kernel(const float* __restrict__ qpart, SePtrs sp, const float* __restrict__ gc, int B, int T,
                                                            float* __restrict__ se_q, float* __restrict__ se_h, float* __restrict__ se_a,
                                                            float* __restrict__ out_d) {
    __shared__ __align__(16) float q[HC];
    __shared__ __align__(16) float h[HSE];
    const int b = blockIdx.x, d = blockIdx.y, i = blockIdx.z;
    const int tau = d ? T - 1 - i : i;
    const size_t slot = ((size_t)i * 2 + d) * B + b;
    const float* qp = qpart + slot * 4 * HC;
    for (int c = threadIdx.x; c < HC; c += SE_THREADS) {
        const float v = (qp[c] + qp[HC + c] + qp[2 * HC + c] + qp[3 * HC + c]) * (1.f / HS);
        q[c] = v;
        se_q[slot * HC + c] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = SE_THREADS / 32;
    for (int j = warp; j < HSE; j += NW) {
        const float* w = sp.l1[d] + (size_t)j * HC;
        float acc = 0.f;
#pragma unroll 4
        for (int k = lane * 4; k < HC; k += 128) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k));
            const float4 qv = *reinterpret_cast<const float4*>(q + k);
            acc += wv.x * qv.x + wv.y * qv.y + wv.z * qv.z + wv.w * qv.w;
        }
        acc = warp_sum(acc);
        if (lane == 0) { const float r = fmaxf(acc, 0.f); h[j] = r; se_h[slot * HSE + j] = r; }
    }
    __syncthreads();
    const float4 hv = *reinterpret_cast<const float4*>(h + lane * 4);
    const size_t n = (size_t)b * T + tau;
    for (int c = warp; c < HC; c += NW) {                 // one warp per output channel: a 512-byte row of L2, coalesced
        const float4 wv = __ldg(reinterpret_cast<const float4*>(sp.l2[d] + (size_t)c * HSE + lane * 4));
        float acc = wv.x * hv.x + wv.y * hv.y + wv.z * hv.z + wv.w * hv.w;
        acc = warp_sum(acc);
        if (lane == 0) {
            const float a = 1.f / (1.f + expf(-acc));
            se_a[slot * HC + c] = a;
            out_d[((size_t)d * B * T + n) * HC + c] = (1.f + a) * gc[n * HC + c];
        }
    }
}

// ------------------------------------------------------------------ BN + ReLU + re-split   grid (Cn/64, R/128, 2)
__global__ void __launch_bounds__(256) bnrelu_split_kernel(const float* __restrict__ hraw, const float* __restrict__ stat, int Cn, int R,
                                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                           __half* __restrict__ p16 = nullptr) {
    const int z = blockIdx.z, c0 = blockIdx.x * 64;
    const Tile t;
    const float* st = stat + (size_t)z * 4 * Cn;
    float a[8], c[8];
    load8(st + c0 + t.cg * 8, a); load8(st + Cn + c0 + t.cg * 8, c);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t off = ((size_t)z * R + blockIdx.y * 128 + t.r0 + 32 * k) * Cn + c0 + t.cg * 8;
        float v[8];
        load8(hraw + off, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(a[i] * v[i] + c[i], 0.f);
        store8_planes(hi + off, lo + off, v);
        if (p16) store8_f16(p16 + off, v);            // training: the operand of the single-pass weight-gradient GEMM
    }
}

// ------------------------------------------------------------------ K14 tail: M' = relu(bn3(H3) + Z);  Z' = M' + Xu[tau_next(d)]
// grid (C/64, B, 2)
__global__ void __launch_bounds__(256) memo_update_kernel(const float* __restrict__ h3, const float* __restrict__ stat3,
                                                          const __nv_bfloat16* __restrict__ z_hi, const __nv_bfloat16* __restrict__ z_lo,
                                                          const __nv_bfloat16* __restrict__ uh, const __nv_bfloat16* __restrict__ ul, int T, int R,
                                                          int has_next, int tau_next0, int tau_next1,
                                                          __nv_bfloat16* __restrict__ mem_hi, __nv_bfloat16* __restrict__ mem_lo,
                                                          __nv_bfloat16* __restrict__ zn_hi, __nv_bfloat16* __restrict__ zn_lo,
                                                          __half* __restrict__ mem16, __half* __restrict__ zn16 = nullptr) {
    const int z = blockIdx.z, b = blockIdx.y, c0 = blockIdx.x * 64;
    const Tile t;
    const float* st = stat3 + (size_t)z * 4 * HC;
    float a[8], c[8];
    load8(st + c0 + t.cg * 8, a); load8(st + HC + c0 + t.cg * 8, c);
    const int tau = z ? tau_next1 : tau_next0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int s = t.r0 + 32 * k;
        const size_t off = ((size_t)z * R + (size_t)b * HS + s) * HC + c0 + t.cg * 8;
        float h[8], zz[8], mn[8];
        load8(h3 + off, h);
        load8_planes(z_hi + off, z_lo + off, zz);
#pragma unroll
        for (int i = 0; i < 8; ++i) mn[i] = fmaxf(a[i] * h[i] + c[i] + zz[i], 0.f);
        store8_planes(mem_hi + off, mem_lo + off, mn);
        if (mem16) store8_f16(mem16 + off, mn);
        if (has_next) {
            const size_t xoff = (((size_t)b * T + tau) * HS + s) * HC + c0 + t.cg * 8;
            float xu[8];
            load8_planes(uh + xoff, ul + xoff, xu);
#pragma unroll
            for (int i = 0; i < 8; ++i) xu[i] += mn[i];
            store8_planes(zn_hi + off, zn_lo + off, xu);
            if (zn16) store8_f16(zn16 + off, xu);
        }
    }
}

// ------------------------------------------------------------------ K15: outputs
// f_uncorr[b][c] = mean_s M_fwd + mean_s M_bwd    grid (C/64, B)
__global__ void __launch_bounds__(256) trl_final_kernel(const __nv_bfloat16* __restrict__ mem_hi, const __nv_bfloat16* __restrict__ mem_lo, int R,
                                                        float* __restrict__ f_uncorr) {
    __shared__ float red[32 * 65];
    const int b = blockIdx.y, c0 = blockIdx.x * 64;
    const Tile t;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int z = 0; z < 2; ++z)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const size_t off = ((size_t)z * R + (size_t)b * HS + t.r0 + 32 * k) * HC + c0 + t.cg * 8;
            float v[8];
            load8_planes(mem_hi + off, mem_lo + off, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += v[i] * (1.f / HS);
        }
    tile_colsum(acc, red, f_uncorr + (size_t)b * HC + c0, t);
}

__global__ void add2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, size_t n4) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    reinterpret_cast<float4*>(out)[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
}

// ------------------------------------------------------------------ host orchestration
static int check_params(grl_handle* h, const grl_head_params* p, int which) {      // which: 1 = GCE, 2 = TRL, 3 = both
    if (which & 1) {
    const void* req[] = {p->glo_fc_w, p->glo_fc_b, p->glo_bn.weight, p->glo_bn.bias, p->glo_bn.running_mean, p->glo_bn.running_var,
                         p->atte0_w, p->atte_bn1.weight, p->atte_bn1.bias, p->atte_bn1.running_mean, p->atte_bn1.running_var,
                         p->atte2_w, p->atte_bn3.weight, p->atte_bn3.bias, p->atte_bn3.running_mean, p->atte_bn3.running_var,
                         p->atte5_w, p->atte_bn6.weight, p->atte_bn6.bias, p->atte_bn6.running_mean, p->atte_bn6.running_var};
    for (const void* q : req)
        if (!q) return set_error(h, GRL_EINVAL, "grl_head: NULL GCE parameter pointer");
    }
    for (int d = 0; d < 2 && (which & 2); ++d) {
        const void* r2[] = {p->f1_w[d], p->f1_b[d], p->f2_w[d], p->f2_b[d], p->se1_w[d], p->se2_w[d], p->memo_conv1_w[d],
                            p->memo_conv2_w[d], p->memo_conv3_w[d], p->memo_bn1[d].weight, p->memo_bn1[d].bias,
                            p->memo_bn1[d].running_mean, p->memo_bn1[d].running_var, p->memo_bn2[d].weight, p->memo_bn2[d].bias,
                            p->memo_bn2[d].running_mean, p->memo_bn2[d].running_var, p->memo_bn3[d].weight, p->memo_bn3[d].bias,
                            p->memo_bn3[d].running_mean, p->memo_bn3[d].running_var};
        for (const void* q : r2)
            if (!q) return set_error(h, GRL_EINVAL, "grl_head: NULL TRL parameter pointer (direction %d)", d);
    }
    return GRL_OK;
}

static BnPtrs bn_ptrs(const grl_bn_params& a, const grl_bn_params& b) {
    BnPtrs r;
    r.gamma[0] = a.weight; r.beta[0] = a.bias; r.rmean[0] = a.running_mean; r.rvar[0] = a.running_var;
    r.gamma[1] = b.weight; r.beta[1] = b.bias; r.rmean[1] = b.running_mean; r.rvar[1] = b.running_var;
    return r;
}

int head_prepare_weights(grl_handle* h, cudaStream_t st, const grl_head_params* p, const HeadWs& w, int which) {
    if (which & 1) {
        GRL_TRY(split_planes(h, st, p->atte0_w, HC + HG, WS_BF(w, w1a_hi), WS_BF(w, w1a_lo), HC, HG, HC));
        GRL_TRY(split_planes(h, st, p->atte2_w, HG, WS_BF(w, w2_hi), WS_BF(w, w2_lo), HG, HMID, HG));
        GRL_TRY(split_planes(h, st, p->glo_fc_w, HC, WS_BF(w, wg_hi), WS_BF(w, wg_lo), HC, HG, HC));
        GRL_TRY(split_planes(h, st, p->atte0_w + HC, HC + HG, WS_BF(w, w1b_hi), WS_BF(w, w1b_lo), HG, HG, HG));
    }
    for (int d = 0; d < 2 && (which & 2); ++d) {
        GRL_TRY(split_planes(h, st, p->f2_w[d], HC, WS_BF(w, wf2_hi) + (size_t)d * HC * HC, WS_BF(w, wf2_lo) + (size_t)d * HC * HC, HC, HC, HC));
        GRL_TRY(split_planes(h, st, p->f1_w[d], HC, WS_BF(w, wf1_hi) + (size_t)d * HC * HC, WS_BF(w, wf1_lo) + (size_t)d * HC * HC, HC, HC, HC));
        GRL_TRY(split_planes(h, st, p->memo_conv1_w[d], HC, WS_BF(w, wc1_hi) + (size_t)d * HB * HC, WS_BF(w, wc1_lo) + (size_t)d * HB * HC, HC, HB, HC));
        GRL_TRY(split_planes(h, st, p->memo_conv2_w[d], HB, WS_BF(w, wc2_hi) + (size_t)d * HB * HB, WS_BF(w, wc2_lo) + (size_t)d * HB * HB, HB, HB, HB));
        GRL_TRY(split_planes(h, st, p->memo_conv3_w[d], HB, WS_BF(w, wc3_hi) + (size_t)d * HC * HB, WS_BF(w, wc3_lo) + (size_t)d * HC * HB, HB, HC, HB));
        GRL_CUDA(h, cudaMemcpyAsync(WS_F32(w, bf2cat) + (size_t)d * HC, p->f2_b[d], HC * 4, cudaMemcpyDeviceToDevice, st));
        GRL_CUDA(h, cudaMemcpyAsync(WS_F32(w, bf1cat) + (size_t)d * HC, p->f1_b[d], HC * 4, cudaMemcpyDeviceToDevice, st));
    }
    return GRL_OK;
}

static int bn_finalize(grl_handle* h, cudaStream_t st, const float* psum, const float* psq, int nparts, long long bstride, int Cn,
                       double count, const BnPtrs& bp, float* stat, int train, int nz) {
    dim3 grid((Cn + 31) / 32, nz);
    bn_finalize_kernel<<<grid, dim3(32, nparts > 256 ? 32 : 8), 0, st>>>(psum, psq, nparts, bstride, Cn, count, bp, stat, train);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

}  // namespace grl

using namespace grl;

extern "C" size_t grl_head_workspace_bytes(int B, int T, int save_for_backward) {
    if (B <= 0 || T <= 0) return 0;
    return head_ws_layout(B, T, save_for_backward ? 1 : 0).total;
}

extern "C" int grl_head_ws_lookup(int B, int T, int save_for_backward, const char* name, size_t* offset, size_t* bytes) {
    if (B <= 0 || T <= 0 || !name || !offset || !bytes) return GRL_EINVAL;
    const HeadWs w = head_ws_layout(B, T, save_for_backward ? 1 : 0);
#define X(nm, esz, count)                                                  \
    if (strcmp(name, #nm) == 0) { *offset = w.off_##nm; *bytes = w.bytes_##nm; return GRL_OK; }
    GRL_HEAD_BUFFERS(X)
#undef X
    return GRL_EINVAL;
}

namespace grl {

// GCE: layer4 maps -> corr_map m, gated planes Xc / Xu, GAP(x_corr)      (basebranch.py:56-68)
static int gce_forward_part(grl_handle* h, cudaStream_t st, const grl_head_params* p, const HeadWs& w, const float* x, int train,
                            float* corr_map, float* x_uncorr, float* x_corr) {
    const int B = w.B, T = w.T, N = w.N, P = w.P;
    // ---------------- GCE ----------------
    nchw_to_planes_kernel<<<dim3(HC / 64, N), 256, 0, st>>>(x, WS_BF(w, xp_hi), WS_BF(w, xp_lo), WS_F32(w, gx), 1.f);
    GRL_LAUNCH_CHECK(h);
    glo_mean_kernel<<<(B * HC + 255) / 256, 256, 0, st>>>(WS_F32(w, gx), WS_F32(w, g), B, T);
    GRL_LAUNCH_CHECK(h);
    small_linear_kernel<<<(HG * 32 + 255) / 256, 256, 0, st>>>(WS_F32(w, g), HC, p->glo_fc_w, HC, p->glo_fc_b, WS_F32(w, u), HG, B, HC, HG);
    GRL_LAUNCH_CHECK(h);
    bn1d_relu_kernel<<<(HG + 127) / 128, 128, 0, st>>>(WS_F32(w, u), WS_F32(w, glo), B, HG, p->glo_bn.weight, p->glo_bn.bias,
                                                       p->glo_bn.running_mean, p->glo_bn.running_var, WS_F32(w, glo_stat), train);
    GRL_LAUNCH_CHECK(h);
    small_linear_kernel<<<(HG * 32 + 255) / 256, 256, 0, st>>>(WS_F32(w, glo), HG, p->atte0_w + HC, HC + HG, nullptr, WS_F32(w, bias1), HG, B, HG, HG);
    GRL_LAUNCH_CHECK(h);
    {   // corr_atte.0: Y1 = X W1a^T + bias1[clip]   (planes out + BN statistics)
        GemmEpi e = epi_default();
        e.Phi = WS_BF(w, y1_hi); e.Plo = WS_BF(w, y1_lo); e.ldp = HG;
        e.grp_bias = WS_F32(w, bias1); e.grp_rows = T * HS; e.ld_gb = HG;
        if (train) { e.col_sum = WS_F32(w, part_a); e.col_sq = WS_F32(w, part_b); }
        Operand a{WS_BF(w, xp_hi), WS_BF(w, xp_lo), HC, 0, 0}, b{WS_BF(w, w1a_hi), WS_BF(w, w1a_lo), HC, 0, 0};
        GRL_TRY(gemm_launch(h, st, P, HG, HC, 1, a, b, e, 0));
    }
    {
        BnPtrs bp = bn_ptrs(p->atte_bn1, p->atte_bn1);
        GRL_TRY(bn_finalize(h, st, WS_F32(w, part_a), WS_F32(w, part_b), 4 * N, 0, HG, (double)P, bp, WS_F32(w, bn1_stat), train, 1));
    }
    fold_bn_into_w2_kernel<<<(HMID * 32 + 255) / 256, 256, 0, st>>>(p->atte2_w, WS_F32(w, bn1_stat), WS_BF(w, w2s_hi), WS_BF(w, w2s_lo), WS_F32(w, bias2));
    GRL_LAUNCH_CHECK(h);
    {   // corr_atte.2 on the raw Y1 with BN-folded weights
        GemmEpi e = epi_default();
        e.C = WS_F32(w, y2); e.ldc = HMID;
        e.col_bias = WS_F32(w, bias2);
        if (train) { e.col_sum = WS_F32(w, part_a); e.col_sq = WS_F32(w, part_b); }
        Operand a{WS_BF(w, y1_hi), WS_BF(w, y1_lo), HG, 0, 0}, b{WS_BF(w, w2s_hi), WS_BF(w, w2s_lo), HG, 0, 0};
        GRL_TRY(gemm_launch(h, st, P, HMID, HG, 1, a, b, e, 0));
    }
    {
        BnPtrs bp = bn_ptrs(p->atte_bn3, p->atte_bn3);
        GRL_TRY(bn_finalize(h, st, WS_F32(w, part_a), WS_F32(w, part_b), 4 * N, 0, HMID, (double)P, bp, WS_F32(w, bn2_stat), train, 1));
    }
    const int y3_blocks = (P + 63) / 64;
    gce_y3_kernel<<<y3_blocks, 256, 0, st>>>(WS_F32(w, y2), WS_F32(w, bn2_stat), p->atte5_w, WS_F32(w, y3), WS_F32(w, part_a), WS_F32(w, part_b), P);
    GRL_LAUNCH_CHECK(h);
    gce_m_kernel<<<(P + 255) / 256, 256, 0, st>>>(WS_F32(w, y3), WS_F32(w, part_a), WS_F32(w, part_b), y3_blocks, P, p->atte_bn6,
                                                  WS_F32(w, bn3_stat), WS_F32(w, m), corr_map, train);
    GRL_LAUNCH_CHECK(h);
    gate_planes_kernel<<<dim3(HC / 64, N), 256, 0, st>>>(WS_BF(w, xp_hi), WS_BF(w, xp_lo), WS_F32(w, m), WS_BF(w, xc_hi), WS_BF(w, xc_lo),
                                                         WS_BF(w, xu_hi), WS_BF(w, xu_lo), WS_F32(w, gc));
    GRL_LAUNCH_CHECK(h);
    if (x_corr) { planes_to_nchw_kernel<<<dim3(HC / 64, N), 256, 0, st>>>(WS_BF(w, xc_hi), WS_BF(w, xc_lo), x_corr); GRL_LAUNCH_CHECK(h); }
    if (x_uncorr) { planes_to_nchw_kernel<<<dim3(HC / 64, N), 256, 0, st>>>(WS_BF(w, xu_hi), WS_BF(w, xu_lo), x_uncorr); GRL_LAUNCH_CHECK(h); }

    return GRL_OK;
}

// TRL: from the Xc / Xu planes and gc = GAP(x_corr)      (grl_model.py:131-180)
static int trl_forward_part(grl_handle* h, cudaStream_t st, const grl_head_params* p, const HeadWs& w, int train, float* f_uncorr,
                            float* f_corr) {
    const int B = w.B, T = w.T, N = w.N, P = w.P, R = w.R;
    // ---------------- TRL ----------------
    // Two streams: the caller's stream `st` carries the recurrence (memory update: conv1 -> BN -> conv2 -> BN -> conv3 -> BN ->
    // ReLU, 8 dependent steps); the f2 convolution of all frames, the per-step f1 convolution (+ pooled squared difference) and
    // the squeeze-excite MLP only feed f_corr, so they run on the handle's low-priority side stream `sd` and overlap the
    // HBM-bound BN / update kernels of the chain.  Events: EV_M(i) = memory M_i ready, EV_F1(i) = f1 of step i done.
    cudaStream_t sd = (h->overlap & 1) ? h->side : st;
    const bool two = sd != st;
    auto EV_M = [&](int i) { return 8 + i; };
    auto EV_F1 = [&](int i) { return 8 + (T + 1) + i; };
    if (two) GRL_TRY(stream_wait(h, st, sd, 0));
    {   // f2 for every frame and both directions at once (F2): [P][4096]
        GemmEpi e = epi_default();
        e.C = WS_F32(w, f2); e.ldc = 2 * HC;
        e.col_bias = WS_F32(w, bf2cat); e.relu = 1;
        Operand a{WS_BF(w, xc_hi), WS_BF(w, xc_lo), HC, 0, 0}, b{WS_BF(w, wf2_hi), WS_BF(w, wf2_lo), HC, 0, 0};
        GRL_TRY(gemm_launch(h, sd, P, 2 * HC, HC, 1, a, b, e, 0));
    }
    const size_t slotM = (size_t)2 * R * HC;     // elements per mem / z slot (both directions)
    // training keeps every memory slot also as ONE fp16 plane: the operand of the single-pass f1 weight-gradient GEMM (head_bwd.cu)
    __half* mem16 = w.save ? reinterpret_cast<__half*>(WS_BF(w, mem_16)) : nullptr;
    __half* z16 = w.save ? reinterpret_cast<__half*>(WS_BF(w, z_16)) : nullptr;
    __half* h1p16 = w.save ? reinterpret_cast<__half*>(WS_BF(w, h1p_16)) : nullptr;
    __half* h2p16 = w.save ? reinterpret_cast<__half*>(WS_BF(w, h2p_16)) : nullptr;
    trl_init_kernel<<<dim3(HC / 64, B), 256, 0, st>>>(WS_BF(w, xu_hi), WS_BF(w, xu_lo), T, R, WS_BF(w, mem_hi), WS_BF(w, mem_lo),
                                                      WS_BF(w, z_hi), WS_BF(w, z_lo), mem16, z16);
    GRL_LAUNCH_CHECK(h);
    if (two) GRL_TRY(ev_record(h, EV_M(0), st));
    const int save = w.save;
    for (int i = 0; i < T; ++i) {
        const int sl = save ? i : 0;                         // per-step slot
        const int ms = save ? i : (i & 1), ms_next = save ? i + 1 : ((i + 1) & 1);
        const int zs = save ? i : (i & 1), zs_next = save ? i + 1 : ((i + 1) & 1);
        const int tau0 = i, tau1 = T - 1 - i;
        const __nv_bfloat16 *mh = WS_BF(w, mem_hi) + ms * slotM, *ml = WS_BF(w, mem_lo) + ms * slotM;
        const __nv_bfloat16 *zh = WS_BF(w, z_hi) + zs * slotM, *zl = WS_BF(w, z_lo) + zs * slotM;
        float* qpart = WS_F32(w, qpart) + (size_t)i * 2 * 4 * B * HC;
        {   // f1 on the memory + squared difference against f2[tau], pooled over the 128 pixels of each clip
            GemmEpi e = epi_default();
            if (save) { e.C = WS_F32(w, f1) + (size_t)sl * 2 * R * HC; e.ldc = HC; e.c_bstride = (long long)R * HC; }
            e.col_bias = WS_F32(w, bf1cat); e.cb_bstride = HC; e.relu = 1;
            e.sub = WS_F32(w, f2); e.ld_sub = 2 * HC; e.sub_tile_rows = (long long)T * HS;
            e.sub_row_off[0] = (long long)tau0 * HS; e.sub_row_off[1] = (long long)tau1 * HS;
            e.sub_col_off[0] = 0; e.sub_col_off[1] = HC;
            e.col_sq = qpart; e.stat_bstride = (long long)4 * B * HC;
            Operand a{mh, ml, HC, (long long)R * HC, 0}, b{WS_BF(w, wf1_hi), WS_BF(w, wf1_lo), HC, (long long)HC * HC, 0};
            if (two) GRL_TRY(ev_wait(h, EV_M(i), sd));
            GRL_TRY(gemm_launch(h, sd, R, HC, HC, 2, a, b, e, 0));
            if (two) GRL_TRY(ev_record(h, EV_F1(i), sd));
        }
        // ---- memory update: BasicBlock(M, Xu[tau]) ----
        float* h1 = WS_F32(w, h1) + (size_t)sl * 2 * R * HB;
        float* h2 = WS_F32(w, h2) + (size_t)sl * 2 * R * HB;
        float* h3 = WS_F32(w, h3) + (size_t)sl * 2 * R * HC;
        __nv_bfloat16 *h1ph = WS_BF(w, h1p_hi) + (size_t)sl * 2 * R * HB, *h1pl = WS_BF(w, h1p_lo) + (size_t)sl * 2 * R * HB;
        __nv_bfloat16 *h2ph = WS_BF(w, h2p_hi) + (size_t)sl * 2 * R * HB, *h2pl = WS_BF(w, h2p_lo) + (size_t)sl * 2 * R * HB;
        float* s1 = WS_F32(w, sbn1) + (size_t)sl * 2 * 4 * HB;
        float* s2 = WS_F32(w, sbn2) + (size_t)sl * 2 * 4 * HB;
        float* s3 = WS_F32(w, sbn3) + (size_t)sl * 2 * 4 * HC;
        {
            GemmEpi e = epi_default();
            e.C = h1; e.ldc = HB; e.c_bstride = (long long)R * HB;
            if (train) { e.col_sum = WS_F32(w, part_a); e.col_sq = WS_F32(w, part_b); e.stat_bstride = (long long)4 * B * HB; }
            Operand a{zh, zl, HC, (long long)R * HC, 0}, b{WS_BF(w, wc1_hi), WS_BF(w, wc1_lo), HC, (long long)HB * HC, 0};
            GRL_TRY(gemm_launch(h, st, R, HB, HC, 2, a, b, e, 0));
            GRL_TRY(bn_finalize(h, st, WS_F32(w, part_a), WS_F32(w, part_b), 4 * B, (long long)4 * B * HB, HB, (double)R,
                                bn_ptrs(p->memo_bn1[0], p->memo_bn1[1]), s1, train, 2));
            bnrelu_split_kernel<<<dim3(HB / 64, R / 128, 2), 256, 0, st>>>(h1, s1, HB, R, h1ph, h1pl,
                                                                           h1p16 ? h1p16 + (size_t)sl * 2 * R * HB : nullptr);
            GRL_LAUNCH_CHECK(h);
        }
        {
            GemmEpi e = epi_default();
            e.C = h2; e.ldc = HB; e.c_bstride = (long long)R * HB;
            if (train) { e.col_sum = WS_F32(w, part_a); e.col_sq = WS_F32(w, part_b); e.stat_bstride = (long long)4 * B * HB; }
            Operand a{h1ph, h1pl, HB, (long long)R * HB, 0}, b{WS_BF(w, wc2_hi), WS_BF(w, wc2_lo), HB, (long long)HB * HB, 0};
            GRL_TRY(gemm_launch(h, st, R, HB, HB, 2, a, b, e, 0));
            GRL_TRY(bn_finalize(h, st, WS_F32(w, part_a), WS_F32(w, part_b), 4 * B, (long long)4 * B * HB, HB, (double)R,
                                bn_ptrs(p->memo_bn2[0], p->memo_bn2[1]), s2, train, 2));
            bnrelu_split_kernel<<<dim3(HB / 64, R / 128, 2), 256, 0, st>>>(h2, s2, HB, R, h2ph, h2pl,
                                                                           h2p16 ? h2p16 + (size_t)sl * 2 * R * HB : nullptr);
            GRL_LAUNCH_CHECK(h);
        }
        {
            GemmEpi e = epi_default();
            e.C = h3; e.ldc = HC; e.c_bstride = (long long)R * HC;
            if (train) { e.col_sum = WS_F32(w, part_a); e.col_sq = WS_F32(w, part_b); e.stat_bstride = (long long)4 * B * HC; }
            Operand a{h2ph, h2pl, HB, (long long)R * HB, 0}, b{WS_BF(w, wc3_hi), WS_BF(w, wc3_lo), HB, (long long)HC * HB, 0};
            GRL_TRY(gemm_launch(h, st, R, HC, HB, 2, a, b, e, 0));
            GRL_TRY(bn_finalize(h, st, WS_F32(w, part_a), WS_F32(w, part_b), 4 * B, (long long)4 * B * HC, HC, (double)R,
                                bn_ptrs(p->memo_bn3[0], p->memo_bn3[1]), s3, train, 2));
        }
        const int has_next = (i + 1 < T) ? 1 : 0;
        // inference ping-pongs two memory slots: this update overwrites the slot M_{i-1} lives in, which f1(i-1) may still read
        if (two && !save && i >= 1) GRL_TRY(ev_wait(h, EV_F1(i - 1), st));
        memo_update_kernel<<<dim3(HC / 64, B, 2), 256, 0, st>>>(h3, s3, zh, zl, WS_BF(w, xu_hi), WS_BF(w, xu_lo), T, R, has_next, i + 1,
                                                                T - 2 - i, WS_BF(w, mem_hi) + ms_next * slotM, WS_BF(w, mem_lo) + ms_next * slotM,
                                                                WS_BF(w, z_hi) + (has_next ? zs_next : zs) * slotM,
                                                                WS_BF(w, z_lo) + (has_next ? zs_next : zs) * slotM,
                                                                mem16 ? mem16 + ms_next * slotM : nullptr,
                                                                (z16 && has_next) ? z16 + zs_next * slotM : nullptr);
        GRL_LAUNCH_CHECK(h);
        if (two) GRL_TRY(ev_record(h, EV_M(i + 1), st));
    }
    {   // squeeze-excite + F4 pooled shortcut for all steps at once (after the last f1, in order on the side stream)
        SePtrs sp; sp.l1[0] = p->se1_w[0]; sp.l1[1] = p->se1_w[1]; sp.l2[0] = p->se2_w[0]; sp.l2[1] = p->se2_w[1];
        se_fwd_kernel<<<dim3(B, 2, T), SE_THREADS, 0, sd>>>(WS_F32(w, qpart), sp, WS_F32(w, gc), B, T, WS_F32(w, se_q), WS_F32(w, se_h),
                                                            WS_F32(w, se_a), WS_F32(w, out_d));
        GRL_LAUNCH_CHECK(h);
    }
    if (two) GRL_TRY(stream_wait(h, sd, st, 1));
    const int mfin = save ? T : (T & 1);
    trl_final_kernel<<<dim3(HC / 64, B), 256, 0, st>>>(WS_BF(w, mem_hi) + mfin * slotM, WS_BF(w, mem_lo) + mfin * slotM, R, f_uncorr);
    GRL_LAUNCH_CHECK(h);
    const size_t n4 = (size_t)N * HC / 4;
    add2_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(WS_F32(w, out_d), WS_F32(w, out_d) + (size_t)N * HC, f_corr, n4);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

static int head_setup(grl_handle* h, const char* who, const grl_head_params* p, int B, int T, int train, void* workspace,
                      size_t workspace_bytes, int save, int which, HeadWs* w) {
    if (!h || !p || !workspace) return set_error(h, GRL_EINVAL, "%s: NULL argument", who);
    if (B <= 0 || T <= 0) return set_error(h, GRL_EINVAL, "%s: need B, T > 0", who);
    if (train && (which & 1) && B < 2) return set_error(h, GRL_EINVAL, "%s: train-mode BatchNorm1d needs B >= 2 (got %d)", who, B);
    GRL_TRY(check_params(h, p, which));
    *w = head_ws_layout(B, T, save ? 1 : 0);
    if (workspace_bytes < w->total) return set_error(h, GRL_ENOMEM, "%s: workspace %zu < %zu bytes", who, workspace_bytes, w->total);
    if (reinterpret_cast<uintptr_t>(workspace) & 1023) return set_error(h, GRL_EINVAL, "%s: workspace must be 1024-byte aligned", who);
    w->base = (uint8_t*)workspace;
    return GRL_OK;
}

}  // namespace grl

extern "C" int grl_head_forward(grl_handle* h, const grl_head_params* p, const float* x, int B, int T, int train, float* f_uncorr,
                                float* f_corr, float* corr_map, float* x_uncorr, float* x_corr, void* workspace,
                                size_t workspace_bytes, int save_for_backward, void* stream) {
    if (!x || !f_uncorr || !f_corr) return set_error(h, GRL_EINVAL, "grl_head_forward: NULL argument");
    HeadWs w;
    GRL_TRY(head_setup(h, "grl_head_forward", p, B, T, train, workspace, workspace_bytes, save_for_backward, 3, &w));
    cudaStream_t st = (cudaStream_t)stream;
    GRL_TRY(head_prepare_weights(h, st, p, w, 3));
    GRL_TRY(gce_forward_part(h, st, p, w, x, train, corr_map, x_uncorr, x_corr));
    return trl_forward_part(h, st, p, w, train, f_uncorr, f_corr);
}

extern "C" int grl_gce_forward(grl_handle* h, const grl_head_params* p, const float* x, int B, int T, int train, float* x_uncorr,
                               float* x_corr, float* corr_map, void* workspace, size_t workspace_bytes, int save_for_backward,
                               void* stream) {
    if (!x || !x_uncorr || !x_corr || !corr_map) return set_error(h, GRL_EINVAL, "grl_gce_forward: NULL argument");
    HeadWs w;
    GRL_TRY(head_setup(h, "grl_gce_forward", p, B, T, train, workspace, workspace_bytes, save_for_backward, 1, &w));
    cudaStream_t st = (cudaStream_t)stream;
    GRL_TRY(head_prepare_weights(h, st, p, w, 1));
    return gce_forward_part(h, st, p, w, x, train, corr_map, x_uncorr, x_corr);
}

extern "C" int grl_trl_forward(grl_handle* h, const grl_head_params* p, const float* x_uncorr, const float* x_corr, int B, int T,
                               int train, float* f_uncorr, float* f_corr, void* workspace, size_t workspace_bytes,
                               int save_for_backward, void* stream) {
    if (!x_uncorr || !x_corr || !f_uncorr || !f_corr) return set_error(h, GRL_EINVAL, "grl_trl_forward: NULL argument");
    HeadWs w;
    GRL_TRY(head_setup(h, "grl_trl_forward", p, B, T, train, workspace, workspace_bytes, save_for_backward, 2, &w));
    cudaStream_t st = (cudaStream_t)stream;
    GRL_TRY(head_prepare_weights(h, st, p, w, 2));
    // the caller's maps -> pixel-major planes; gc = GAP(x_corr) per frame rides along
    nchw_to_planes_kernel<<<dim3(HC / 64, w.N), 256, 0, st>>>(x_corr, WS_BF(w, xc_hi), WS_BF(w, xc_lo), WS_F32(w, gc), 1.f / HS);
    GRL_LAUNCH_CHECK(h);
    nchw_to_planes_kernel<<<dim3(HC / 64, w.N), 256, 0, st>>>(x_uncorr, WS_BF(w, xu_hi), WS_BF(w, xu_lo), nullptr, 1.f);
    GRL_LAUNCH_CHECK(h);
    return trl_forward_part(h, st, p, w, train, f_uncorr, f_corr);
}

// ------------------------------------------------------------------ BasicBlock.forward(x1, x2) as a stand-alone operator
// reid/models/grl_model.py:67-85: relu(bn3(conv3(relu(bn2(conv2(relu(bn1(conv1(x1 + x2)))))))) + (x1 + x2)), x* [n][2048][16][8].
// The same kernels as one memory-update step of the TRL recurrence (one "direction", n frames = n 128-row tiles).
struct BlockWs { size_t z_hi, z_lo, w1_hi, w1_lo, w2_hi, w2_lo, w3_hi, w3_lo, h1, h1p_hi, h1p_lo, h2, h2p_hi, h2p_lo, h3, s1, s2, s3, pa, pb, o_hi, o_lo, total; };
static BlockWs block_ws_layout(int n) {
    BlockWs L;
    const size_t R = (size_t)n * HS;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    L.z_hi = take(R * HC * 2); L.z_lo = take(R * HC * 2);
    L.w1_hi = take((size_t)HB * HC * 2); L.w1_lo = take((size_t)HB * HC * 2);
    L.w2_hi = take((size_t)HB * HB * 2); L.w2_lo = take((size_t)HB * HB * 2);
    L.w3_hi = take((size_t)HC * HB * 2); L.w3_lo = take((size_t)HC * HB * 2);
    L.h1 = take(R * HB * 4); L.h1p_hi = take(R * HB * 2); L.h1p_lo = take(R * HB * 2);
    L.h2 = take(R * HB * 4); L.h2p_hi = take(R * HB * 2); L.h2p_lo = take(R * HB * 2);
    L.h3 = take(R * HC * 4);
    L.s1 = take((size_t)4 * HB * 4); L.s2 = take((size_t)4 * HB * 4); L.s3 = take((size_t)4 * HC * 4);
    L.pa = take((size_t)4 * n * HC * 4); L.pb = take((size_t)4 * n * HC * 4);
    L.o_hi = take(R * HC * 2); L.o_lo = take(R * HC * 2);
    L.total = off;
    return L;
}
extern "C" size_t grl_basic_block_workspace_bytes(int n) { return n > 0 ? block_ws_layout(n).total : 0; }

extern "C" int grl_basic_block_forward(grl_handle* h, const float* conv1_w, const grl_bn_params* bn1, const float* conv2_w, const grl_bn_params* bn2,
                                       const float* conv3_w, const grl_bn_params* bn3, const float* x1, const float* x2, int n, int train,
                                       float* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !conv1_w || !conv2_w || !conv3_w || !bn1 || !bn2 || !bn3 || !x1 || !x2 || !out || !workspace)
        return set_error(h, GRL_EINVAL, "grl_basic_block_forward: NULL argument");
    if (n <= 0) return set_error(h, GRL_EINVAL, "grl_basic_block_forward: need n > 0");
    const BlockWs L = block_ws_layout(n);
    if (workspace_bytes < L.total) return set_error(h, GRL_ENOMEM, "grl_basic_block_forward: workspace %zu < %zu bytes", workspace_bytes, L.total);
    if (reinterpret_cast<uintptr_t>(workspace) & 1023) return set_error(h, GRL_EINVAL, "grl_basic_block_forward: workspace must be 1024-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* w = (uint8_t*)workspace;
    auto BF = [&](size_t o) { return reinterpret_cast<__nv_bfloat16*>(w + o); };
    auto F32 = [&](size_t o) { return reinterpret_cast<float*>(w + o); };
    const int R = n * HS;
    nchw_to_planes_kernel<<<dim3(HC / 64, n), 256, 0, st>>>(x1, BF(L.z_hi), BF(L.z_lo), nullptr, 1.f, x2);
    GRL_LAUNCH_CHECK(h);
    GRL_TRY(split_planes(h, st, conv1_w, HC, BF(L.w1_hi), BF(L.w1_lo), HC, HB, HC));
    GRL_TRY(split_planes(h, st, conv2_w, HB, BF(L.w2_hi), BF(L.w2_lo), HB, HB, HB));
    GRL_TRY(split_planes(h, st, conv3_w, HB, BF(L.w3_hi), BF(L.w3_lo), HB, HC, HB));
    struct Stage { const __nv_bfloat16 *a_hi, *a_lo; int K; const __nv_bfloat16 *w_hi, *w_lo; int N; float* hraw; const grl_bn_params* bn; float* stat;
                   __nv_bfloat16 *p_hi, *p_lo; };
    const Stage stages[3] = {
        {BF(L.z_hi), BF(L.z_lo), HC, BF(L.w1_hi), BF(L.w1_lo), HB, F32(L.h1), bn1, F32(L.s1), BF(L.h1p_hi), BF(L.h1p_lo)},
        {BF(L.h1p_hi), BF(L.h1p_lo), HB, BF(L.w2_hi), BF(L.w2_lo), HB, F32(L.h2), bn2, F32(L.s2), BF(L.h2p_hi), BF(L.h2p_lo)},
        {BF(L.h2p_hi), BF(L.h2p_lo), HB, BF(L.w3_hi), BF(L.w3_lo), HC, F32(L.h3), bn3, F32(L.s3), nullptr, nullptr}};
    for (int i = 0; i < 3; ++i) {
        const Stage& S = stages[i];
        GemmEpi e = epi_default();
        e.C = S.hraw; e.ldc = S.N;
        if (train) { e.col_sum = F32(L.pa); e.col_sq = F32(L.pb); }
        Operand a{S.a_hi, S.a_lo, S.K, 0, 0}, b{S.w_hi, S.w_lo, S.K, 0, 0};
        GRL_TRY(gemm_launch(h, st, R, S.N, S.K, 1, a, b, e, 0));
        GRL_TRY(bn_finalize(h, st, F32(L.pa), F32(L.pb), 4 * n, 0, S.N, (double)R, bn_ptrs(*S.bn, *S.bn), S.stat, train, 1));
        if (S.p_hi) {
            bnrelu_split_kernel<<<dim3(S.N / 64, R / 128, 1), 256, 0, st>>>(S.hraw, S.stat, S.N, R, S.p_hi, S.p_lo);
            GRL_LAUNCH_CHECK(h);
        }
    }
    memo_update_kernel<<<dim3(HC / 64, n, 1), 256, 0, st>>>(F32(L.h3), F32(L.s3), BF(L.z_hi), BF(L.z_lo), nullptr, nullptr, 1, R, 0, 0, 0, BF(L.o_hi),
                                                            BF(L.o_lo), nullptr, nullptr, nullptr);
    GRL_LAUNCH_CHECK(h);
    planes_to_nchw_kernel<<<dim3(HC / 64, n), 256, 0, st>>>(BF(L.o_hi), BF(L.o_lo), out);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}
