// grl_b200 — the verification head in training (third piece), sm_100a.
//
//   Siamese.forward(x_corr)   reid/models/Siamese.py:108-142   temporal self-attention pooling of the probe and the gallery half
//                                                              (:79-106, train-mode BatchNorm over each half separately),
//                                                              squared pair differences -> classifierBN -> classifierlinear
//   PairLoss.forward          reid/loss/pairloss.py:19-48      BCE of the pair scores against identity equality, precision
// and their backward.  Sizes: 2n = 32 clips of T = 8 frames, 2048 -> 512 projections, n^2 = 256 pairs: microseconds of work,
// so these are plain fp32 kernels with fixed reduction orders (one thread or one block per output).
//
// Forward of one half (n clips):  Qp = A Wq^T + bq;  Qb = BN(Qp);  Qh = Qb / |Qb|  (same for K);  W = softmax(Qh Kh^T) per clip;
// c[s] = sum_t W[t][s];  P = sum_s c[s] A[s];  out = P / |P|.   Pair (p, g): diff = (out_p - out_g)^2;  y = BN_c(diff);
// z = y Wc^T + bc.
#include <math_constants.h>

#include "api.h"

namespace grl {

constexpr int SD = 2048;      // feature channels
constexpr int SA = 512;       // attention width
constexpr float S_EPS = 1e-5f, S_MOM = 0.1f;

// ------------------------------------------------------------------ small fp32 GEMM: C[m][n] (+)= sum_k A(m,k) * B(n,k)
// AT = 0: A(m,k) = a[m*K + k]; AT = 1: A(m,k) = a[k*M + m].   BT likewise with N.
template <int AT, int BT>
__global__ void __launch_bounds__(256) sia_gemm_kernel(const float* __restrict__ a, const float* __restrict__ b, int M, int N, int K,
                                                       int accumulate, float* __restrict__ c) {
    __shared__ float sa[16][65], sb[16][65];
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int t = threadIdx.x; t < 64 * 16; t += 256) {
            if (AT == 0) { const int r = t >> 4, kk = t & 15; sa[kk][r] = (m0 + r < M && k0 + kk < K) ? a[(long long)(m0 + r) * K + k0 + kk] : 0.f; }
            else { const int kk = t >> 6, r = t & 63; sa[kk][r] = (m0 + r < M && k0 + kk < K) ? a[(long long)(k0 + kk) * M + m0 + r] : 0.f; }
            if (BT == 0) { const int r = t >> 4, kk = t & 15; sb[kk][r] = (n0 + r < N && k0 + kk < K) ? b[(long long)(n0 + r) * K + k0 + kk] : 0.f; }
            else { const int kk = t >> 6, r = t & 63; sb[kk][r] = (n0 + r < N && k0 + kk < K) ? b[(long long)(k0 + kk) * N + n0 + r] : 0.f; }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { av[i] = sa[kk][ty * 4 + i]; bv[i] = sb[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) {
                float* dst = c + (long long)m * N + n;
                *dst = accumulate ? *dst + acc[i][j] : acc[i][j];
            }
        }
}

__device__ __forceinline__ float block_sum(float v, float* sh) {       // blockDim.x == 256
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sh[i];
    return s;
}

// ------------------------------------------------------------------ forward
// BatchNorm1d(512) statistics of Q (which = 0) / K (1) projections, per half; running buffers updated probe first, gallery
// second (two module calls, Siamese.py:124-125).  grid (SA / 32, 2), block (32, 8).   stat [which][half][mean | rstd][SA]
__global__ void sia_bn_stats_kernel(const float* __restrict__ qp, const float* __restrict__ kp, const float* __restrict__ bq,
                                    const float* __restrict__ bk, int rows_half, int train, grl_bn_params bnq, grl_bn_params bnk,
                                    float* __restrict__ stat) {
    __shared__ float sh[2][8][33];
    const int which = blockIdx.y;
    const int c = blockIdx.x * 32 + threadIdx.x;
    const float* src = which ? kp : qp;
    const float bias = (which ? bk : bq)[c];
    const grl_bn_params& bn = which ? bnk : bnq;
    for (int half = 0; half < 2; ++half) {
        float mean, var;
        if (train) {
            float s = 0.f, sq = 0.f;                      // two passes: mean, then centred second moment
            for (int r = threadIdx.y; r < rows_half; r += 8) s += src[(size_t)(half * rows_half + r) * SA + c] + bias;
            __syncthreads();
            sh[0][threadIdx.y][threadIdx.x] = s;
            __syncthreads();
            s = 0.f;
            for (int i = 0; i < 8; ++i) s += sh[0][i][threadIdx.x];
            mean = s / rows_half;
            for (int r = threadIdx.y; r < rows_half; r += 8) {
                const float v = src[(size_t)(half * rows_half + r) * SA + c] + bias - mean;
                sq += v * v;
            }
            sh[1][threadIdx.y][threadIdx.x] = sq;
            __syncthreads();
            sq = 0.f;
            for (int i = 0; i < 8; ++i) sq += sh[1][i][threadIdx.x];
            var = sq / rows_half;
            if (threadIdx.y == 0) {
                const float unb = rows_half > 1 ? var * rows_half / (rows_half - 1) : var;
                bn.running_mean[c] = (1.f - S_MOM) * bn.running_mean[c] + S_MOM * mean;
                bn.running_var[c] = (1.f - S_MOM) * bn.running_var[c] + S_MOM * unb;
            }
        } else {
            mean = bn.running_mean[c]; var = bn.running_var[c];
        }
        if (threadIdx.y == 0) {
            float* st = stat + ((size_t)which * 2 + half) * 2 * SA;
            st[c] = mean;
            st[SA + c] = rsqrtf(var + S_EPS);
        }
    }
}
// Qh = BN(Qp + b) / |.|  per row.  grid (R, 2), block 256 (2 channels per thread)
__global__ void __launch_bounds__(256) sia_bn_norm_kernel(const float* __restrict__ qp, const float* __restrict__ kp, const float* __restrict__ bq,
                                                          const float* __restrict__ bk, int rows_half, grl_bn_params bnq, grl_bn_params bnk,
                                                          const float* __restrict__ stat, float* __restrict__ qh, float* __restrict__ kh,
                                                          float* __restrict__ qnorm, float* __restrict__ knorm) {
    __shared__ float sh[8];
    const int row = blockIdx.x, which = blockIdx.y, half = row / rows_half;
    const float* src = (which ? kp : qp) + (size_t)row * SA;
    const float* bias = which ? bk : bq;
    const grl_bn_params& bn = which ? bnk : bnq;
    const float* st = stat + ((size_t)which * 2 + half) * 2 * SA;
    float v[2], ss = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = threadIdx.x + 256 * i;
        v[i] = bn.weight[c] * (src[c] + bias[c] - st[c]) * st[SA + c] + bn.bias[c];
        ss += v[i] * v[i];
    }
    const float nrm = sqrtf(block_sum(ss, sh));
    float* dst = (which ? kh : qh) + (size_t)row * SA;
#pragma unroll
    for (int i = 0; i < 2; ++i) dst[threadIdx.x + 256 * i] = v[i] / nrm;
    if (threadIdx.x == 0) (which ? knorm : qnorm)[row] = nrm;
}
// One block per clip: W = softmax(Qh Kh^T) [T][T], c = column sums, P = sum_s c[s] x[s], out = P / |P|.   T <= 32
__global__ void __launch_bounds__(256) sia_attn_pool_kernel(const float* __restrict__ x, const float* __restrict__ qh, const float* __restrict__ kh,
                                                            int T, float* __restrict__ W, float* __restrict__ cs, float* __restrict__ P,
                                                            float* __restrict__ pnorm, float* __restrict__ out) {
    __shared__ float S[32][33];
    __shared__ float c_s[32];
    __shared__ float sh[8];
    const int i = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    for (int pair = warp; pair < T * T; pair += 8) {
        const int t = pair / T, s = pair - t * T;
        const float* q = qh + ((size_t)i * T + t) * SA;
        const float* k = kh + ((size_t)i * T + s) * SA;
        float d = 0.f;
        for (int a = lane; a < SA; a += 32) d = fmaf(q[a], k[a], d);
        d = warp_sum(d);
        if (lane == 0) S[t][s] = d;
    }
    __syncthreads();
    if (threadIdx.x < T) {
        const int t = threadIdx.x;
        float mx = -CUDART_INF_F, sum = 0.f;
        for (int s = 0; s < T; ++s) mx = fmaxf(mx, S[t][s]);
        for (int s = 0; s < T; ++s) { const float e = expf(S[t][s] - mx); S[t][s] = e; sum += e; }
        for (int s = 0; s < T; ++s) { S[t][s] /= sum; W[((size_t)i * T + t) * T + s] = S[t][s]; }
    }
    __syncthreads();
    if (threadIdx.x < T) {
        float c = 0.f;
        for (int t = 0; t < T; ++t) c += S[t][threadIdx.x];
        c_s[threadIdx.x] = c;
        cs[(size_t)i * T + threadIdx.x] = c;
    }
    __syncthreads();
    float p[8], ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int d = threadIdx.x + 256 * j;
        float acc = 0.f;
        for (int s = 0; s < T; ++s) acc = fmaf(c_s[s], x[((size_t)i * T + s) * SD + d], acc);
        p[j] = acc; ss += acc * acc;
        P[(size_t)i * SD + d] = acc;
    }
    const float nrm = sqrtf(block_sum(ss, sh));
#pragma unroll
    for (int j = 0; j < 8; ++j) out[(size_t)i * SD + threadIdx.x + 256 * j] = p[j] / nrm;
    if (threadIdx.x == 0) pnorm[i] = nrm;
}
// classifierBN statistics over the n^2 pair rows; one thread per channel, pairs in (p, g) order.  cstat [mean | rstd][SD]
__global__ void sia_pair_stats_kernel(const float* __restrict__ out, int n, int train, grl_bn_params bn, float* __restrict__ cstat) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= SD) return;
    float mean, var;
    if (train) {
        float s = 0.f, sq = 0.f;
        const float cnt = (float)n * n;
        for (int p = 0; p < n; ++p) {
            const float a = out[(size_t)p * SD + d];
            for (int g = 0; g < n; ++g) { const float t = a - out[(size_t)(n + g) * SD + d]; s += t * t; }
        }
        mean = s / cnt;
        for (int p = 0; p < n; ++p) {
            const float a = out[(size_t)p * SD + d];
            for (int g = 0; g < n; ++g) { const float t = a - out[(size_t)(n + g) * SD + d]; const float v = t * t - mean; sq += v * v; }
        }
        var = sq / cnt;
        const float unb = cnt > 1.f ? var * cnt / (cnt - 1.f) : var;
        bn.running_mean[d] = (1.f - S_MOM) * bn.running_mean[d] + S_MOM * mean;
        bn.running_var[d] = (1.f - S_MOM) * bn.running_var[d] + S_MOM * unb;
    } else {
        mean = bn.running_mean[d]; var = bn.running_var[d];
    }
    cstat[d] = mean;
    cstat[SD + d] = rsqrtf(var + S_EPS);
}
// one block per pair: z[j] = sum_d BN_c(diff)[d] Wc[j][d] + bc[j]
__global__ void __launch_bounds__(256) sia_pair_cls_kernel(const float* __restrict__ out, int n, grl_bn_params bn, const float* __restrict__ cstat,
                                                           const float* __restrict__ wc, const float* __restrict__ bc, float* __restrict__ z) {
    __shared__ float sh[8];
    const int p = blockIdx.x / n, g = blockIdx.x - p * n;
    float z0 = 0.f, z1 = 0.f;
    for (int d = threadIdx.x; d < SD; d += 256) {
        const float t = out[(size_t)p * SD + d] - out[(size_t)(n + g) * SD + d];
        const float y = bn.weight[d] * (t * t - cstat[d]) * cstat[SD + d] + bn.bias[d];
        z0 = fmaf(y, wc[d], z0);
        z1 = fmaf(y, wc[SD + d], z1);
    }
    z0 = block_sum(z0, sh);
    z1 = block_sum(z1, sh);
    if (threadIdx.x == 0) { z[(size_t)blockIdx.x * 2] = z0 + bc[0]; z[(size_t)blockIdx.x * 2 + 1] = z1 + bc[1]; }
}

// ------------------------------------------------------------------ backward
// per channel d over the pair rows: s1 = sum dy, s2 = sum dy * xhat (dy = dz0 Wc[0][d] + dz1 Wc[1][d]); d gamma/beta, d Wc
__global__ void sia_pair_bwd_reduce_kernel(const float* __restrict__ out, const float* __restrict__ dz, int n, grl_bn_params bn,
                                           const float* __restrict__ cstat, const float* __restrict__ wc, float* __restrict__ s12,
                                           float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dwc) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= SD) return;
    const float w0 = wc[d], w1 = wc[SD + d], mu = cstat[d], rs = cstat[SD + d], ga = bn.weight[d], be = bn.bias[d];
    float s1 = 0.f, s2 = 0.f, a0 = 0.f, a1 = 0.f;
    for (int p = 0; p < n; ++p) {
        const float a = out[(size_t)p * SD + d];
        for (int g = 0; g < n; ++g) {
            const float t = a - out[(size_t)(n + g) * SD + d];
            const float xh = (t * t - mu) * rs;
            const float d0 = dz[((size_t)p * n + g) * 2], d1 = dz[((size_t)p * n + g) * 2 + 1];
            const float dy = d0 * w0 + d1 * w1;
            s1 += dy; s2 += dy * xh;
            const float y = ga * xh + be;
            a0 = fmaf(d0, y, a0); a1 = fmaf(d1, y, a1);
        }
    }
    s12[d] = s1; s12[SD + d] = s2;
    dgamma[d] = s2; dbeta[d] = s1;
    dwc[d] = a0; dwc[SD + d] = a1;
}
// d out[sample][d] = upstream + pair part.  grid (2n), block 256
__global__ void __launch_bounds__(256) sia_pair_bwd_apply_kernel(const float* __restrict__ out, const float* __restrict__ dz, int n, grl_bn_params bn,
                                                                 const float* __restrict__ cstat, const float* __restrict__ wc,
                                                                 const float* __restrict__ s12, const float* __restrict__ d_out_up,
                                                                 float* __restrict__ d_out) {
    const int smp = blockIdx.x;
    const bool probe = smp < n;
    const float cnt = (float)n * n;
    for (int d = threadIdx.x; d < SD; d += 256) {
        const float w0 = wc[d], w1 = wc[SD + d], mu = cstat[d], rs = cstat[SD + d], k0 = bn.weight[d] * rs;
        const float m1 = s12[d] / cnt, m2 = s12[SD + d] / cnt;
        const float me = out[(size_t)smp * SD + d];
        float acc = d_out_up ? d_out_up[(size_t)smp * SD + d] : 0.f;
        for (int o = 0; o < n; ++o) {
            const int p = probe ? smp : o, g = probe ? o : smp - n;
            const float t = probe ? me - out[(size_t)(n + g) * SD + d] : out[(size_t)p * SD + d] - me;
            const float dy = dz[((size_t)p * n + g) * 2] * w0 + dz[((size_t)p * n + g) * 2 + 1] * w1;
            const float dd = k0 * (dy - m1 - (t * t - mu) * rs * m2);        // d diff
            acc += (probe ? 2.f : -2.f) * t * dd;
        }
        d_out[(size_t)smp * SD + d] = acc;
    }
}
__global__ void sia_dbc_kernel(const float* __restrict__ dz, int npairs, float* __restrict__ dbc) {
    if (threadIdx.x < 2) {
        float s = 0.f;
        for (int i = 0; i < npairs; ++i) s += dz[(size_t)i * 2 + threadIdx.x];
        dbc[threadIdx.x] = s;
    }
}
// One block per clip: through the output normalisation, the pooling, the softmax and the Q/K normalisations.
// writes dx[i] = c[s] dP (the V path; the projection path is added afterwards), dqb / dkb [R][SA]
__global__ void __launch_bounds__(256) sia_attn_pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ qh, const float* __restrict__ kh,
                                                                const float* __restrict__ qnorm, const float* __restrict__ knorm,
                                                                const float* __restrict__ W, const float* __restrict__ cs,
                                                                const float* __restrict__ out, const float* __restrict__ pnorm,
                                                                const float* __restrict__ d_out, int T, float* __restrict__ dx,
                                                                float* __restrict__ dqb, float* __restrict__ dkb) {
    __shared__ float dP[SD];
    __shared__ float dS[32][33];
    __shared__ float dc[32];
    __shared__ float sh[8];
    const int i = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    float dot = 0.f;
    for (int d = threadIdx.x; d < SD; d += 256) dot = fmaf(out[(size_t)i * SD + d], d_out[(size_t)i * SD + d], dot);
    dot = block_sum(dot, sh);
    const float inv = 1.f / pnorm[i];
    for (int d = threadIdx.x; d < SD; d += 256) dP[d] = (d_out[(size_t)i * SD + d] - out[(size_t)i * SD + d] * dot) * inv;
    __syncthreads();
    for (int s = warp; s < T; s += 8) {                       // dc[s] = x[s] . dP ; dx[s] = c[s] dP
        const float* xs = x + ((size_t)i * T + s) * SD;
        float* dxs = dx + ((size_t)i * T + s) * SD;
        const float c = cs[(size_t)i * T + s];
        float a = 0.f;
        for (int d = lane; d < SD; d += 32) { a = fmaf(xs[d], dP[d], a); dxs[d] = c * dP[d]; }
        a = warp_sum(a);
        if (lane == 0) dc[s] = a;
    }
    __syncthreads();
    if (threadIdx.x < T) {                                    // softmax backward, row t
        const int t = threadIdx.x;
        const float* w = W + ((size_t)i * T + t) * T;
        float m = 0.f;
        for (int s = 0; s < T; ++s) m = fmaf(w[s], dc[s], m);
        for (int s = 0; s < T; ++s) dS[t][s] = w[s] * (dc[s] - m);
    }
    __syncthreads();
    // dQh[t] = sum_s dS[t][s] Kh[s]; dKh[s] = sum_t dS[t][s] Qh[t]; then through x / |x|
    for (int r = warp; r < 2 * T; r += 8) {
        const bool isq = r < T;
        const int t = isq ? r : r - T;
        const float* self = (isq ? qh : kh) + ((size_t)i * T + t) * SA;
        const float* other = (isq ? kh : qh) + (size_t)i * T * SA;
        float g[SA / 32], dotg = 0.f;
#pragma unroll
        for (int j = 0; j < SA / 32; ++j) {
            const int a = lane + 32 * j;
            float acc = 0.f;
            for (int s = 0; s < T; ++s) acc = fmaf(isq ? dS[t][s] : dS[s][t], other[(size_t)s * SA + a], acc);
            g[j] = acc;
            dotg = fmaf(acc, self[a], dotg);
        }
        dotg = warp_sum(dotg);
        const float invn = 1.f / (isq ? qnorm : knorm)[(size_t)i * T + t];
        float* dst = (isq ? dqb : dkb) + ((size_t)i * T + t) * SA;
#pragma unroll
        for (int j = 0; j < SA / 32; ++j) { const int a = lane + 32 * j; dst[a] = (g[j] - self[a] * dotg) * invn; }
    }
}
// BatchNorm1d backward of the Q (which 0) / K (1) projections, both halves; in place on dqb/dkb (-> d Qp / d Kp); d gamma, d beta
// and d bias accumulate over the two halves.  grid (SA / 32, 2), block (32, 8)
__global__ void sia_bn_bwd_kernel(const float* __restrict__ qp, const float* __restrict__ kp, const float* __restrict__ bq,
                                  const float* __restrict__ bk, int rows_half, grl_bn_params bnq, grl_bn_params bnk,
                                  const float* __restrict__ stat, float* __restrict__ dqb, float* __restrict__ dkb,
                                  float* __restrict__ dgq, float* __restrict__ dbq_bn, float* __restrict__ dgk, float* __restrict__ dbk_bn,
                                  float* __restrict__ dbias_q, float* __restrict__ dbias_k) {
    __shared__ float sh[2][8][33];
    const int which = blockIdx.y;
    const int c = blockIdx.x * 32 + threadIdx.x;
    const float* src = which ? kp : qp;
    float* dy = which ? dkb : dqb;
    const float bias = (which ? bk : bq)[c];
    const float gamma = (which ? bnk : bnq).weight[c];
    float dg = 0.f, db = 0.f, dbias = 0.f;
    for (int half = 0; half < 2; ++half) {
        const float* st = stat + ((size_t)which * 2 + half) * 2 * SA;
        const float mu = st[c], rs = st[SA + c];
        float s1 = 0.f, s2 = 0.f;
        for (int r = threadIdx.y; r < rows_half; r += 8) {
            const size_t off = (size_t)(half * rows_half + r) * SA + c;
            const float g = dy[off];
            s1 += g; s2 += g * (src[off] + bias - mu) * rs;
        }
        __syncthreads();
        sh[0][threadIdx.y][threadIdx.x] = s1; sh[1][threadIdx.y][threadIdx.x] = s2;
        __syncthreads();
        s1 = 0.f; s2 = 0.f;
        for (int i = 0; i < 8; ++i) { s1 += sh[0][i][threadIdx.x]; s2 += sh[1][i][threadIdx.x]; }
        dg += s2; db += s1;
        const float m1 = s1 / rows_half, m2 = s2 / rows_half;
        float colsum = 0.f;
        for (int r = threadIdx.y; r < rows_half; r += 8) {
            const size_t off = (size_t)(half * rows_half + r) * SA + c;
            const float v = gamma * rs * (dy[off] - m1 - (src[off] + bias - mu) * rs * m2);
            dy[off] = v;
            colsum += v;
        }
        __syncthreads();
        sh[0][threadIdx.y][threadIdx.x] = colsum;
        __syncthreads();
        for (int i = 0; i < 8; ++i) dbias += sh[0][i][threadIdx.x];
    }
    if (threadIdx.y == 0) {
        (which ? dgk : dgq)[c] = dg;
        (which ? dbk_bn : dbq_bn)[c] = db;
        (which ? dbias_k : dbias_q)[c] = dbias;
    }
}

// ------------------------------------------------------------------ PairLoss (pairloss.py:19-48)
// label[p][g] = (tar_probe[g] == tar_gallery[p]) (the reference's expand/eq order), loss = mean BCE (log clamped at -100 like
// torch.nn.BCELoss), prec = fraction of pairs with (score > 0.5) == label.   One block.
__global__ void __launch_bounds__(256) pair_loss_fwd_kernel(const float* __restrict__ score, const int64_t* __restrict__ tp, const int64_t* __restrict__ tg,
                                                            int n, float* __restrict__ loss, float* __restrict__ prec) {
    __shared__ float sh[8];
    float l = 0.f, ok = 0.f;
    for (int i = threadIdx.x; i < n * n; i += 256) {
        const int p = i / n, g = i - p * n;
        const float y = (tp[g] == tg[p]) ? 1.f : 0.f;
        const float s = score[i];
        l -= y * fmaxf(logf(s), -100.f) + (1.f - y) * fmaxf(logf(1.f - s), -100.f);
        ok += ((s > 1.f - s) ? 1.f : 0.f) == y ? 1.f : 0.f;
    }
    l = block_sum(l, sh);
    ok = block_sum(ok, sh);
    if (threadIdx.x == 0) { *loss = l / (float)(n * n); *prec = ok / (float)(n * n); }
}
__global__ void pair_loss_bwd_kernel(const float* __restrict__ score, const int64_t* __restrict__ tp, const int64_t* __restrict__ tg, int n,
                                     const float* __restrict__ d_loss, float* __restrict__ d_score) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n) return;
    const int p = i / n, g = i - p * n;
    const float y = (tp[g] == tg[p]) ? 1.f : 0.f;
    const float s = score[i];
    d_score[i] = *d_loss * (s - y) / fmaxf((1.f - s) * s, 1e-12f) / (float)(n * n);      // torch's binary_cross_entropy_backward
}

struct SiaWs {
    size_t qp, kp, stat, qh, kh, qnorm, knorm, W, cs, P, pnorm, cstat, s12, dout, dqb, dkb, total;
};
static SiaWs sia_layout(int n2, int T) {
    SiaWs w;
    const size_t R = (size_t)n2 * T;
    size_t off = 0;
    auto take = [&](size_t floats) { size_t o = off; off = align_up(off + floats * 4, 256); return o; };
    w.qp = take(R * SA); w.kp = take(R * SA); w.stat = take(2 * 2 * 2 * SA);
    w.qh = take(R * SA); w.kh = take(R * SA); w.qnorm = take(R); w.knorm = take(R);
    w.W = take(R * T); w.cs = take(R); w.P = take((size_t)n2 * SD); w.pnorm = take(n2);
    w.cstat = take(2 * SD); w.s12 = take(2 * SD); w.dout = take((size_t)n2 * SD);
    w.dqb = take(R * SA); w.dkb = take(R * SA);
    w.total = off;
    return w;
}

}  // namespace grl

using namespace grl;

extern "C" size_t grl_siamese_workspace_bytes(int n2, int T) {
    if (n2 <= 0 || (n2 & 1) || T <= 0 || T > 32) return 0;
    return sia_layout(n2, T).total;
}

#define SIA_F(name) ((float*)((char*)workspace + w.name))

extern "C" int grl_siamese_forward(grl_handle* h, const grl_siamese_params* p, const float* x, int n2, int T, int train, float* cls_encode,
                                   float* siamese_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return GRL_EINVAL;
    if (!p || !x || !cls_encode || !siamese_out || !workspace) return set_error(h, GRL_EINVAL, "grl_siamese_forward: NULL argument");
    if (n2 <= 0 || (n2 & 1)) return set_error(h, GRL_EINVAL, "grl_siamese_forward: the batch size should be even number! (got %d)", n2);
    if (T <= 0 || T > 32) return set_error(h, GRL_EINVAL, "grl_siamese_forward: need 0 < T <= 32");
    const SiaWs w = sia_layout(n2, T);
    if (workspace_bytes < w.total) return set_error(h, GRL_ENOMEM, "grl_siamese_forward: workspace %zu < %zu bytes", workspace_bytes, w.total);
    cudaStream_t st = (cudaStream_t)stream;
    const int n = n2 / 2, R = n2 * T, Rh = n * T;
    // the reference views x as [n, 2, T, D]: probe = samples 0, 2, 4, ..., gallery = 1, 3, 5, ...  The caller passes x already
    // regrouped as [probe clips | gallery clips] (the Python mirror does the index shuffle), so halves are contiguous here.
    sia_gemm_kernel<0, 0><<<dim3(SA / 64, (R + 63) / 64), 256, 0, st>>>(x, p->featQ_w, R, SA, SD, 0, SIA_F(qp));
    GRL_LAUNCH_CHECK(h);
    sia_gemm_kernel<0, 0><<<dim3(SA / 64, (R + 63) / 64), 256, 0, st>>>(x, p->featK_w, R, SA, SD, 0, SIA_F(kp));
    GRL_LAUNCH_CHECK(h);
    sia_bn_stats_kernel<<<dim3(SA / 32, 2), dim3(32, 8), 0, st>>>(SIA_F(qp), SIA_F(kp), p->featQ_b, p->featK_b, Rh, train, p->featQ_bn, p->featK_bn,
                                                                  SIA_F(stat));
    GRL_LAUNCH_CHECK(h);
    sia_bn_norm_kernel<<<dim3(R, 2), 256, 0, st>>>(SIA_F(qp), SIA_F(kp), p->featQ_b, p->featK_b, Rh, p->featQ_bn, p->featK_bn, SIA_F(stat),
                                                   SIA_F(qh), SIA_F(kh), SIA_F(qnorm), SIA_F(knorm));
    GRL_LAUNCH_CHECK(h);
    sia_attn_pool_kernel<<<n2, 256, 0, st>>>(x, SIA_F(qh), SIA_F(kh), T, SIA_F(W), SIA_F(cs), SIA_F(P), SIA_F(pnorm), siamese_out);
    GRL_LAUNCH_CHECK(h);
    sia_pair_stats_kernel<<<SD / 128, 128, 0, st>>>(siamese_out, n, train, p->cls_bn, SIA_F(cstat));
    GRL_LAUNCH_CHECK(h);
    sia_pair_cls_kernel<<<n * n, 256, 0, st>>>(siamese_out, n, p->cls_bn, SIA_F(cstat), p->cls_w, p->cls_b, cls_encode);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

extern "C" int grl_siamese_backward(grl_handle* h, const grl_siamese_params* p, const float* x, const float* siamese_out, int n2, int T,
                                    const float* d_cls_encode, const float* d_siamese_out, float* dx, const grl_siamese_grads* g,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return GRL_EINVAL;
    if (!p || !x || !siamese_out || !d_cls_encode || !dx || !g || !workspace) return set_error(h, GRL_EINVAL, "grl_siamese_backward: NULL argument");
    if (n2 <= 0 || (n2 & 1) || T <= 0 || T > 32) return set_error(h, GRL_EINVAL, "grl_siamese_backward: bad sizes");
    const SiaWs w = sia_layout(n2, T);
    if (workspace_bytes < w.total) return set_error(h, GRL_ENOMEM, "grl_siamese_backward: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = n2 / 2, R = n2 * T, Rh = n * T;
    sia_pair_bwd_reduce_kernel<<<SD / 128, 128, 0, st>>>(siamese_out, d_cls_encode, n, p->cls_bn, SIA_F(cstat), p->cls_w, SIA_F(s12),
                                                         g->cls_bn_w, g->cls_bn_b, g->cls_w);
    GRL_LAUNCH_CHECK(h);
    sia_dbc_kernel<<<1, 32, 0, st>>>(d_cls_encode, n * n, g->cls_b);
    GRL_LAUNCH_CHECK(h);
    sia_pair_bwd_apply_kernel<<<n2, 256, 0, st>>>(siamese_out, d_cls_encode, n, p->cls_bn, SIA_F(cstat), p->cls_w, SIA_F(s12), d_siamese_out,
                                                  SIA_F(dout));
    GRL_LAUNCH_CHECK(h);
    sia_attn_pool_bwd_kernel<<<n2, 256, 0, st>>>(x, SIA_F(qh), SIA_F(kh), SIA_F(qnorm), SIA_F(knorm), SIA_F(W), SIA_F(cs), siamese_out,
                                                 SIA_F(pnorm), SIA_F(dout), T, dx, SIA_F(dqb), SIA_F(dkb));
    GRL_LAUNCH_CHECK(h);
    sia_bn_bwd_kernel<<<dim3(SA / 32, 2), dim3(32, 8), 0, st>>>(SIA_F(qp), SIA_F(kp), p->featQ_b, p->featK_b, Rh, p->featQ_bn, p->featK_bn,
                                                                SIA_F(stat), SIA_F(dqb), SIA_F(dkb), g->featQ_bn_w, g->featQ_bn_b, g->featK_bn_w,
                                                                g->featK_bn_b, g->featQ_b, g->featK_b);
    GRL_LAUNCH_CHECK(h);
    // d W = d Qp^T x   [SA][SD], contraction over the R rows
    sia_gemm_kernel<1, 1><<<dim3(SD / 64, SA / 64), 256, 0, st>>>(SIA_F(dqb), x, SA, SD, R, 0, g->featQ_w);
    GRL_LAUNCH_CHECK(h);
    sia_gemm_kernel<1, 1><<<dim3(SD / 64, SA / 64), 256, 0, st>>>(SIA_F(dkb), x, SA, SD, R, 0, g->featK_w);
    GRL_LAUNCH_CHECK(h);
    // d x += d Qp Wq + d Kp Wk   [R][SD], contraction over SA
    sia_gemm_kernel<0, 1><<<dim3(SD / 64, (R + 63) / 64), 256, 0, st>>>(SIA_F(dqb), p->featQ_w, R, SD, SA, 1, dx);
    GRL_LAUNCH_CHECK(h);
    sia_gemm_kernel<0, 1><<<dim3(SD / 64, (R + 63) / 64), 256, 0, st>>>(SIA_F(dkb), p->featK_w, R, SD, SA, 1, dx);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

extern "C" int grl_pair_loss_forward(grl_handle* h, const float* score, const int64_t* tar_probe, const int64_t* tar_gallery, int n,
                                     float* loss, float* prec, void* stream) {
    if (!h) return GRL_EINVAL;
    if (!score || !tar_probe || !tar_gallery || !loss || !prec || n <= 0) return set_error(h, GRL_EINVAL, "grl_pair_loss_forward: bad argument");
    pair_loss_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(score, tar_probe, tar_gallery, n, loss, prec);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

extern "C" int grl_pair_loss_backward(grl_handle* h, const float* score, const int64_t* tar_probe, const int64_t* tar_gallery, int n,
                                      const float* d_loss, float* d_score, void* stream) {
    if (!h) return GRL_EINVAL;
    if (!score || !tar_probe || !tar_gallery || !d_loss || !d_score || n <= 0) return set_error(h, GRL_EINVAL, "grl_pair_loss_backward: bad argument");
    pair_loss_bwd_kernel<<<(n * n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(score, tar_probe, tar_gallery, n, d_loss, d_score);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}
