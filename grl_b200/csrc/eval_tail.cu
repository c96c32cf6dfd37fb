// grl_b200 — eval feature tail (sm_100a): head outputs -> the 6144-d per-clip descriptor the evaluator matches on.
//
// Reference lines replaced (all in eval mode, running-statistics BatchNorm):
//   corr_bn / uncorr_bn + F.normalize                       reid/models/grl_model.py:222-226
//   Siamese.self_attention (Q/K Linear + BN + L2, softmax(QK^T), weights . V, sum, L2)   reid/models/Siamese.py:79-106
//   out_feat = cat(x_uncorr, out_frame, feats_corr.mean(1))  reid/evaluator/attevaluator.py:79-80
// The two 2048 -> 512 projections are one split-bf16 tcgen05 GEMM ([n*T] x 1024 x 2048, stacked Q|K weights);
// everything else is a handful of row-wise reductions over 2048 channels.
#include "api.h"

namespace grl {

constexpr int TC = 2048;      // feature channels
constexpr int TA = 512;       // attention width (Siamese output_num)
constexpr float TAIL_BN_EPS = 1e-5f;

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane_id() == 0) sh[warp] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sh[i];
    return s;
}

// y = normalize(bn_eval(x)) over the 2048 channels of one row; one block (256 threads, 8 channels each) per row.
// rows [0, n_corr) use the corr_bn parameters and go to xc; rows [n_corr, n_corr + n_unc) use uncorr_bn and go to
// out[:, 0:2048] (leading dimension ld_out).  apply_bn = 0: inputs are already BN'd + normalised (plain copy).
__global__ void __launch_bounds__(256) tail_bn_norm_kernel(const float* __restrict__ f_corr, const float* __restrict__ f_unc, int n_corr, int n_unc,
                                                           grl_bn_params bn_c, grl_bn_params bn_u, int apply_bn, float* __restrict__ xc,
                                                           float* __restrict__ out, long long ld_out) {
    __shared__ float sh[8];
    const int row = blockIdx.x;
    const bool corr = row < n_corr;
    const float* src = corr ? f_corr + (size_t)row * TC : f_unc + (size_t)(row - n_corr) * TC;
    float* dst = corr ? xc + (size_t)row * TC : out + (size_t)(row - n_corr) * ld_out;
    const grl_bn_params& bn = corr ? bn_c : bn_u;
    const int c = threadIdx.x * 8;
    float v[8];
    {
        const float4 a = *reinterpret_cast<const float4*>(src + c), b = *reinterpret_cast<const float4*>(src + c + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    if (apply_bn) {
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float a = bn.weight[c + i] * rsqrtf(bn.running_var[c + i] + TAIL_BN_EPS);
            v[i] = a * (v[i] - bn.running_mean[c + i]) + bn.bias[c + i];
            ss += v[i] * v[i];
        }
        ss = block_sum_256(ss, sh);
        const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);        // F.normalize: x / max(||x||, eps)
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= inv;
    }
    *reinterpret_cast<float4*>(dst + c) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// stacked projection weights [Q | K] -> planes, stacked bias
__global__ void tail_stack_bias_kernel(const float* __restrict__ bq, const float* __restrict__ bk, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < TA) { out[i] = bq[i]; out[TA + i] = bk[i]; }
}

// One block per clip.  qk [n*T][1024] = raw projections (Q | K, bias added).  BN(eval) + L2-normalise every Q_t / K_t,
// logits = Q K^T (T x T), softmax over the key axis, w[j] = sum_i softmax[i][j], pooled = normalize(sum_j w[j] x_j),
// out[n][2048:4096] = pooled, out[n][4096:6144] = mean_t x_t.      dynamic smem: 2*T*512 + T*T + T floats
__global__ void __launch_bounds__(256) tail_attention_kernel(const float* __restrict__ qk, const float* __restrict__ xc, int T,
                                                             grl_bn_params bn_q, grl_bn_params bn_k, float* __restrict__ out,
                                                             long long ld_out) {
    extern __shared__ float smem[];
    __shared__ float sh[8];
    float* Q = smem;                         // [T][512]
    float* K = Q + (size_t)T * TA;           // [T][512]
    float* L = K + (size_t)T * TA;           // [T][T] logits -> softmax
    float* W = L + T * T;                    // [T] column sums
    const int n = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // BN + L2 norm: one warp per (t, Q|K) vector of 512
    for (int v = warp; v < 2 * T; v += 8) {
        const int t = v >> 1, isk = v & 1;
        const grl_bn_params& bn = isk ? bn_k : bn_q;
        const float* src = qk + ((size_t)n * T + t) * (2 * TA) + isk * TA;
        float* dst = (isk ? K : Q) + (size_t)t * TA;
        float vals[16], ss = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int c = lane + 32 * i;
            const float a = bn.weight[c] * rsqrtf(bn.running_var[c] + TAIL_BN_EPS);
            vals[i] = a * (src[c] - bn.running_mean[c]) + bn.bias[c];
            ss += vals[i] * vals[i];
        }
        ss = warp_sum(ss);
        const float inv = 1.f / sqrtf(ss);                       // Siamese.py:88: plain division by the norm
#pragma unroll
        for (int i = 0; i < 16; ++i) dst[lane + 32 * i] = vals[i] * inv;
    }
    __syncthreads();
    for (int e = warp; e < T * T; e += 8) {                      // logits[i][j] = Q_i . K_j
        const int i = e / T, j = e - i * T;
        float acc = 0.f;
#pragma unroll 4
        for (int c = lane; c < TA; c += 32) acc += Q[(size_t)i * TA + c] * K[(size_t)j * TA + c];
        acc = warp_sum(acc);
        if (lane == 0) L[e] = acc;
    }
    __syncthreads();
    if (threadIdx.x < T) {                                       // softmax over j for row i
        float* r = L + threadIdx.x * T;
        float mx = r[0];
        for (int j = 1; j < T; ++j) mx = fmaxf(mx, r[j]);
        float s = 0.f;
        for (int j = 0; j < T; ++j) { r[j] = expf(r[j] - mx); s += r[j]; }
        const float inv = 1.f / s;
        for (int j = 0; j < T; ++j) r[j] *= inv;
    }
    __syncthreads();
    if (threadIdx.x < T) {                                       // (weights @ V).sum(1) == (column sums of weights) @ V
        float s = 0.f;
        for (int i = 0; i < T; ++i) s += L[i * T + threadIdx.x];
        W[threadIdx.x] = s;
    }
    __syncthreads();
    const int c = threadIdx.x * 8;
    float pooled[8] = {0, 0, 0, 0, 0, 0, 0, 0}, mean[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < T; ++t) {
        const float* x = xc + ((size_t)n * T + t) * TC + c;
        const float4 a = *reinterpret_cast<const float4*>(x), b = *reinterpret_cast<const float4*>(x + 4);
        const float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const float wt = W[t];
#pragma unroll
        for (int i = 0; i < 8; ++i) { pooled[i] += wt * xv[i]; mean[i] += xv[i]; }
    }
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += pooled[i] * pooled[i];
    ss = block_sum_256(ss, sh);
    const float inv = 1.f / sqrtf(ss), invT = 1.f / (float)T;
    float* o = out + (size_t)n * ld_out;
    *reinterpret_cast<float4*>(o + TC + c) = make_float4(pooled[0] * inv, pooled[1] * inv, pooled[2] * inv, pooled[3] * inv);
    *reinterpret_cast<float4*>(o + TC + c + 4) = make_float4(pooled[4] * inv, pooled[5] * inv, pooled[6] * inv, pooled[7] * inv);
    *reinterpret_cast<float4*>(o + 2 * TC + c) = make_float4(mean[0] * invT, mean[1] * invT, mean[2] * invT, mean[3] * invT);
    *reinterpret_cast<float4*>(o + 2 * TC + c + 4) = make_float4(mean[4] * invT, mean[5] * invT, mean[6] * invT, mean[7] * invT);
}

static void tail_ws_layout(int n, int T, size_t* xc, size_t* xpl, size_t* wpl, size_t* qk, size_t* bias) {
    const size_t rows = (size_t)n * T;
    *xc = align_up(rows * TC * 4, 1024);
    *xpl = align_up(rows * TC * 2, 1024);
    *wpl = align_up((size_t)2 * TA * TC * 2, 1024);
    *qk = align_up(rows * 2 * TA * 4, 1024);
    *bias = align_up((size_t)2 * TA * 4, 1024);
}

}  // namespace grl

using namespace grl;

extern "C" size_t grl_eval_descriptor_workspace_bytes(int n, int T) {
    if (n <= 0 || T <= 0) return 0;
    size_t a, b, c, d, e;
    tail_ws_layout(n, T, &a, &b, &c, &d, &e);
    return a + 2 * b + 2 * c + d + e;
}

extern "C" int grl_eval_descriptor(grl_handle* h, const grl_tail_params* p, const float* f_uncorr, const float* f_corr, int n, int T,
                                   int apply_tail_bn, float* out, long long ld_out, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    if (!h || !p || !f_uncorr || !f_corr || !out || !workspace) return set_error(h, GRL_EINVAL, "grl_eval_descriptor: NULL argument");
    if (n <= 0 || T <= 0 || T > 64 || ld_out < 3 * TC || (ld_out & 3)) return set_error(h, GRL_EINVAL, "grl_eval_descriptor: need n > 0, 0 < T <= 64, ld_out >= 6144");
    const void* req[] = {p->featQ_w, p->featQ_b, p->featK_w, p->featK_b, p->featQ_bn.weight, p->featQ_bn.bias, p->featQ_bn.running_mean,
                         p->featQ_bn.running_var, p->featK_bn.weight, p->featK_bn.bias, p->featK_bn.running_mean, p->featK_bn.running_var};
    for (const void* q : req)
        if (!q) return set_error(h, GRL_EINVAL, "grl_eval_descriptor: NULL attention parameter");
    if (apply_tail_bn) {
        const void* r2[] = {p->corr_bn.weight, p->corr_bn.bias, p->corr_bn.running_mean, p->corr_bn.running_var,
                            p->uncorr_bn.weight, p->uncorr_bn.bias, p->uncorr_bn.running_mean, p->uncorr_bn.running_var};
        for (const void* q : r2)
            if (!q) return set_error(h, GRL_EINVAL, "grl_eval_descriptor: NULL corr_bn / uncorr_bn parameter");
    }
    if (workspace_bytes < grl_eval_descriptor_workspace_bytes(n, T)) return set_error(h, GRL_ENOMEM, "grl_eval_descriptor: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    size_t b_xc, b_xpl, b_wpl, b_qk, b_bias;
    tail_ws_layout(n, T, &b_xc, &b_xpl, &b_wpl, &b_qk, &b_bias);
    uint8_t* w = (uint8_t*)workspace;
    float* xc = (float*)w; w += b_xc;
    __nv_bfloat16* x_hi = (__nv_bfloat16*)w; w += b_xpl;
    __nv_bfloat16* x_lo = (__nv_bfloat16*)w; w += b_xpl;
    __nv_bfloat16* w_hi = (__nv_bfloat16*)w; w += b_wpl;
    __nv_bfloat16* w_lo = (__nv_bfloat16*)w; w += b_wpl;
    float* qk = (float*)w; w += b_qk;
    float* bias = (float*)w;
    const int rows = n * T;
    tail_bn_norm_kernel<<<rows + n, 256, 0, st>>>(f_corr, f_uncorr, rows, n, p->corr_bn, p->uncorr_bn, apply_tail_bn ? 1 : 0, xc, out, ld_out);
    GRL_LAUNCH_CHECK(h);
    GRL_TRY(split_planes(h, st, xc, TC, x_hi, x_lo, TC, rows, TC));
    GRL_TRY(split_planes(h, st, p->featQ_w, TC, w_hi, w_lo, TC, TA, TC));
    GRL_TRY(split_planes(h, st, p->featK_w, TC, w_hi + (size_t)TA * TC, w_lo + (size_t)TA * TC, TC, TA, TC));
    tail_stack_bias_kernel<<<(TA + 255) / 256, 256, 0, st>>>(p->featQ_b, p->featK_b, bias);
    GRL_LAUNCH_CHECK(h);
    {
        GemmEpi e = epi_default();
        e.C = qk; e.ldc = 2 * TA;
        e.col_bias = bias;
        Operand a{x_hi, x_lo, TC, 0, 0}, b{w_hi, w_lo, TC, 0, 0};
        GRL_TRY(gemm_launch(h, st, rows, 2 * TA, TC, 1, a, b, e, 128));
    }
    const size_t smem = ((size_t)2 * T * TA + (size_t)T * T + T) * sizeof(float);
    if (smem > 220 * 1024) return set_error(h, GRL_EINVAL, "grl_eval_descriptor: T = %d needs %zu bytes of shared memory", T, smem);
    GRL_TRY(ensure_dyn_smem(h, (const void*)tail_attention_kernel, (int)smem));
    tail_attention_kernel<<<n, 256, smem, st>>>(qk, xc, T, p->featQ_bn, p->featK_bn, out, ld_out);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}
