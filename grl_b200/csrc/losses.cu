// grl_b200 — loss neighbours of the head, sm_100a.
//
//   OIMLoss        reid/loss/oim.py:8-58      logits = scalar * x @ lut^T, mean cross-entropy; backward: dx through the
//                                             softmax, THEN the look-up table's momentum update, sample by sample in order
//   TripletLoss    reid/loss/triplet.py:5-90  ('soft' | float margin, batch_hard=True, mode='id', dis_func='eu', n_dis=0)
//
// All of it is B x D / B x C sized work (B <= a few hundred rows, D = 2048, C = 625 identities): latency-bound, so the
// kernels are plain fp32 CUDA-core code with deterministic reduction orders.  The legacy `OIM(autograd.Function)` of the
// reference cannot execute on current PyTorch (non-static forward); these entry points are its forward/backward contract.
#include <math_constants.h>

#include "api.h"

namespace grl {

// ------------------------------------------------------------------ small fp32 GEMM: C[m][n] = alpha * sum_k A(m,k) * B(n,k)
// MODE 0: A = a[m][k] row-major.  MODE 1 (OIM backward): A(m,k) = probs[m][k] - (k == target[m]), B(n,k) = lut[k][n].
template <int MODE>
__global__ void __launch_bounds__(256) loss_gemm_kernel(const float* __restrict__ a, const float* __restrict__ b, const int64_t* __restrict__ targets,
                                                        int M, int N, int K, const float* __restrict__ alpha_dev, float alpha, float* __restrict__ c) {
    __shared__ float sa[16][65], sb[16][65];
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int t = threadIdx.x; t < 64 * 16; t += 256) {
            if (MODE == 0) {
                const int r = t >> 4, kk = t & 15;              // k fastest: both operands are k-contiguous
                const int m = m0 + r, n = n0 + r, k = k0 + kk;
                sa[kk][r] = (m < M && k < K) ? a[(long long)m * K + k] : 0.f;
                sb[kk][r] = (n < N && k < K) ? b[(long long)n * K + k] : 0.f;
            } else {
                const int kk = t >> 6, r = t & 63;              // lut rows are n-contiguous, probs rows k-contiguous
                const int n = n0 + r, k = k0 + kk;
                sb[kk][r] = (n < N && k < K) ? b[(long long)k * N + n] : 0.f;
                const int r2 = t >> 4, kk2 = t & 15, m2 = m0 + r2, k2 = k0 + kk2;
                sa[kk2][r2] = (m2 < M && k2 < K) ? a[(long long)m2 * K + k2] - ((int64_t)k2 == targets[m2] ? 1.f : 0.f) : 0.f;
            }
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
    const float al = alpha * (alpha_dev ? *alpha_dev : 1.f);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) c[(long long)m * N + n] = al * acc[i][j];
        }
}

// ------------------------------------------------------------------ OIM: softmax cross-entropy rows, mean
__global__ void __launch_bounds__(256) oim_ce_rows_kernel(const float* __restrict__ logits, const int64_t* __restrict__ targets, int C,
                                                          float* __restrict__ probs, float* __restrict__ row_loss) {
    __shared__ float red[8];
    const int b = blockIdx.x;
    const float* l = logits + (long long)b * C;
    float mx = -CUDART_INF_F;
    for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, l[c]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if (lane_id() == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s += expf(l[c] - mx);
    s = warp_sum(s);
    if (lane_id() == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    const float lse = mx + logf(s);
    for (int c = threadIdx.x; c < C; c += blockDim.x) probs[(long long)b * C + c] = expf(l[c] - lse);
    if (threadIdx.x == 0) row_loss[b] = lse - l[targets[b]];
}
__global__ void mean_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {      // one block, fixed order
    __shared__ float red[8];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
    s = warp_sum(s);
    if (lane_id() == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        *out = t / (float)n;
    }
}

// ------------------------------------------------------------------ OIM: look-up table update (oim.py:24-26)
// for x, y in zip(inputs, targets): lut[y] = m * lut[y] + (1 - m) * x; lut[y] /= lut[y].norm()
// Samples of different identities touch different rows: one block per identity applies ITS samples in batch order.
__global__ void __launch_bounds__(256) oim_lut_update_kernel(const float* __restrict__ x, const int64_t* __restrict__ targets, int B, int D,
                                                             float momentum, float* __restrict__ lut) {
    __shared__ float red[8];
    __shared__ float s_norm;
    const int y = blockIdx.x;
    float* row = lut + (long long)y * D;
    for (int i = 0; i < B; ++i) {
        if (targets[i] != y) continue;                           // block-uniform
        const float* xi = x + (long long)i * D;
        float ss = 0.f;
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const float v = momentum * row[d] + (1.f - momentum) * xi[d];
            row[d] = v;
            ss += v * v;
        }
        ss = warp_sum(ss);
        if (lane_id() == 0) red[threadIdx.x >> 5] = ss;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int w = 0; w < 8; ++w) t += red[w];
            s_norm = sqrtf(t);
        }
        __syncthreads();
        const float nrm = s_norm;
        for (int d = threadIdx.x; d < D; d += blockDim.x) row[d] = row[d] / nrm;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ batch-hard triplet (triplet.py:16-90)
// dist[i][j] = sqrt(sum (a_i - a_j)^2 + 1e-12); max_positive = max_j dist*posmask, min_negative = min_j dist + 1e5*same_id;
// z = max_positive - min_negative; loss = log(1 + exp(z)) ('soft') | clamp(z + margin, 0).
__global__ void __launch_bounds__(256) triplet_fwd_kernel(const float* __restrict__ feat, const int64_t* __restrict__ ids, int B, int D, int soft,
                                                          float margin, float* __restrict__ loss, float* __restrict__ zout,
                                                          int32_t* __restrict__ pos_idx, int32_t* __restrict__ neg_idx,
                                                          float* __restrict__ pos_d, float* __restrict__ neg_d) {
    extern __shared__ float tsm[];                               // a_i [D] | dist row [B]
    float* ai = tsm;
    float* drow = tsm + D;
    const int i = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += blockDim.x) ai[d] = feat[(long long)i * D + d];
    __syncthreads();
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = warp; j < B; j += nw) {
        const float* aj = feat + (long long)j * D;
        float s = 0.f;
        for (int d = lane_id(); d < D; d += 32) { const float t = ai[d] - aj[d]; s = fmaf(t, t, s); }
        s = warp_sum(s);
        if (lane_id() == 0) drow[j] = sqrtf(s + 1e-12f);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int64_t idi = ids[i];
        float mp = -CUDART_INF_F, mn = CUDART_INF_F;
        int jp = 0, jn = 0;
        for (int j = 0; j < B; ++j) {                            // first maximum / minimum wins, like torch's reductions
            const bool same = ids[j] == idi;
            const float vp = (same && j != i) ? drow[j] : 0.f;   // dist * positive_mask
            const float vn = drow[j] + (same ? 1e5f : 0.f);      // dist + 1e5 * same_id_mask
            if (vp > mp) { mp = vp; jp = j; }
            if (vn < mn) { mn = vn; jn = j; }
        }
        const float z = mp - mn;
        zout[i] = z;
        loss[i] = soft ? logf(1.f + expf(z)) : fmaxf(z + margin, 0.f);   // literally torch.log(1 + torch.exp(z)), :83
        const bool pos_is_real = (ids[jp] == idi) && jp != i;    // else the masked product is identically 0: no gradient
        pos_idx[i] = pos_is_real ? jp : -1;
        neg_idx[i] = jn;
        pos_d[i] = drow[jp];
        neg_d[i] = drow[jn];
    }
}
// d feat[m] = sum over rows i of the terms that touch a_m, in row order (deterministic, no atomics)
__global__ void __launch_bounds__(256) triplet_bwd_kernel(const float* __restrict__ feat, int B, int D, int soft, float margin,
                                                          const float* __restrict__ z, const int32_t* __restrict__ pos_idx,
                                                          const int32_t* __restrict__ neg_idx, const float* __restrict__ pos_d,
                                                          const float* __restrict__ neg_d, const float* __restrict__ d_loss,
                                                          float* __restrict__ dfeat) {
    const int m = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float am = feat[(long long)m * D + d];
        float g = 0.f;
        for (int i = 0; i < B; ++i) {
            const float dz = d_loss[i] * (soft ? 1.f / (1.f + expf(-z[i])) : (z[i] + margin >= 0.f ? 1.f : 0.f));
            const int p = pos_idx[i], n = neg_idx[i];
            if (i == m) {
                if (p >= 0) g += dz * (am - feat[(long long)p * D + d]) / pos_d[i];
                g -= dz * (am - feat[(long long)n * D + d]) / neg_d[i];
            }
            if (p == m && i != m) g -= dz * (feat[(long long)i * D + d] - am) / pos_d[i];
            if (n == m && i != m) g += dz * (feat[(long long)i * D + d] - am) / neg_d[i];
        }
        dfeat[(long long)m * D + d] = g;
    }
}

}  // namespace grl

using namespace grl;

extern "C" int grl_oim_forward(grl_handle* h, const float* x, const int64_t* targets, const float* lut, int B, int C, int D, float scalar,
                               float* logits, float* probs, float* row_loss, float* loss, void* stream) {
    if (!h) return GRL_EINVAL;
    if (!x || !targets || !lut || !logits || !probs || !row_loss || !loss) return set_error(h, GRL_EINVAL, "grl_oim_forward: NULL argument");
    if (B <= 0 || C <= 0 || D <= 0) return set_error(h, GRL_EINVAL, "grl_oim_forward: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    loss_gemm_kernel<0><<<dim3((C + 63) / 64, (B + 63) / 64), 256, 0, st>>>(x, lut, nullptr, B, C, D, nullptr, scalar, logits);
    GRL_LAUNCH_CHECK(h);
    oim_ce_rows_kernel<<<B, 256, 0, st>>>(logits, targets, C, probs, row_loss);
    GRL_LAUNCH_CHECK(h);
    mean_kernel<<<1, 256, 0, st>>>(row_loss, B, loss);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

extern "C" int grl_oim_backward(grl_handle* h, const float* x, const int64_t* targets, float* lut, const float* probs, int B, int C, int D,
                                float scalar, float momentum, const float* d_loss, float* dx, void* stream) {
    if (!h) return GRL_EINVAL;
    if (!x || !targets || !lut || !probs || !d_loss) return set_error(h, GRL_EINVAL, "grl_oim_backward: NULL argument");
    if (B <= 0 || C <= 0 || D <= 0) return set_error(h, GRL_EINVAL, "grl_oim_backward: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (dx) {   // grad_inputs = grad_outputs.mm(lut) with grad_outputs = d_loss * scalar / B * (softmax - onehot), BEFORE the update
        loss_gemm_kernel<1><<<dim3((D + 63) / 64, (B + 63) / 64), 256, 0, st>>>(probs, lut, targets, B, D, C, d_loss, scalar / (float)B, dx);
        GRL_LAUNCH_CHECK(h);
    }
    oim_lut_update_kernel<<<C, 256, 0, st>>>(x, targets, B, D, momentum, lut);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

extern "C" int grl_triplet_forward(grl_handle* h, const float* feat, const int64_t* ids, int B, int D, int soft, float margin, float* loss,
                                   float* z, int32_t* pos_idx, int32_t* neg_idx, float* pos_d, float* neg_d, void* stream) {
    if (!h) return GRL_EINVAL;
    if (!feat || !ids || !loss || !z || !pos_idx || !neg_idx || !pos_d || !neg_d) return set_error(h, GRL_EINVAL, "grl_triplet_forward: NULL argument");
    if (B <= 0 || D <= 0 || (size_t)(D + B) * 4 > 200 * 1024) return set_error(h, GRL_EINVAL, "grl_triplet_forward: bad sizes (B=%d, D=%d)", B, D);
    const size_t smem = (size_t)(D + B) * 4;
    GRL_TRY(ensure_dyn_smem(h, (const void*)triplet_fwd_kernel, (int)smem));
    triplet_fwd_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(feat, ids, B, D, soft, margin, loss, z, pos_idx, neg_idx, pos_d, neg_d);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}

extern "C" int grl_triplet_backward(grl_handle* h, const float* feat, int B, int D, int soft, float margin, const float* z,
                                    const int32_t* pos_idx, const int32_t* neg_idx, const float* pos_d, const float* neg_d,
                                    const float* d_loss, float* dfeat, void* stream) {
    if (!h) return GRL_EINVAL;
    if (!feat || !z || !pos_idx || !neg_idx || !pos_d || !neg_d || !d_loss || !dfeat) return set_error(h, GRL_EINVAL, "grl_triplet_backward: NULL argument");
    if (B <= 0 || D <= 0) return set_error(h, GRL_EINVAL, "grl_triplet_backward: bad sizes");
    triplet_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(feat, B, D, soft, margin, z, pos_idx, neg_idx, pos_d, neg_d, d_loss, dfeat);
    GRL_LAUNCH_CHECK(h);
    return GRL_OK;
}
