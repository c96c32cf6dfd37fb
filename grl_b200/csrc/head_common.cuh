// grl_b200 — GCE + TRL head: workspace layout and device helpers shared by forward / backward.
//
// Data layout in HBM (everything internal is PIXEL-MAJOR):
//   P = B*T*128 pixels, row p = (b, t, s);  activations are [P][C] with C contiguous, so a 1x1
//   conv is a plain GEMM  Y[P][Cout] = X[P][Cin] * W[Cout][Cin]^T  and TMA boxes are 128-byte rows.
//   GEMM operands are stored as two bf16 planes (hi, lo) — same bytes as one fp32 copy.
//   One TRL step works on R = B*128 rows (frame tau of every clip); both temporal directions are
//   batched as z = 0 (forward), 1 (backward) in every launch.
#pragma once
#include "api.h"

namespace grl {

constexpr int HC = 2048;    // layer4 channels
constexpr int HG = 1024;    // glo_fc / corr_atte.0 width
constexpr int HMID = 256;   // corr_atte.2 width
constexpr int HB = 512;     // BasicBlock planes
constexpr int HSE = 128;    // channel attention bottleneck
constexpr int HS = 128;     // 16 x 8 pixels per frame
constexpr float BN_EPS = 1e-5f;
constexpr float BN_MOM = 0.1f;

// ------------------------------------------------------------------ workspace
#define GRL_HEAD_BUFFERS(X)                                                                       \
    /* --- GCE --- */                                                                             \
    X(xp_hi, 2, (size_t)P * HC) X(xp_lo, 2, (size_t)P * HC)                                       \
    X(gx, 4, (size_t)N * HC) X(g, 4, (size_t)B * HC) X(u, 4, (size_t)B * HG) X(glo, 4, (size_t)B * HG) \
    X(bias1, 4, (size_t)B * HG) X(glo_stat, 4, (size_t)4 * HG)                                     \
    X(part_a, 4, PART) X(part_b, 4, PART) X(part_c, 4, PART)                                      \
    X(y1_hi, 2, (size_t)P * HG) X(y1_lo, 2, (size_t)P * HG) X(bn1_stat, 4, (size_t)4 * HG)        \
    X(w2s_hi, 2, (size_t)HMID * HG) X(w2s_lo, 2, (size_t)HMID * HG) X(bias2, 4, HMID)             \
    X(y2, 4, (size_t)P * HMID) X(bn2_stat, 4, (size_t)4 * HMID)                                   \
    X(y3, 4, (size_t)P) X(bn3_stat, 4, 8) X(m, 4, (size_t)P)                                      \
    X(xc_hi, 2, (size_t)P * HC) X(xc_lo, 2, (size_t)P * HC) X(xu_hi, 2, (size_t)P * HC) X(xu_lo, 2, (size_t)P * HC) \
    X(gc, 4, (size_t)N * HC) X(f2, 4, (size_t)P * 2 * HC)                                         \
    /* --- weight planes --- */                                                                   \
    X(w1a_hi, 2, (size_t)HG * HC) X(w1a_lo, 2, (size_t)HG * HC) X(w2_hi, 2, (size_t)HMID * HG) X(w2_lo, 2, (size_t)HMID * HG) \
    X(wf2_hi, 2, (size_t)2 * HC * HC) X(wf2_lo, 2, (size_t)2 * HC * HC) X(wf1_hi, 2, (size_t)2 * HC * HC) X(wf1_lo, 2, (size_t)2 * HC * HC) \
    X(wc1_hi, 2, (size_t)2 * HB * HC) X(wc1_lo, 2, (size_t)2 * HB * HC) X(wc2_hi, 2, (size_t)2 * HB * HB) X(wc2_lo, 2, (size_t)2 * HB * HB) \
    X(wc3_hi, 2, (size_t)2 * HC * HB) X(wc3_lo, 2, (size_t)2 * HC * HB) X(bf2cat, 4, (size_t)2 * HC) X(bf1cat, 4, (size_t)2 * HC) \
    X(wg_hi, 2, (size_t)HG * HC) X(wg_lo, 2, (size_t)HG * HC) X(w1b_hi, 2, (size_t)HG * HG) X(w1b_lo, 2, (size_t)HG * HG) \
    X(sa_hi, 2, (size_t)B * HC) X(sa_lo, 2, (size_t)B * HC)                                       \
    /* --- TRL per-step state (SL = T slots when saving for backward, else 1/2) --- */            \
    X(mem_hi, 2, (size_t)SLM * 2 * R * HC) X(mem_lo, 2, (size_t)SLM * 2 * R * HC)                 \
    X(z_hi, 2, (size_t)SLZ * 2 * R * HC) X(z_lo, 2, (size_t)SLZ * 2 * R * HC)                     \
    X(f1, 4, (size_t)SL * 2 * R * HC) X(qpart, 4, (size_t)T * 2 * 4 * B * HC)                     \
    X(se_q, 4, (size_t)T * 2 * B * HC) X(se_a, 4, (size_t)T * 2 * B * HC) X(se_h, 4, (size_t)T * 2 * B * HSE) \
    X(h1, 4, (size_t)SL * 2 * R * HB) X(h1p_hi, 2, (size_t)SL * 2 * R * HB) X(h1p_lo, 2, (size_t)SL * 2 * R * HB) \
    X(h2, 4, (size_t)SL * 2 * R * HB) X(h2p_hi, 2, (size_t)SL * 2 * R * HB) X(h2p_lo, 2, (size_t)SL * 2 * R * HB) \
    X(h3, 4, (size_t)SL * 2 * R * HC)                                                             \
    X(sbn1, 4, (size_t)SL * 2 * 4 * HB) X(sbn2, 4, (size_t)SL * 2 * 4 * HB) X(sbn3, 4, (size_t)SL * 2 * 4 * HC) \
    X(out_d, 4, (size_t)2 * N * HC)                                                               \
    /* --- backward scratch (only when saving for backward) --- */                                \
    X(dmem, 4, BW * (T + 1) * 2 * R * HC) X(dh3_hi, 2, BW * T * 2 * R * HC) X(dh3_lo, 2, BW * T * 2 * R * HC) \
    X(dh2p, 4, BW * 2 * R * HB) X(dh2_hi, 2, BW * T * 2 * R * HB) X(dh2_lo, 2, BW * T * 2 * R * HB) \
    X(dh1p, 4, BW * 2 * R * HB) X(dh1_hi, 2, BW * T * 2 * R * HB) X(dh1_lo, 2, BW * T * 2 * R * HB) \
    X(dzc, 4, BW * 2 * R * HC)                                                                      \
    X(dxu, 4, BW * 2 * P * HC) X(dxc, 4, BW * P * HC) X(dgc, 4, BW * N * HC)                       \
    X(kcoef, 4, BW * 2 * 3 * HC) X(se_ds, 4, BW * T * 2 * B * HC) X(se_dh, 4, BW * T * 2 * B * HSE) X(se_dq, 4, BW * T * 2 * B * HC) \
    X(dbf1_part, 4, BW * T * 2 * B * HC) X(dbf2_part, 4, BW * T * 2 * B * HC)                      \
    X(gw_f1, 4, BW * 2 * HC * HC) X(gw_f2, 4, BW * 2 * HC * HC) X(gw_c1, 4, BW * 2 * HB * HC)      \
    X(gw_c2, 4, BW * 2 * HB * HB) X(gw_c3, 4, BW * 2 * HC * HB)                                    \
    X(gbn1, 4, BW * 2 * 2 * HB) X(gbn2, 4, BW * 2 * 2 * HB) X(gbn3, 4, BW * 2 * 2 * HC)            \
    X(dm, 4, BW * P) X(dy3, 4, BW * P) X(dy2_hi, 2, BW * P * HMID) X(dy2_lo, 2, BW * P * HMID)     \
    X(g2, 4, BW * 16 * HMID * HG) X(dz1, 4, BW * P * HG) X(dy1_hi, 2, BW * P * HG) X(dy1_lo, 2, BW * P * HG) \
    X(dbias1_part, 4, BW * N * HG) X(dbias1, 4, BW * B * HG) X(du, 4, BW * B * HG) X(dg, 4, BW * B * HC) \
    X(gsmall, 4, BW * 8192)                                                                        \
    /* --- single-plane fp16 operands of the f1 / f2 gradient GEMMs (one MMA per k-step) + their device-side scales --- */ \
    X(df1_16, 2, BW * 2 * R * HC) X(mem_16, 2, BW * SLM * 2 * R * HC) X(df2_16, 2, BW * P * 2 * HC) X(xc_16, 2, BW * P * HC) \
    X(wf1_16, 2, BW * 2 * HC * HC) X(wf2_16, 2, BW * 2 * HC * HC) X(f16_scal, 4, BW * (64 + 12 * T))  \
    /* --- ... and of the memory block's WEIGHT gradients: activations (written by the training forward), gradients (bn_bwd_apply) --- */ \
    X(h1p_16, 2, BW * SL * 2 * R * HB) X(h2p_16, 2, BW * SL * 2 * R * HB) X(z_16, 2, BW * SLZ * 2 * R * HC)            \
    X(dh3_16, 2, BW * T * 2 * R * HC) X(dh2_16, 2, BW * T * 2 * R * HB) X(dh1_16, 2, BW * T * 2 * R * HB)

struct HeadWs {
    int B, T, N, P, R, save, SL, SLM, SLZ;
    size_t total;
#define X(name, esz, count) size_t off_##name, bytes_##name;
    GRL_HEAD_BUFFERS(X)
#undef X
    uint8_t* base;
    template <class Tp> Tp* ptr(size_t off) const { return reinterpret_cast<Tp*>(base + off); }
};

inline HeadWs head_ws_layout(int B, int T, int save) {
    HeadWs w;
    memset(&w, 0, sizeof(w));
    w.B = B; w.T = T; w.save = save;
    const int N = B * T, R = B * HS;
    const size_t P = (size_t)N * HS;
    w.N = N; w.P = (int)P; w.R = R;
    const int SL = save ? T : 1, SLM = save ? T + 1 : 2, SLZ = save ? T : 2;
    w.SL = SL; w.SLM = SLM; w.SLZ = SLZ;
    const size_t BW = save ? 1 : 0;
    const size_t p1 = (size_t)4 * N * HG, p2 = (size_t)16 * B * HG;
    const size_t PART = p1 > p2 ? p1 : p2;
    size_t off = 0;
#define X(name, esz, count)                                  \
    w.off_##name = off;                                      \
    w.bytes_##name = (size_t)(esz) * (size_t)(count);        \
    off += align_up(w.bytes_##name, 1024);
    GRL_HEAD_BUFFERS(X)
#undef X
    w.total = off;
    return w;
}

#define WS_F32(w, name) ((w).ptr<float>((w).off_##name))
#define WS_BF(w, name) ((w).ptr<__nv_bfloat16>((w).off_##name))

// out[rows][N] = in[rows][K] . W (+ bias) for the B-row global-descriptor branch, on the tcgen05 GEMM (rows << 128: the TMA
// box zero-fills the missing rows).  w_mn = 0: W planes stored [N][K] (nn.Linear forward); 1: stored [K][N] (its input gradient).
int small_gemm(grl_handle* h, cudaStream_t st, const HeadWs& w, const float* in, int rows, int K, const __nv_bfloat16* w_hi,
               const __nv_bfloat16* w_lo, long long ldw, int w_mn, const float* bias, float* out, int N);

// ------------------------------------------------------------------ device helpers
// Standard elementwise tile: 256 threads cover 128 rows x 64 channels; thread -> channel group
// cg = t & 7 (8 channels = 32 B fp32 / 16 B bf16), row slot r0 = t >> 3, rows r0 + 32*k (k < 4).
// A warp touches 4 rows x 256 B (fp32): fully coalesced.
struct Tile {
    int cg, r0;
    __device__ Tile() : cg(threadIdx.x & 7), r0(threadIdx.x >> 3) {}
};

__device__ __forceinline__ void load8(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
// planes -> fp32 (hi + lo)
__device__ __forceinline__ void load8_planes(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float* v) {
    const uint4 h = *reinterpret_cast<const uint4*>(hi), l = *reinterpret_cast<const uint4*>(lo);
    v[0] = bf16_lo_f(h.x) + bf16_lo_f(l.x); v[1] = bf16_hi_f(h.x) + bf16_hi_f(l.x);
    v[2] = bf16_lo_f(h.y) + bf16_lo_f(l.y); v[3] = bf16_hi_f(h.y) + bf16_hi_f(l.y);
    v[4] = bf16_lo_f(h.z) + bf16_lo_f(l.z); v[5] = bf16_hi_f(h.z) + bf16_hi_f(l.z);
    v[6] = bf16_lo_f(h.w) + bf16_lo_f(l.w); v[7] = bf16_hi_f(h.w) + bf16_hi_f(l.w);
}
// only the hi plane (sign / >0 tests)
__device__ __forceinline__ void load8_hi(const __nv_bfloat16* hi, float* v) {
    const uint4 h = *reinterpret_cast<const uint4*>(hi);
    v[0] = bf16_lo_f(h.x); v[1] = bf16_hi_f(h.x); v[2] = bf16_lo_f(h.y); v[3] = bf16_hi_f(h.y);
    v[4] = bf16_lo_f(h.z); v[5] = bf16_hi_f(h.z); v[6] = bf16_lo_f(h.w); v[7] = bf16_hi_f(h.w);
}
__device__ __forceinline__ void store8_planes(__nv_bfloat16* hi, __nv_bfloat16* lo, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(v[2 * j], h0, l0); split_bf16(v[2 * j + 1], h1, l1);
        h[j] = pack_bf16(h0, h1); l[j] = pack_bf16(l0, l1);
    }
    *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

// Sum acc[8] over the 32 row slots of the tile and write 64 column sums to out[0..63].
// `red` is a [32][65] float scratch in shared memory.  Contains __syncthreads().
__device__ __forceinline__ void tile_colsum(const float* acc, float* red, float* out, const Tile& t) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[t.r0 * 65 + t.cg * 8 + i] = acc[i];
    __syncthreads();
    if (threadIdx.x < 64) {
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) s += red[r * 65 + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// one fp16 plane (single-pass operands of the f1 / f2 gradient GEMMs); saturating, so a diverged activation cannot become inf
__device__ __forceinline__ void store8_f16(__half* p, const float* v, float scale = 1.f) {
    __half2 h2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        h2[j] = __floats2half2_rn(fminf(fmaxf(v[2 * j] * scale, -65504.f), 65504.f), fminf(fmaxf(v[2 * j + 1] * scale, -65504.f), 65504.f));
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(h2);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

}  // namespace grl
