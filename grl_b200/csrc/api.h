// grl_b200 — internal host-side declarations shared by the translation units.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/grl_b200.h"
#include "gemm.cuh"

typedef CUresult (*grl_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct grl_prof_rec { cudaEvent_t e0, e1; double flops; };
struct grl_prof;   // std::vector<grl_prof_rec>, owned by the handle (gemm.cu)

struct grl_handle {
    int device;
    int num_sms;
    long long launches;
    grl_encode_tiled_fn encode;
    int prof_on;
    grl_prof* prof;
    cudaStream_t side;          // low-priority internal stream: work that is off the recurrence's critical path
    cudaEvent_t* events;        // pool of timing-disabled events for fork/join between the caller's stream and `side`
    int n_events, overlap;
    void* func_attrs;           // per-device record of cudaFuncAttributeMaxDynamicSharedMemorySize settings (gemm.cu)
    // communicator of the gallery-sharded search (comm.cu): an ncclComm_t created by grl_comm_init (owned) or handed in by
    // grl_comm_attach (borrowed); NULL == single rank
    void* comm;
    int comm_world, comm_rank, comm_owned;
    // per-stage events of the last grl_sharded_topk call (search.cu), recorded when stage_prof is on
    int stage_prof, n_stage_ev;
    cudaEvent_t* stage_ev;
    char err[512];
};

namespace grl {

int set_error(grl_handle* h, int code, const char* fmt, ...);

#define GRL_CUDA(h, expr)                                                                              \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return grl::set_error((h), GRL_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                  __FILE__, __LINE__);                                                 \
    } while (0)

#define GRL_LAUNCH_CHECK(h)                                                                            \
    do {                                                                                               \
        (h)->launches++;                                                                               \
        cudaError_t _e = cudaGetLastError();                                                           \
        if (_e != cudaSuccess)                                                                         \
            return grl::set_error((h), GRL_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                                  __FILE__, __LINE__);                                                 \
    } while (0)

#define GRL_TRY(expr)            \
    do {                         \
        int _r = (expr);         \
        if (_r != GRL_OK) return _r; \
    } while (0)

// A GEMM operand given as bf16 hi/lo planes.
struct Operand {
    const __nv_bfloat16* hi;
    const __nv_bfloat16* lo;
    long long ld;        // leading dimension in elements
    long long bstride;   // batch stride in elements (ignored when batch == 1)
    int mn_major;        // 0: [rows][K], 1: [K][rows]
};

// D[z] = A[z] * B[z]^T with the fused epilogue `epi`; bn = 0 picks the N tile automatically.
int gemm_launch(grl_handle* h, cudaStream_t st, int M, int N, int K, int batch, const Operand& A, const Operand& B,
                GemmEpi epi, int bn);

// D = A * B^T with ONE fp16 plane per operand (K-major, one MMA per k-step): coarse pass of the retrieval search.
int gemm_launch_f16(grl_handle* h, cudaStream_t st, int M, int N, int K, const __half* A, long long lda, const __half* B,
                    long long ldb, GemmEpi epi, int bn);

// D[z] = A[z] * B[z]^T with ONE fp16 plane per operand, any majors, batches (single MMA per k-step).
int gemm_launch_x1(grl_handle* h, cudaStream_t st, int M, int N, int K, int batch, const __half* A, long long lda, long long a_bstride,
                   int a_mn, const __half* B, long long ldb, long long b_bstride, int b_mn, GemmEpi epi);

// The same contraction on 256 x 256 tiles (coarse_gemm.cuh) when the problem is large enough, else gemm_launch_f16.
int coarse_gemm_launch(grl_handle* h, cudaStream_t st, int M, int N, int K, const __half* A, long long lda, const __half* B,
                       long long ldb, GemmEpi epi);

// fp32 [rows][cols] (ld_src) -> bf16 hi/lo planes [rows][cols] (ld_dst); optional per-row scale.
int split_planes(grl_handle* h, cudaStream_t st, const float* src, long long ld_src, __nv_bfloat16* hi,
                 __nv_bfloat16* lo, long long ld_dst, long long rows, int cols);
// fp32 [rows][cols] -> transposed planes [cols][rows] (ld_dst = leading dim of the transposed planes)
int split_planes_transposed(grl_handle* h, cudaStream_t st, const float* src, long long ld_src, __nv_bfloat16* hi,
                            __nv_bfloat16* lo, long long ld_dst, int rows, int cols);

// Function attributes are per device: a process that drives several GPUs (nn.DataParallel, mars_train.py:80) must opt every
// kernel in on every device.  Raises `func`'s dynamic shared-memory limit to `bytes` on the handle's device (once per size).
int ensure_dyn_smem(grl_handle* h, const void* func, int bytes);

// Event k of the handle's pool (grown on demand).  Events are only ever recorded/waited by the thread driving the handle.
cudaEvent_t pool_event(grl_handle* h, int k);
// `waiter` waits until everything enqueued on `signaler` so far has finished (uses pool event k).
int stream_wait(grl_handle* h, cudaStream_t signaler, cudaStream_t waiter, int k);
// record pool event k on `s` / make `s` wait for pool event k
int ev_record(grl_handle* h, int k, cudaStream_t s);
int ev_wait(grl_handle* h, int k, cudaStream_t s);

// Drops the handle's communicator (destroys it when the library created it).
void comm_release(grl_handle* h);

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline GemmEpi epi_default() {
    GemmEpi e;
    memset(&e, 0, sizeof(e));
    e.alpha = 1.f;
    e.grp_rows = 1;
    return e;
}

}  // namespace grl
