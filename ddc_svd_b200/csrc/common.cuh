// common.cuh — shared device/host helpers for the sm_100a SVD kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

// Abort-on-error, the convention of the reference's CHECK_CL_ERROR / CALL_CL_GUARDED
// (cl-helper.h:47-87): svd_gpu() returns void, every failure prints and abort()s.
#define SVD_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            fprintf(stderr, "*** '%s' in '%s' on line %d failed with error '%s'.\n",      \
                    #expr, __FILE__, __LINE__, cudaGetErrorString(e__));                  \
            abort();                                                                      \
        }                                                                                 \
    } while (0)

// every kernel launch site is followed by SVD_KERNEL_CHECK(); it also counts the launch so that
// bench.py can report how many of OUR kernels ran inside a timed region (svdgpu_launch_count)
extern unsigned long long g_svdgpu_launches;
#define SVD_KERNEL_CHECK() do { ++g_svdgpu_launches; SVD_CUDA_CHECK(cudaGetLastError()); } while (0)

// true the first time it is called for the current device with this flag set (function attributes such as the
// dynamic shared-memory limit are per device; setting them on every launch costs host time on the launch path)
struct DeviceOnce { bool done[64] = {}; };
static inline bool first_on_device(DeviceOnce &f)
{
    int dev = 0;
    SVD_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return true;
    if (f.done[dev]) return false;
    f.done[dev] = true;
    return true;
}

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }
static inline long round_up(long a, long b) { return (a + b - 1) / b * b; }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_prod(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit load that does not allocate in L1 (the trailing matrix is read once per pass)
__device__ __forceinline__ double2 ldg_stream2(const double *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                 : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

namespace svdgpu {

// ---- FP64 GEMM on the DMMA pipe (dgemm_dmma.cu) -------------------------------------
// C(MxN, ldc) = beta*C + alpha * op(A)(MxK) * op(B)(KxN), column-major, batched over grid.z:
// batch z uses A + z*sA, B + z*sB, C + z*sC and K - z*dK (and M - z*dM) so one launch can
// walk trapezoidal Householder panels.  With splitk > 1 the K range of every batch entry
// is cut into splitk slices and slice s writes its partial product (alpha applied,
// beta ignored) to C + s*sSplit; the caller reduces.
struct GemmArgs {
    int M, N, K;
    const double *A; long lda; int transA;
    const double *B; long ldb; int transB;
    double *C; long ldc;
    double alpha, beta;
    int batch; long sA, sB, sC; int dK, dM;
    int splitk; long sSplit;
};
void dgemm_dmma(const GemmArgs &g, cudaStream_t st);
// persistent warp-specialised variant (dgemm_ws.cu); dgemm_dmma() routes eligible shapes to it
constexpr int DGEMM_WS_DEFAULT = 1;
bool dgemm_ws_enabled();
int dgemm_ws_mode();
bool dgemm_ws_eligible(const GemmArgs &g);
void dgemm_ws(const GemmArgs &g, cudaStream_t st);
void transpose_device(int m, int n, const double *A, long lda, double *At, long ldat, cudaStream_t st);
// out(MxN, ldo) = beta*out + alpha * sum_s part[s]  (deterministic split-K reduction)
void sum_partials(double *out, long ldo, const double *part, long ldp, long sSplit, int nsplit,
                  int M, int N, double alpha, double beta, cudaStream_t st);

} // namespace svdgpu
