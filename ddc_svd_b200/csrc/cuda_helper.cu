// cuda_helper.cu — the thin C-ABI CUDA layer of libsvdgpu.so (include/cuda-helper.h).
// Takes the place of the reference's OpenCL glue cl-helper.c: context/queue creation
// (create_context_on, cl-helper.c:209-366), buffer management, blocking copies
// (bidiag_par.c:298-301, :405-418) and per-kernel argument marshalling
// (SET_n_KERNEL_ARGS + clEnqueueNDRangeKernel).  No interactive prompt, no run-time
// compilation of kernel source files: the kernels are compiled for sm_100a at build time.
#include "common.cuh"
#include "bidiag.cuh"
#include "ddc.cuh"
#include "twisted.cuh"
#include "backtransform.cuh"
#include "../../include/cuda-helper.h"

namespace svdgpu {
void scale_matrix_device(int m, int n, double *A, long lda, double *sc, double *work, cudaStream_t st);
void scale_vector_device(int n, double *x, const double *factor, cudaStream_t st);
}
using namespace svdgpu;
unsigned long long g_svdgpu_launches = 0;
static inline cudaStream_t S(void *s) { return (cudaStream_t)s; }

extern "C" {

int svdgpu_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        fprintf(stderr, "*** libsvdgpu: no usable CUDA device (%s). There is no CPU fallback.\n",
                cudaGetErrorString(e));
        abort();
    }
    return n;
}
void svdgpu_set_device(int dev) { SVD_CUDA_CHECK(cudaSetDevice(dev)); }
int svdgpu_get_device(void) { int d = 0; SVD_CUDA_CHECK(cudaGetDevice(&d)); return d; }
const char *svdgpu_device_name(void)
{
    static char name[320];
    cudaDeviceProp p;
    SVD_CUDA_CHECK(cudaGetDeviceProperties(&p, svdgpu_get_device()));
    snprintf(name, sizeof name, "%s (sm_%d%d, %d SMs)", p.name, p.major, p.minor, p.multiProcessorCount);
    return name;
}
void *svdgpu_malloc(size_t bytes)
{
    void *p = nullptr;
    SVD_CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 1));
    return p;
}
void svdgpu_free(void *dptr) { if (dptr) SVD_CUDA_CHECK(cudaFree(dptr)); }
void svdgpu_memset(void *dptr, int value, size_t bytes, void *stream)
{
    SVD_CUDA_CHECK(cudaMemsetAsync(dptr, value, bytes, S(stream)));
}
void svdgpu_h2d(void *dst, const void *src, size_t bytes, void *stream)
{
    SVD_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, S(stream)));
}
void svdgpu_d2h(void *dst, const void *src, size_t bytes, void *stream)
{
    SVD_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, S(stream)));
}
void svdgpu_d2d(void *dst, const void *src, size_t bytes, void *stream)
{
    SVD_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, S(stream)));
}
void svdgpu_h2d_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes,
                   size_t height, void *stream)
{
    SVD_CUDA_CHECK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height,
                                     cudaMemcpyHostToDevice, S(stream)));
}
void svdgpu_d2h_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes,
                   size_t height, void *stream)
{
    SVD_CUDA_CHECK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height,
                                     cudaMemcpyDeviceToHost, S(stream)));
}
void *svdgpu_stream_create(void)
{
    cudaStream_t s;
    SVD_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return (void *)s;
}
void svdgpu_stream_destroy(void *stream) { SVD_CUDA_CHECK(cudaStreamDestroy(S(stream))); }
void svdgpu_stream_sync(void *stream) { SVD_CUDA_CHECK(cudaStreamSynchronize(S(stream))); }
void svdgpu_stream_wait_event(void *stream, void *event)
{
    SVD_CUDA_CHECK(cudaStreamWaitEvent(S(stream), (cudaEvent_t)event, 0));
}
void *svdgpu_event_create(void)
{
    cudaEvent_t e;
    SVD_CUDA_CHECK(cudaEventCreate(&e));
    return (void *)e;
}
void svdgpu_event_destroy(void *event) { SVD_CUDA_CHECK(cudaEventDestroy((cudaEvent_t)event)); }
void svdgpu_event_record(void *event, void *stream)
{
    SVD_CUDA_CHECK(cudaEventRecord((cudaEvent_t)event, S(stream)));
}
float svdgpu_event_elapsed_ms(void *start, void *stop)
{
    float ms = 0.f;
    SVD_CUDA_CHECK(cudaEventSynchronize((cudaEvent_t)stop));
    SVD_CUDA_CHECK(cudaEventElapsedTime(&ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return ms;
}
void *svdgpu_host_alloc(size_t bytes)
{
    void *p = nullptr;
    SVD_CUDA_CHECK(cudaMallocHost(&p, bytes ? bytes : 1));
    return p;
}
void svdgpu_host_free(void *p) { if (p) SVD_CUDA_CHECK(cudaFreeHost(p)); }

unsigned long long svdgpu_launch_count(void) { return g_svdgpu_launches; }

// ---- kernel families -------------------------------------------------------------------
size_t svdgpu_bidiag_workspace(int m, int n, long lda) { return bidiag_workspace_bytes(m, n, lda); }
int svdgpu_bidiag_tail_start(int m, int n, int nb, int ctas) { return bidiag_tail_start(m, n, nb, ctas); }
void svdgpu_bidiag(int m, int n, double *dA, long lda, double *dalpha, double *dbeta, void *dwork, int nb,
                   void *stream)
{
    bidiag_device(m, n, dA, lda, dalpha, dbeta, dwork, nb, S(stream));
}
size_t svdgpu_ddc_workspace(int N) { return ddc_workspace_bytes(N); }
void svdgpu_ddc_values(int N, const double *db1, const double *db2, double *dsigma, void *dwork,
                       void *stream)
{
    ddc_values_device(N, db1, db2, dsigma, dwork, S(stream));
}
size_t svdgpu_twisted_workspace(int n, int mb, int ns) { return twisted_workspace_bytes(n, mb, ns); }
void svdgpu_twisted_vectors(int n, int mb, const double *da, const double *db, const double *dsigma_all,
                            int ntot, int i0, int ns, double *dX, long ldx, double *dY, long ldy,
                            double *dsigma_out, int rqi_steps, void *dwork, void *stream)
{
    twisted_vectors_device(n, mb, da, db, dsigma_all, ntot, i0, ns, dX, ldx, dY, ldy, dsigma_out,
                           rqi_steps, dwork, S(stream));
}
size_t svdgpu_backtransform_workspace(int rows, int nref, int nc)
{
    return backtransform_workspace_bytes(rows, nref, nc);
}
size_t svdgpu_qr_workspace(int m, int n) { return qr_workspace_bytes(m, n); }
void svdgpu_qr(int m, int n, double *dA, long lda, double *dR, long ldr, void *dwork, void *stream)
{
    if (m < n) { fprintf(stderr, "svdgpu_qr: needs m >= n (m=%d n=%d)\n", m, n); abort(); }
    qr_device(m, n, dA, lda, dR, ldr, dwork, S(stream));
}
void svdgpu_wy_apply(int left, int rows, int nref, const double *dA_mod, long lda, double *dC, long ldc,
                     int nc, void *dwork, void *stream)
{
    wy_apply_device(left, rows, nref, dA_mod, lda, dC, ldc, nc, dwork, S(stream));
}
void svdgpu_dgemm(int transA, int transB, int M, int N, int K, double alpha, const double *dA, long lda,
                  const double *dB, long ldb, double beta, double *dC, long ldc, void *stream)
{
    GemmArgs g = {};
    g.M = M; g.N = N; g.K = K;
    g.A = dA; g.lda = lda; g.transA = transA;
    g.B = dB; g.ldb = ldb; g.transB = transB;
    g.C = dC; g.ldc = ldc; g.alpha = alpha; g.beta = beta; g.batch = 1; g.splitk = 1;
    dgemm_dmma(g, S(stream));
}
void svdgpu_scale_matrix(int m, int n, double *dA, long lda, double *dscale, double *dwork, void *stream)
{
    scale_matrix_device(m, n, dA, lda, dscale, dwork, S(stream));
}
void svdgpu_transpose(int m, int n, const double *dA, long lda, double *dAt, long ldat, void *stream)
{
    transpose_device(m, n, dA, lda, dAt, ldat, S(stream));
}
void svdgpu_scale_vector(int n, double *dx, const double *dfactor, void *stream)
{
    scale_vector_device(n, dx, dfactor, S(stream));
}
void svdgpu_bidiag_pass_probe(int m, int n, const double *dA, long lda, void *dwork, int which,
                              void *stream)
{
    bidiag_pass_probe(m, n, dA, lda, dwork, which, S(stream));
}

} // extern "C"
