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
#include "check.cuh"
#include "ozaki.cuh"
#include "../../include/cuda-helper.h"
#include <cstring>
#include <nccl.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>

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
void *svdgpu_stream_create_priority(int high)
{
    int lo = 0, hi = 0;
    cudaStream_t s;
    SVD_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));      /* lo = least, hi = greatest priority */
    SVD_CUDA_CHECK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high ? hi : lo));
    return (void *)s;
}
int svdgpu_host_register(void *p, size_t bytes)
{
    /* 0 = registered here (unregister later), 1 = already page-locked / not registrable: use it as it is */
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) return 1;
    (void)cudaGetLastError();
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { (void)cudaGetLastError(); return 1; }
    return 0;
}
void svdgpu_host_unregister(void *p) { if (cudaHostUnregister(p) != cudaSuccess) (void)cudaGetLastError(); }
void svdgpu_range_push(const char *name) { nvtxRangePushA(name); }
void svdgpu_range_pop(void) { nvtxRangePop(); }
void svdgpu_device_sync(void) { SVD_CUDA_CHECK(cudaDeviceSynchronize()); }
int svdgpu_enable_peer_access(int dev, int peer)
{
    int can = 0, cur = 0;
    SVD_CUDA_CHECK(cudaGetDevice(&cur));
    SVD_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, dev, peer));
    if (!can) return 0;
    SVD_CUDA_CHECK(cudaSetDevice(dev));
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); SVD_CUDA_CHECK(cudaSetDevice(cur)); return 0; }
    (void)cudaGetLastError();
    SVD_CUDA_CHECK(cudaSetDevice(cur));
    return 1;
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
static void hook_thunk(void *user, int done, cudaStream_t st)
{
    const svdgpu_progress *p = (const svdgpu_progress *)user;
    p->fn(p->user, done, (void *)st);
}
void svdgpu_bidiag_progress(int m, int n, double *dA, long lda, double *dalpha, double *dbeta, void *dwork, int nb,
                            const svdgpu_progress *progress, void *stream)
{
    ProgressHook h = {hook_thunk, (void *)progress, progress ? progress->every : 0};
    bidiag_device(m, n, dA, lda, dalpha, dbeta, dwork, nb, S(stream), (progress && progress->fn) ? &h : nullptr);
}
void svdgpu_qr_progress(int m, int n, double *dA, long lda, double *dR, long ldr, void *dwork,
                        const svdgpu_progress *progress, void *stream)
{
    if (m < n) { fprintf(stderr, "svdgpu_qr: needs m >= n (m=%d n=%d)\n", m, n); abort(); }
    ProgressHook h = {hook_thunk, (void *)progress, progress ? progress->every : 0};
    qr_device(m, n, dA, lda, dR, ldr, dwork, S(stream), (progress && progress->fn) ? &h : nullptr);
}
int svdgpu_wy_panel_width(void) { return wy_panel_width(); }
int svdgpu_wy_panel_count(int nref) { return wy_panel_count(nref); }
size_t svdgpu_wy_panels_bytes(int rows, int nref) { return wy_panels_bytes(rows, nref); }
size_t svdgpu_wy_apply_workspace(int nc) { return wy_apply_workspace_bytes(nc); }
void svdgpu_wy_setup(int left, int rows, int nref, const double *dA_mod, long lda, void *dpanels, int pb, int pe,
                     void *stream)
{
    wy_setup_device(left, rows, nref, dA_mod, lda, dpanels, pb, pe, S(stream));
}
void svdgpu_wy_apply_prepared(int left, int rows, int nref, const void *dpanels, double *dC, long ldc, int nc,
                              void *dwork, void *stream)
{
    wy_apply_prepared(left, rows, nref, dpanels, dC, ldc, nc, dwork, S(stream));
}
void svdgpu_wy_panel_slices(void *dpanels, int rows, int nref, int pb, int pe, double **dV, double **dVT,
                            size_t *count)
{
    wy_panel_slices(dpanels, rows, nref, pb, pe, dV, dVT, count);
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
size_t svdgpu_ozaki_workspace(int M, int N) { return ozaki_workspace_bytes(M, N); }
void svdgpu_ozaki_update(int M, int N, double sign, const double *dA, long lda, const double *dB, long ldb, double *dC,
                         long ldc, void *dwork, void *stream)
{
    ozaki_update_device(M, N, sign, dA, lda, dB, ldb, dC, ldc, dwork, S(stream));
}
size_t svdgpu_check_workspace(int m, int n, int nc) { return check_workspace_bytes(m, n, nc); }
void svdgpu_check(int m, int n, const double *dA0, long lda, const double *dsigma, const double *dU, long ldu,
                  const double *dV, long ldv, int nc, double *dout6, void *dwork, void *stream)
{
    check_device(m, n, dA0, lda, dsigma, dU, ldu, dV, ldv, nc, dout6, dwork, S(stream));
}
void svdgpu_bidiag_pass_probe(int m, int n, const double *dA, long lda, void *dwork, int which,
                              void *stream)
{
    bidiag_pass_probe(m, n, dA, lda, dwork, which, S(stream));
}

// ---- NCCL wrappers ("NCCL wrappers (broadcast, all-gather)") -----------------------------
// libnccl.so.2 is opened on first use: a process that never shards never loads it, and a process that has
// already loaded a copy (PyTorch bundles one under the same soname) shares that copy.
namespace {
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*CommCount)(const ncclComm_t, int *);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int *);
};
NcclApi *nccl_api()
{
    static NcclApi api;
    static bool loaded = false;
    if (loaded) return &api;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "*** libsvdgpu: multi-GPU run requested but libnccl.so.2 cannot be loaded (%s)\n", dlerror()); abort(); }
#define SVD_NCCL_SYM(f) do { *(void **)(&api.f) = dlsym(h, "nccl" #f); \
        if (!api.f) { fprintf(stderr, "*** libsvdgpu: nccl" #f " missing from libnccl\n"); abort(); } } while (0)
    SVD_NCCL_SYM(GetUniqueId); SVD_NCCL_SYM(CommInitRank); SVD_NCCL_SYM(CommInitAll); SVD_NCCL_SYM(CommDestroy);
    SVD_NCCL_SYM(CommCount); SVD_NCCL_SYM(GroupStart); SVD_NCCL_SYM(GroupEnd); SVD_NCCL_SYM(Broadcast);
    SVD_NCCL_SYM(AllGather); SVD_NCCL_SYM(GetErrorString); SVD_NCCL_SYM(GetVersion);
#undef SVD_NCCL_SYM
    loaded = true;
    return &api;
}
}
#define SVD_NCCL_CHECK(expr)                                                                         \
    do {                                                                                             \
        ncclResult_t r__ = (expr);                                                                   \
        if (r__ != ncclSuccess) {                                                                    \
            fprintf(stderr, "*** '%s' in '%s' on line %d failed with NCCL error '%s'.\n", #expr,     \
                    __FILE__, __LINE__, nccl_api()->GetErrorString(r__));                            \
            abort();                                                                                 \
        }                                                                                            \
    } while (0)

int svdgpu_nccl_version(void) { int v = 0; SVD_NCCL_CHECK(nccl_api()->GetVersion(&v)); return v; }
void svdgpu_nccl_unique_id(void *id128)
{
    static_assert(sizeof(ncclUniqueId) == SVDGPU_NCCL_ID_BYTES, "ncclUniqueId size");
    SVD_NCCL_CHECK(nccl_api()->GetUniqueId((ncclUniqueId *)id128));
}
void *svdgpu_nccl_comm_init_rank(int nranks, int rank, const void *id128)
{
    ncclComm_t c;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    SVD_NCCL_CHECK(nccl_api()->CommInitRank(&c, nranks, id, rank));
    return (void *)c;
}
void svdgpu_nccl_comm_init_all(int ndev, const int *devices, void **comms_out)
{
    ncclComm_t c[64];
    if (ndev > 64) { fprintf(stderr, "svdgpu_nccl_comm_init_all: too many devices\n"); abort(); }
    SVD_NCCL_CHECK(nccl_api()->CommInitAll(c, ndev, devices));
    for (int i = 0; i < ndev; ++i) comms_out[i] = (void *)c[i];
}
void svdgpu_nccl_comm_destroy(void *comm) { if (comm) SVD_NCCL_CHECK(nccl_api()->CommDestroy((ncclComm_t)comm)); }
int svdgpu_nccl_comm_count(void *comm) { int n = 0; SVD_NCCL_CHECK(nccl_api()->CommCount((ncclComm_t)comm, &n)); return n; }
void svdgpu_nccl_group_start(void) { SVD_NCCL_CHECK(nccl_api()->GroupStart()); }
void svdgpu_nccl_group_end(void) { SVD_NCCL_CHECK(nccl_api()->GroupEnd()); }
void svdgpu_nccl_bcast(void *comm, void *dbuf, size_t count, int root, void *stream)
{
    SVD_NCCL_CHECK(nccl_api()->Broadcast(dbuf, dbuf, count, ncclDouble, root, (ncclComm_t)comm, S(stream)));
}
void svdgpu_nccl_allgather(void *comm, const void *dsend, void *drecv, size_t count_per_rank, void *stream)
{
    SVD_NCCL_CHECK(nccl_api()->AllGather(dsend, drecv, count_per_rank, ncclDouble, (ncclComm_t)comm, S(stream)));
}

} // extern "C"
