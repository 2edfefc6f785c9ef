/* cuda-helper.h — thin C-ABI device layer of libsvdgpu.so (plain pointers and sizes only).
 *
 * This is what replaces the reference's OpenCL glue cl-helper.c / cl-helper.h:
 *   create_context_on (cl-helper.h:99-101)        -> svdgpu_set_device / svdgpu_stream_create
 *   clCreateBuffer / clReleaseMemObject            -> svdgpu_malloc / svdgpu_free
 *   clEnqueueWriteBuffer / clEnqueueReadBuffer     -> svdgpu_h2d / svdgpu_d2h (+ _2d)
 *   clFinish                                       -> svdgpu_stream_sync
 *   CHECK_CL_ERROR / CALL_CL_GUARDED (:47-69)      -> every entry point aborts on CUDA errors
 *   read_file + kernel_from_string + SET_n_KERNEL_ARGS + clEnqueueNDRangeKernel
 *                                                  -> one launcher per kernel family below
 * All `stream` arguments are cudaStream_t passed as void* (NULL = the legacy default stream);
 * all d* pointers are device pointers.  Launchers only enqueue work unless stated.
 */
#ifndef SVDGPU_CUDA_HELPER_H
#define SVDGPU_CUDA_HELPER_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- device / memory / stream glue -------------------------------------------------- */
int    svdgpu_device_count(void);
void   svdgpu_set_device(int dev);
int    svdgpu_get_device(void);
const char *svdgpu_device_name(void);           /* static buffer, current device */
void  *svdgpu_malloc(size_t bytes);             /* stream-ordered pool allocation, aborts on failure */
void   svdgpu_free(void *dptr);
void   svdgpu_memset(void *dptr, int value, size_t bytes, void *stream);
void   svdgpu_h2d(void *dst, const void *src, size_t bytes, void *stream);
void   svdgpu_d2h(void *dst, const void *src, size_t bytes, void *stream);
void   svdgpu_d2d(void *dst, const void *src, size_t bytes, void *stream);
void   svdgpu_h2d_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes,
                     size_t height, void *stream);
void   svdgpu_d2h_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes,
                     size_t height, void *stream);
void  *svdgpu_stream_create(void);
void  *svdgpu_stream_create_priority(int high);  /* high = 1: greatest priority (the factorization), 0: least (side work) */
void   svdgpu_stream_destroy(void *stream);
void   svdgpu_stream_sync(void *stream);
void   svdgpu_stream_wait_event(void *stream, void *event);
void  *svdgpu_event_create(void);
void   svdgpu_event_destroy(void *event);
void   svdgpu_event_record(void *event, void *stream);
float  svdgpu_event_elapsed_ms(void *start, void *stop);   /* synchronises on stop */
/* number of kernels this library has launched so far in this process */
unsigned long long svdgpu_launch_count(void);
void  *svdgpu_host_alloc(size_t bytes);         /* pinned host memory */
void   svdgpu_host_free(void *p);
/* page-lock a caller's malloc'd buffer for the duration of a call (the reference's callers hand pageable
 * memory, test-whole-svd.c:44-66).  Returns 0 if it was registered here (call svdgpu_host_unregister
 * afterwards), 1 if it already is page-locked or cannot be registered (use it as it is). */
int    svdgpu_host_register(void *p, size_t bytes);
void   svdgpu_host_unregister(void *p);
void   svdgpu_device_sync(void);
int    svdgpu_enable_peer_access(int dev, int peer);   /* 1 = dev can now map peer's memory */
/* NVTX ranges around the phases (visible in nsys / ncu --nvtx) */
void   svdgpu_range_push(const char *name);
void   svdgpu_range_pop(void);

/* ---- NCCL wrappers (multi-GPU: broadcast of the prepared reflector panels and of the bidiagonal,
 * all-gather of the singular values / of U, V blocks).  libnccl.so.2 is opened on first use. -------- */
#define SVDGPU_NCCL_ID_BYTES 128
int    svdgpu_nccl_version(void);
void   svdgpu_nccl_unique_id(void *id128);                                   /* ncclGetUniqueId */
void  *svdgpu_nccl_comm_init_rank(int nranks, int rank, const void *id128);  /* one process per GPU, current device */
void   svdgpu_nccl_comm_init_all(int ndev, const int *devices, void **comms_out);   /* one process, ndev GPUs */
void   svdgpu_nccl_comm_destroy(void *comm);
int    svdgpu_nccl_comm_count(void *comm);
void   svdgpu_nccl_group_start(void);
void   svdgpu_nccl_group_end(void);
void   svdgpu_nccl_bcast(void *comm, void *dbuf, size_t count_doubles, int root, void *stream);   /* in place */
void   svdgpu_nccl_allgather(void *comm, const void *dsend, void *drecv, size_t count_doubles_per_rank,
                             void *stream);

/* ---- kernel families (one launcher each) -------------------------------------------- */
/* Householder bidiagonalization (bidiag_par.c:34-450 + its 10 .cl kernels).
 * dA: m x n, lda even and >= m (rows [m,lda) finite); dalpha[min(m,n)], dbeta[n-1 | m]. */
size_t svdgpu_bidiag_workspace(int m, int n, long lda);
void   svdgpu_bidiag(int m, int n, double *dA, long lda, double *dalpha, double *dbeta,
                     void *dwork, int nb, void *stream);
/* the same with a progress callback: fn(user, done, stream) is called on the host, at enqueue time, every
 * `every` steps and once at the end, when the work enqueued so far leaves reflectors [0, done) final in dA
 * (svd_gpu.c uses it to prepare and broadcast compact-WY panels while the factorization still runs) */
typedef struct { void (*fn)(void *user, int done, void *stream); void *user; int every; } svdgpu_progress;
void   svdgpu_bidiag_progress(int m, int n, double *dA, long lda, double *dalpha, double *dbeta,
                              void *dwork, int nb, const svdgpu_progress *progress, void *stream);
/* planning only (no device): the first step that the on-chip tail kernel takes over when `ctas`
 * CTAs are co-resident (148 on a B200), min(m,n) if the trailing block never fits */
int    svdgpu_bidiag_tail_start(int m, int n, int nb, int ctas);
/* dDC singular values (Calculations-Parallel.c:852-874); b1,b2 of length N; only enqueues. */
size_t svdgpu_ddc_workspace(int N);
void   svdgpu_ddc_values(int N, const double *db1, const double *db2, double *dsigma,
                         void *dwork, void *stream);
/* twisted-factorization vectors for sigma_all[i0 .. i0+ns) (parallel-twisted.c:554-637,
 * :530-551); X[t*ldx + j], j < mb; Y[t*ldy + j], j < n (Y may be NULL). */
size_t svdgpu_twisted_workspace(int n, int mb, int ns);
void   svdgpu_twisted_vectors(int n, int mb, const double *da, const double *db,
                              const double *dsigma_all, int ntot, int i0, int ns,
                              double *dX, long ldx, double *dY, long ldy, double *dsigma_out,
                              int rqi_steps, void *dwork, void *stream);
/* compact-WY back-transform (multU / multV, bidiag_par.c:990-1095):
 * C(rows x nc) <- H_0 ... H_{nref-1} C, reflectors read from dA_mod. */
size_t svdgpu_backtransform_workspace(int rows, int nref, int nc);
void   svdgpu_wy_apply(int left, int rows, int nref, const double *dA_mod, long lda, double *dC,
                       long ldc, int nc, void *dwork, void *stream);
/* the same in two halves: panel set-up (depends on the reflectors only; panels [pb, pe) of
 * svdgpu_wy_panel_width() reflectors need only those reflectors to be final) and application from the
 * prepared panels.  svdgpu_wy_panel_slices gives the two contiguous device ranges (V and V*T columns of
 * panels [pb, pe), `count` doubles each) that another device needs to run svdgpu_wy_apply_prepared. */
int    svdgpu_wy_panel_width(void);
int    svdgpu_wy_panel_count(int nref);
size_t svdgpu_wy_panels_bytes(int rows, int nref);
size_t svdgpu_wy_apply_workspace(int nc);
void   svdgpu_wy_setup(int left, int rows, int nref, const double *dA_mod, long lda, void *dpanels,
                       int pb, int pe, void *stream);
void   svdgpu_wy_apply_prepared(int left, int rows, int nref, const void *dpanels, double *dC, long ldc,
                                int nc, void *dwork, void *stream);
void   svdgpu_wy_panel_slices(void *dpanels, int rows, int nref, int pb, int pe, double **dV, double **dVT,
                              size_t *count);
/* Householder QR of a tall matrix (m >= n) in the same reflector convention, used by svd_gpu() for
 * m >> n (LAPACK dgesdd's "QR first" route; the reference has no counterpart and bidiagonalizes the
 * full m x n matrix, bidiag_par.c:310-397).  dA <- reflectors (diagonal and below) + strict upper
 * triangle of R; dR (n x n, ldr) <- R.  svdgpu_wy_apply(1, m, n, dA, ...) then applies Q. */
size_t svdgpu_qr_workspace(int m, int n);
void   svdgpu_qr(int m, int n, double *dA, long lda, double *dR, long ldr, void *dwork, void *stream);
void   svdgpu_qr_progress(int m, int n, double *dA, long lda, double *dR, long ldr, void *dwork,
                          const svdgpu_progress *progress, void *stream);
/* FP64 DMMA GEMM building block: C = beta*C + alpha*op(A)*op(B), column-major */
void   svdgpu_dgemm(int transA, int transB, int M, int N, int K, double alpha, const double *dA,
                    long lda, const double *dB, long ldb, double beta, double *dC, long ldc,
                    void *stream);
/* The same update C += sign * A (M x 128) * B (128 x N) on the 5th-generation tensor cores: tcgen05 has no FP64
 * kind, so the operands are cut into 8 signed 7-bit slices (error-free), the 36 slice products with i + j <= 7
 * run as tcgen05.mma.kind::i8 into int32 accumulators in tensor memory and are recombined in FP64 (Ozaki
 * scheme).  K is the compact-WY panel width (128).  dwork: svdgpu_ozaki_workspace(M, N) bytes. */
size_t svdgpu_ozaki_workspace(int M, int N);
void   svdgpu_ozaki_update(int M, int N, double sign, const double *dA, long lda, const double *dB, long ldb,
                           double *dC, long ldc, void *dwork, void *stream);
/* power-of-two range guard (no counterpart in the reference, which overflows/underflows on inputs
 * near 1e+-150): dscale[0] <- factor applied to dA in place (1.0 unless max|A| is outside
 * [1e-100, 1e100]), dscale[1] <- its inverse; dwork needs 1024 doubles.  svdgpu_scale_vector
 * multiplies dx[0..n) by *dfactor (used to scale sigma back). */
void   svdgpu_scale_matrix(int m, int n, double *dA, long lda, double *dscale, double *dwork, void *stream);
void   svdgpu_scale_vector(int n, double *dx, const double *dfactor, void *stream);
/* dAt (n x m, ldat) = dA^T (m x n, lda), on the device: replaces the host transpose of
 * matrix_helper.c:166-174 (svd_gpu.c:104) and turns a wide problem into a tall one */
void   svdgpu_transpose(int m, int n, const double *dA, long lda, double *dAt, long ldat, void *stream);
/* residual / orthogonality of a computed SVD on the device: the enabled form of the check the reference's
 * driver carries switched off (test-whole-svd.c:81-96).  dA0 is the ORIGINAL matrix; nc = min(m,n) columns of
 * the factors for a whole SVD, fewer for a column block.  dout6 (device): ||U^T U - I||_F, ||V^T V - I||_F,
 * ||A - U S V^T||_F/||A||_F (block: ||A V - U S||_F/||A||_F), |sum sigma^2 - ||A||_F^2|/||A||_F^2, ||A||_F,
 * 1.0 if sigma ascends. */
size_t svdgpu_check_workspace(int m, int n, int nc);
void   svdgpu_check(int m, int n, const double *dA0, long lda, const double *dsigma, const double *dU, long ldu,
                    const double *dV, long ldv, int nc, double *dout6, void *dwork, void *stream);
/* one streaming pass over the full m x n matrix, for roofline measurement: which = 0 gemvT, 1 gemvN
 * (the split passes), 2 the fused single-read pass of step 0 (writes a reflector into column 0 of dA:
 * hand it a scratch copy); returns nothing, only enqueues. */
void   svdgpu_bidiag_pass_probe(int m, int n, const double *dA, long lda, void *dwork, int which,
                                void *stream);

#ifdef __cplusplus
}
#endif
#endif
