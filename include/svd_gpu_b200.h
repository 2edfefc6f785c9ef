/* svd_gpu_b200.h — extensions around the drop-in svd_gpu(): device-resident entry points,
 * per-phase timings and the multi-GPU (column-sharded) vector phases.  Plain C ABI. */
#ifndef SVDGPU_B200_H
#define SVDGPU_B200_H
#include <stddef.h>
#include "svd_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Whole path on device-resident data (what svd_gpu() runs between its copies).
 * dA m x n (lda even >= m; rows [m,lda) finite) is overwritten with the reflectors;
 * dsigma[min(m,n)] ascending; dU (m x min, ldu) and dV (n x min, ldv) may both be NULL
 * (values only).  Enqueues on `stream` (cudaStream_t as void*), synchronises once inside
 * the dDC phase; per-phase events are kept for svd_gpu_last_phase_ms(). */
void svd_gpu_dev(int m, int n, double *dA, long lda, double *dsigma, double *dU, long ldu,
                 double *dV, long ldv, void *stream);

/* Column-sharded vector phases (north_star: twisted solves + back-transform shard by
 * singular-value blocks): given the bidiagonalization output (dA_mod, dalpha, dbeta with
 * dbeta zero-padded to min(m,n) entries) and ALL singular values, compute columns
 * [i0, i0+ns) of U and V into dUblk (m x ns, ldu) / dVblk (n x ns, ldv) and the polished
 * singular values dsig_out[ns] (may be NULL). */
void svd_gpu_vectors_dev(int m, int n, const double *dA_mod, long lda, const double *dalpha,
                         const double *dbeta, const double *dsigma_all, int i0, int ns,
                         double *dUblk, long ldu, double *dVblk, long ldv, double *dsig_out,
                         void *stream);
/* Bidiagonalization + dDC on device data: the part that runs on one GPU.
 * dalpha[min], dbeta[min] (zero padded), dsigma[min]. */
void svd_gpu_values_dev(int m, int n, double *dA, long lda, double *dalpha, double *dbeta,
                        double *dsigma, void *stream);

/* milliseconds of the last svd_gpu()/svd_gpu_dev() call on this thread's context:
 * [0] h2d  [1] bidiag  [2] dDC  [3] twisted  [4] back-transform  [5] d2h  [6] total
 * (device times from CUDA events on the call's stream; [0],[5] are 0 for svd_gpu_dev). */
void svd_gpu_last_phase_ms(float ms[7]);

/* tunables (also read once from the environment: SVD_GPU_NB, SVD_GPU_RQI, SVD_GPU_DEVICE) */
void svd_gpu_set_option(const char *name, int value);

#ifdef __cplusplus
}
#endif
#endif
