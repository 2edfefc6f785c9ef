/* svd_gpu_b200.h — extensions around the drop-in svd_gpu(): device-resident entry points,
 * per-phase timings, the multi-GPU (singular-value-block sharded) path and the on-device checker.
 * Plain C ABI.
 *
 * Threading: every entry point may be called from any host thread.  State (streams, events, one cached
 * device arena) is kept per DEVICE behind a lock: calls that use different devices run concurrently,
 * calls on the same device are serialised.  svd_gpu_set_option("release", 0) frees the arenas.
 *
 * What svd_gpu() leaves in A (the reference promises the reflectors of bidiag_par, svd_gpu.c:100):
 *   - m >= n, m < 2.5 n (every shape the reference handles correctly, i.e. square): exactly that.
 *   - m >= 2.5 n ("QR first"): the reflectors of A = QR below/on the diagonal and R above it; the
 *     bidiagonalization ran on the n x n factor R.  svd_gpu_set_option("qr_first", 0) (or
 *     SVD_GPU_QR_FIRST=0) restores the reference layout.
 *   - m < n: the transpose of the tall problem's storage (the SVD of A^T is computed).
 *     svd_gpu_set_option("wide_transpose", 0) restores the reference's wide layout (bidiag.c:33-186).
 *   A_mod from the default routes must therefore not be fed to multU / multV / form_u_par for tall (>= 2.5:1)
 *   or wide inputs unless the two options are switched off (tests/test_gpu_parity.py pins both contracts).
 */
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
 * (values only).  Ordered after the work already enqueued on `stream` (cudaStream_t as void*), and
 * `stream` waits for the results; internally the phases run on the library's own streams.  Per-phase
 * events are kept for svd_gpu_last_phase_ms(). */
void svd_gpu_dev(int m, int n, double *dA, long lda, double *dsigma, double *dU, long ldu,
                 double *dV, long ldv, void *stream);

/* milliseconds of the last svd_gpu()/svd_gpu_dev() call:
 * [0] h2d  [1] bidiag (+QR)  [2] dDC  [3] twisted  [4] back-transform  [5] d2h  [6] total
 * (device times from CUDA events; [0],[5] are 0 for svd_gpu_dev). */
void svd_gpu_last_phase_ms(float ms[7]);

/* tunables (also read once from the environment: SVD_GPU_NB, SVD_GPU_RQI, SVD_GPU_DEVICE, SVD_GPU_QR_FIRST,
 * SVD_GPU_WIDE_TRANSPOSE, SVD_GPU_NGPUS, SVD_GPU_HOST_REGISTER, SVD_GPU_WY_OVERLAP):
 * "nb", "rqi", "qr_first", "qr_ratio10", "wide_transpose", "ngpus" (GPUs svd_gpu() uses), "host_register"
 * (page-lock the caller's malloc'd buffers for the call; off by default: measured slower than pageable copies), "wy_overlap" (prepare WY panels during the factorization
 * even on one GPU), "release". */
void svd_gpu_set_option(const char *name, int value);

/* ---- multi-GPU: a group of ranks, one per GPU ------------------------------------------------------
 * north_star / SURVEY.md 8e: the bidiagonalization and the dDC values run on rank 0; the twisted vector
 * solves and the back-transform shard by contiguous blocks of singular values (the reference's parallel
 * loop over vectors, svd_gpu.c:117-121).  Rank r owns singular values [i0, i0 + ns) given by
 * svdgpu_shard_range(min(m,n), world, r): blk = ceil(min/world), i0 = min(r*blk, min).
 * Two ways to form a group:
 *   svdgpu_group_create_local(ndev, devices)   ONE process drives ndev GPUs (what svd_gpu() does for
 *                                              SVD_GPU_NGPUS > 1); all per-rank arguments are arrays of ndev.
 *   svdgpu_group_create_rank(n, rank, id128)   one process per GPU (torchrun / MPI style) on the current
 *                                              device; id128 from svdgpu_nccl_unique_id() on rank 0, handed
 *                                              to the others out of band.  Per-rank arrays have one entry.
 * Both are collective over the group's processes. */
typedef struct svdgpu_group svdgpu_group;
svdgpu_group *svdgpu_group_create_local(int ndev, const int *devices /* NULL: 0..ndev-1 */);
svdgpu_group *svdgpu_group_create_rank(int nranks, int rank, const void *id128);
void svdgpu_group_destroy(svdgpu_group *g);
int  svdgpu_group_size(const svdgpu_group *g);
int  svdgpu_group_nlocal(const svdgpu_group *g);
int  svdgpu_group_rank(const svdgpu_group *g, int local);
int  svdgpu_group_device(const svdgpu_group *g, int local);
void svdgpu_shard_range(int mn, int world, int rank, int *blk, int *i0, int *ns);
/* planning only (no device needed): the route of an m x n input (route[0] transpose, route[1] QR first, route[2..3]
 * the shape the core factorizes) and the canonical order of the compact-WY panel chunks that rank 0 prepares and
 * broadcasts — set (0 Q of the QR, 1 left, 2 right reflectors), panels [pb, pe), reflectors that must be final. */
int  svdgpu_plan_chunks(int m, int n, int world, int route[4], int max_out, int *set_out, int *pb_out, int *pe_out,
                        int *need_out);

/* Device-resident sharded SVD.  dA_root: the matrix on rank 0's device (ignored elsewhere), overwritten
 * as by svd_gpu_dev.  Per local rank lr: dsigma[lr][min(m,n)] receives ALL singular values, dUblk[lr]
 * (m x blk, ldu) / dVblk[lr] (n x blk, ldv) this rank's ns columns of U and V; streams[lr] as in
 * svd_gpu_dev (streams == NULL: the default stream everywhere). */
void svd_gpu_sharded_dev(svdgpu_group *g, int m, int n, double *dA_root, long lda, double *const *dsigma,
                         double *const *dUblk, long ldu, double *const *dVblk, long ldv, void *const *streams);
/* The same from host buffers: A, sigma as svd_gpu() (rank 0's process only); Ublk[lr] / Vblk[lr]: host
 * destination of local rank lr's block (m x ns, ld m / n x ns, ld n), copied from its own GPU.
 * svd_gpu() calls this with Ublk[lr] = U + i0*m on the group SVD_GPU_NGPUS selects.  Ublk == Vblk == NULL:
 * values only (rank 0). */
void svd_gpu_sharded(svdgpu_group *g, int m, int n, double *A, double *sigma, double *const *Ublk,
                     double *const *Vblk);
/* per-phase milliseconds of local rank `local` in the last call on this group: [0]..[6] as
 * svd_gpu_last_phase_ms (ranks other than 0 report 0 for [1],[2]); [7] = time the rank spent waiting for the
 * panels / the bidiagonal to arrive after its previous phase (rank 0: after dDC) */
void svd_gpu_group_phase_ms(svdgpu_group *g, int local, float ms[8]);

/* low-level building blocks of a sharded run, direct route only (no range guard, no QR first, no transpose):
 * bidiagonalization + dDC (dalpha[min], dbeta[min] zero padded, dsigma[min]), and the vector phases of the
 * block [i0, i0+ns) given the reflector matrix and ALL singular values. */
void svd_gpu_values_dev(int m, int n, double *dA, long lda, double *dalpha, double *dbeta,
                        double *dsigma, void *stream);
void svd_gpu_vectors_dev(int m, int n, const double *dA_mod, long lda, const double *dalpha,
                         const double *dbeta, const double *dsigma_all, int i0, int ns,
                         double *dUblk, long ldu, double *dVblk, long ldv, double *dsig_out,
                         void *stream);

/* ---- the reference driver's residual check, enabled (test-whole-svd.c:81-96 is "#if 0" there and calls an
 * l2_norm_mat that exists nowhere).  A0 is the ORIGINAL matrix (svd_gpu destroys its input).  out6:
 *   [0] ||U^T U - I||_F  [1] ||V^T V - I||_F  [2] ||A - U S V^T||_F / ||A||_F
 *   [3] |sum sigma^2 - ||A||_F^2| / ||A||_F^2   [4] ||A||_F   [5] 1.0 if sigma is ascending
 * svd_gpu_check_dev takes device pointers and nc columns of the factors: nc = min(m,n) for a whole SVD; for a
 * column block (one rank of a sharded run) [2] is ||A V - U S||_F / ||A||_F and [3] is 0.  Synchronises. */
void svd_gpu_check(int m, int n, const double *A0, const double *sigma, const double *U, const double *V,
                   double out6[6]);
void svd_gpu_check_dev(int m, int n, const double *dA0, long lda, const double *dsigma, const double *dU,
                       long ldu, const double *dV, long ldv, int nc, double out6[6], void *stream);

/* the reference drivers' input recipe (test-whole-svd.c:18-24,69-73; bidiag_dr.c:54-60,77): uniform [lo,hi)
 * from glibc rand() after srand(seed), in fill order — bench.py and the self-checking driver feed exactly the
 * matrices the reference's own drivers would */
void svdgpu_fill_rand(double *A, size_t count, double lo, double hi, unsigned seed);

#ifdef __cplusplus
}
#endif
#endif
