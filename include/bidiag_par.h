/* bidiag_par.h — phase-level exports with the reference's signatures (bidiag_par.h:30,72-73)
 * so that the bidiagonalization and the back-transform can be parity-tested on their own.
 * Host pointers; each call copies to the device, runs the sm_100a kernels and copies back. */
#ifndef BIDIAG_PAR
#define BIDIAG_PAR
#ifdef __cplusplus
extern "C" {
#endif
/* A m x n column-major (ld m) is overwritten with the reflectors; alpha[min(m,n)];
 * beta[n-1] if m >= n else beta[m]   (bidiag_par.c:34-450) */
void bidiag_par(int m, int n, double *A, double *alpha, double *beta);
/* explicit orthogonal factors, reference signatures (bidiag_par.h:33-34, bidiag_par.c:877-988):
 * U (m x m, ld m) = Q_L, V (n x n, ld n) = Q_R with A = Q_L B Q_R^T */
void form_u_par(int m, int n, const double *A_mod, double *U);
void form_v_par(int m, int n, const double *A_mod, double *V);
/* one back-transformed vector, reference calling convention (bidiag_par.c:990-1095).
 * multV takes the TRANSPOSED reflector matrix (n x m, ld n) like the reference's caller
 * passes (svd_gpu.c:103,120).  Per-vector calls re-upload the reflectors every time; use
 * svd_gpu_backtransform() for anything but tests. */
void multU(int m, int n, int vecnum, double *A_mod, double *Y, double *U);
void multV(int m, int n, int vecnum, double *AT_mod, double *X, double *V);
/* all vectors at once: U (m x mn, ld m) = Q_L [Y;0], V (n x mn, ld n) = Q_R [X;0];
 * Y is mn x mn (Y[i*mn+j]), X is mn vectors of length len_beta+1 (X[i*(len_beta+1)+j]). */
void svd_gpu_backtransform(int m, int n, const double *A_mod, const double *X, const double *Y,
                           double *U, double *V);
#ifdef __cplusplus
}
#endif
#endif
