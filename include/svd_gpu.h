/* svd_gpu.h — the drop-in entry point.  Same name, guard, signature and argument/output
 * layout as the reference's svd_gpu.h:1-6 / svd_gpu.c:53:
 *
 *   A      m x n, column-major, leading dimension m.  OVERWRITTEN with the Householder
 *          reflectors of the bidiagonalization (as the reference's bidiag_par leaves them).
 *   sigma  min(m,n) singular values, ASCENDING (Calculations-Parallel.c:56).
 *   U      m x min(m,n) left vectors in the first min(m,n) columns, leading dimension m.
 *   V      n x min(m,n) right vectors in the first min(m,n) columns, leading dimension n.
 *          U(:,i), V(:,i) pair with sigma[i] (svd_gpu.c:118-121).
 *
 * Errors: none returned; any failure prints to stderr and abort()s (svd_gpu.c:85-96,
 * cl-helper.h:47-55).  New, backward compatible: U == NULL && V == NULL computes the
 * singular values only (the reference would dereference NULL there).
 * Never reads stdin and never opens kernel source files (the reference does both:
 * cl-helper.c:230-255, bidiag_par.c:199-256).
 */
#ifndef SVDGPU
#define SVDGPU
#ifdef __cplusplus
extern "C" {
#endif
void svd_gpu(int m, int n, double* A,double * sigma, double * U, double* V);
#ifdef __cplusplus
}
#endif
#endif
