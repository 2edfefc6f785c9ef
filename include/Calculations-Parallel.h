/* Calculations-Parallel.h — phase-level export, reference signature
 * (Calculations-Parallel.h:41, Calculations-Parallel.c:852-874).  Host pointers.
 * b1 (diagonal) and b2 (super-diagonal) both have N entries: the matrix is N x (N+1);
 * pass b2[N-1] = 0 for a square bidiagonal.  sigma[N] ascending. */
#ifndef SVDGPU_CALCULATIONS_PARALLEL_H
#define SVDGPU_CALCULATIONS_PARALLEL_H
#ifdef __cplusplus
extern "C" {
#endif
void GetSingularValues_Parallel( int N , double b1[] , double b2[] , double sigma[] );
#ifdef __cplusplus
}
#endif
#endif
