/* parallel-twisted.h — phase-level exports, reference signatures
 * (parallel-twisted.h:18-19, parallel-twisted.c:554-637, :530-551).  Host pointers.
 * n singular values / rows of B, m = n or n+1 columns; A diagonal [n], B super-diagonal
 * [m-1]; X[i*m + j] right vectors; Y[i*n + j] left vectors. */
#ifndef PARALLELTWISTED
#define PARALLELTWISTED
#ifdef __cplusplus
extern "C" {
#endif
void CalcRightSingularVectors(int n, int m, double* A, double* B,double* sigma, double* X);
void RighttoLeftSingularVectors(int n, int m, double* A, double * B, double * sigma, double * X, double * Y);
#ifdef __cplusplus
}
#endif
#endif
