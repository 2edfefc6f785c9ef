/* dropin_check.c — the reference driver's flow with its dormant check switched on.
 *
 * test-whole-svd.c:29-99 fills an n x n matrix with rand_d(1,4) (glibc default seed), calls svd_gpu() and
 * then — inside "#if 0", :81-96 — would rebuild U S V^T with dgemm_simple and print the Frobenius error
 * (it cannot link there: l2_norm_mat is defined nowhere in the reference).  This client does the same with
 * malloc'd (pageable) buffers through the public headers, and evaluates the check on the GPU with
 * svd_gpu_check() (include/svd_gpu_b200.h).  Usage: dropin_check m n [ngpus]; exit code 0 iff the
 * north-star bounds hold (100 eps max(m,n) for the three Frobenius figures). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "svd_gpu.h"
#include "svd_gpu_b200.h"
#include "matrix_helper.h"

static double rand_d(double lo, double hi) { return lo + (hi - lo) * (rand() / (RAND_MAX + 1.0)); }   /* test-whole-svd.c:18-24 */

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s m n [ngpus]\n", argv[0]); return 2; }
    const int m = atoi(argv[1]), n = atoi(argv[2]);
    const int mn = m < n ? m : n;
    if (argc > 3) svd_gpu_set_option("ngpus", atoi(argv[3]));
    double *A = malloc(sizeof(double) * (size_t)m * n), *A_Copy = malloc(sizeof(double) * (size_t)m * n);
    double *sigma = malloc(sizeof(double) * mn);
    double *U = malloc(sizeof(double) * (size_t)m * m), *V = malloc(sizeof(double) * (size_t)n * n);   /* test-whole-svd.c:60-63 */
    if (!A || !A_Copy || !sigma || !U || !V) { fprintf(stderr, "out of memory\n"); return 2; }
    for (size_t i = 0; i < (size_t)m * n; ++i) { A[i] = rand_d(1., 4.); A_Copy[i] = A[i]; }
    svd_gpu(m, n, A, sigma, U, V);
    double out[6];
    svd_gpu_check(m, n, A_Copy, sigma, U, V, out);
    const double bound = 100.0 * 2.220446049250313e-16 * (m > n ? m : n);
    printf("relative error in Frobenius norm = %2.15e \n", out[2]);
    printf("orthogonality U %.3e V %.3e  checksum %.3e  ascending %d  bound %.3e\n", out[0], out[1], out[3], (int)out[5], bound);
    const int ok = out[0] <= bound && out[1] <= bound && out[2] <= bound && out[5] == 1.0;
    printf(ok ? "OK\n" : "FAILED\n");
    free(A); free(A_Copy); free(sigma); free(U); free(V);
    return ok ? 0 : 1;
}
