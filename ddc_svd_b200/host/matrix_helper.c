/* matrix_helper.c — plain host C helpers with the reference's signatures
 * (matrix_helper.h:30-45; definitions matrix_helper.c:41-174).  They are not on the GPU hot
 * path; they exist because the reference's client header promises them and its driver's
 * verification block (test-whole-svd.c:81-96) calls transpose / form_bidiag / dgemm_simple
 * and an l2_norm_mat that the reference never defined.  Column-major throughout. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/matrix_helper.h"

void print_matrix(const double *A, long m, long n, char *message)
{
    printf("%s \n", message);
    for (long r = 0; r < m; ++r) {
        for (long c = 0; c < n; ++c) printf("%11.3E", A[r + m * c]);
        putchar('\n');
    }
}

static double strided_dot(int l, const double *a, long sa, const double *b, long sb)
{
    double acc = 0;
    for (int k = 0; k < l; ++k) acc += a[k * sa] * b[k * sb];
    return acc;
}
static void strided_scale(int l, double *a, long sa, double f)
{
    for (int k = 0; k < l; ++k) a[k * sa] *= f;
}

double l2_normv(int l, const double *v) { return sqrt(strided_dot(l, v, 1, v, 1)); }
void scale_vector(int l, double *v, double scale) { strided_scale(l, v, 1, scale); }
double dot_prod(int l, const double *a, const double *b) { return strided_dot(l, a, 1, b, 1); }
double l2_norm_mat_row(int m, int n, int l, const double *row) { (void)n; return sqrt(strided_dot(l, row, m, row, m)); }
void scale_mat_row(int m, int n, int l, double *row, double scale) { (void)n; strided_scale(l, row, m, scale); }
double dot_prod_mat_rows(int m, int n, int l, const double *a, const double *b) { (void)n; return strided_dot(l, a, m, b, m); }
double dot_prod_mat_row_with_vec(int m, int n, int l, const double *row, const double *vec) { (void)n; return strided_dot(l, row, m, vec, 1); }
void set_vec_to_zero(int l, double *v) { memset(v, 0, sizeof(double) * (size_t)l); }

/* C (M x N) += A (M x L) * B (L x N) — accumulates into C like the reference (matrix_helper.c:121-131) */
void dgemm_simple(const int M, const int N, const int L, const double *A, const double *B, double *C)
{
    for (int j = 0; j < N; ++j)
        for (int k = 0; k < L; ++k) {
            const double b = B[k + (size_t)j * L];
            const double *a = A + (size_t)k * M;
            double *c = C + (size_t)j * M;
            for (int i = 0; i < M; ++i) c[i] += a[i] * b;
        }
}

/* M x N matrix with alpha on the diagonal and beta on the super-diagonal (matrix_helper.c:137-160) */
void form_bidiag(const int M, const int N, const double *alpha, const double *beta, double *mat)
{
    memset(mat, 0, sizeof(double) * (size_t)M * N);
    const int nd = M < N ? M : N;
    for (int i = 0; i < nd; ++i) {
        mat[i + (size_t)i * M] = alpha[i];
        if (i + 1 < N && (N > M || i < N - 1)) mat[i + (size_t)(i + 1) * M] = beta[i];
    }
}

/* AT (N x M, ld N) = A^T for A (M x N, ld M) (matrix_helper.c:166-174) */
void transpose(const int M, const int N, const double *A, double *AT)
{
    for (int c = 0; c < N; ++c)
        for (int r = 0; r < M; ++r) AT[c + (size_t)r * N] = A[r + (size_t)c * M];
}

double l2_norm_mat(int m, int n, const double *A)
{
    double acc = 0;
    for (size_t k = 0; k < (size_t)m * n; ++k) acc += A[k] * A[k];
    return sqrt(acc);
}

/* The input recipe of the reference's drivers (test-whole-svd.c:18-24,69-73: rand_d(1,4) in fill order with
 * glibc's default seed 1; bidiag_dr.c:54-60,77 uses [1,2) after srand(4)): A[i] = lo + (hi-lo) * rand()/(RAND_MAX+1). */
void svdgpu_fill_rand(double *A, size_t count, double lo, double hi, unsigned seed)
{
    srand(seed);
    for (size_t i = 0; i < count; ++i) A[i] = lo + (hi - lo) * (rand() / (RAND_MAX + 1.0));
}
