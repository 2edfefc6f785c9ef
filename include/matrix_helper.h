/* matrix_helper.h — the host helpers the reference's client header promises
 * (matrix_helper.h:30-45).  test-whole-svd.c:5 includes this header and its (disabled)
 * check block calls transpose / form_bidiag / dgemm_simple, so they exist here as plain
 * host C with the reference's signatures and column-major conventions. */
#ifndef MATRIX_HELPER
#define MATRIX_HELPER
#ifdef __cplusplus
extern "C" {
#endif
void print_matrix(const double * A, long m, long n, char* message);
double l2_normv(int l, const double* v);
void scale_vector(int l, double* v, double scale);
double dot_prod(int l, const double* a, const double* b);
double l2_norm_mat_row(int m, int n, int l, const double* row);
void scale_mat_row(int m, int n, int l, double* row, double scale);
double dot_prod_mat_rows(int m, int n, int l, const double* a, const double* b);
double dot_prod_mat_row_with_vec(int m, int n, int l, const double* row, const double* vec);
void set_vec_to_zero(int l, double* v);
void dgemm_simple( const int M, const int N, const int L, const double *A, const double *B, double *C);
void form_bidiag( const int M, const int N, const double *alpha, const double *beta, double * mat);
void transpose( const int M, const int N, const double * A, double * AT);
/* the Frobenius norm test-whole-svd.c:93-94 calls but the reference never defines */
double l2_norm_mat(int m, int n, const double *A);
#ifdef __cplusplus
}
#endif
#endif
