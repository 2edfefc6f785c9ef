/* oracle/svd_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, CPU restatement of the reference's svd_gpu() hot path
 * (sddelong/ddc-svd).  It exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg have something to check the CUDA path against on
 * a machine where /root/reference does not exist.  Nothing in the shipped
 * library (ddc_svd_b200/) may include, link or call it.
 *
 * Pinning: tests/test_oracle_vs_ref.py compares every function below with the
 * unmodified reference compiled into oracle/_ref/libddcref.so (same inputs),
 * and tests/golden/ holds outputs of that reference build for machines where
 * the reference is absent.  The reference ships no golden vectors of its own
 * (SURVEY.md §8c).
 *
 * Deliberate deviations from the reference (both are reference bugs, SURVEY.md
 * facts 2 and 4), each behind an explicit, documented rule:
 *   - the super-diagonal handed to the dDC recursion is zero-padded to length N
 *     (the reference reads beta[mn-1] out of bounds, svd_gpu.c:76,106);
 *   - orc_apply_right() uses the true leading dimension n of the transposed
 *     reflector matrix (the reference's multV uses m, bidiag_par.c:1034,1036,
 *     which is only right when m == n).
 * For m == n — the only shape the reference handles correctly — both rules are
 * no-ops and the functions agree with the reference to rounding.
 */
#ifndef DDC_SVD_ORACLE_H
#define DDC_SVD_ORACLE_H

/* test-whole-svd.c:18-24,69-73 and bidiag_dr.c:54-60,133-137: glibc rand() recipe.
 * seed < 0 keeps the generator's current state (the reference driver never seeds). */
void orc_fill_rand(double *A, long count, double lo, double hi, int seed);

/* bidiag.c:33-186 (== bidiag_par.c:310-397 semantics). A is m x n column-major, overwritten
 * with unit-norm Householder vectors; alpha[min(m,n)], beta[n-1 if m>=n else m]. */
void orc_bidiag(int m, int n, double *A, double *alpha, double *beta);

/* Calculations-Parallel.c:852-874. b1,b2 both of length N (b2[N-1]=0 for a square B). */
void orc_ddc_values(int N, const double *b1, const double *b2, double *sigma);

/* parallel-twisted.c:554-637. n singular values, vectors of length m (= n or n+1), X[i*m+j]. */
void orc_right_vectors(int n, int m, const double *a, const double *b, const double *sigma,
                       double *X);
/* parallel-twisted.c:530-551. Y[i*n+j] = (B x_i)_j / sigma_i. */
void orc_left_vectors(int n, int m, const double *a, const double *b, const double *sigma,
                      const double *X, double *Y);

/* bidiag_par.c:1046-1095: out(0:m) = H_0 ... H_{last} [Y(:,vec); 0]. */
void orc_apply_left(int m, int n, int vec, const double *A_mod, const double *Y, double *out);
/* bidiag_par.c:990-1043 with the leading-dimension fix: out(0:n) = G_0 ... [X(:,vec); 0].
 * AT is the n x m transpose of A_mod (matrix_helper.c:166-174). */
void orc_apply_right(int m, int n, int vec, const double *AT, const double *X, double *out);

/* bidiag.c:252-359 (form_u, form_v; same results as bidiag_par.c:877-988 form_u_par/form_v_par):
 * U (m x m) = Q_L, V (n x n) = Q_R from the stored reflectors. */
void orc_form_u(int m, int n, const double *A_mod, double *U);
void orc_form_v(int m, int n, const double *A_mod, double *V);

/* svd_gpu.c:53-131. sigma ascending; first min(m,n) columns of U (ld m) and V (ld n). */
void orc_svd(int m, int n, double *A, double *sigma, double *U, double *V);

/* per-phase wall-clock of the last orc_svd() call, seconds:
 * [0] bidiag [1] transpose [2] dDC [3] right vectors [4] left vectors [5] back-transform */
void orc_last_timings(double t[6]);
#endif
