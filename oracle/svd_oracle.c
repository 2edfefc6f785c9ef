/* oracle/svd_oracle.c — TEST INFRASTRUCTURE ONLY (see svd_oracle.h).
 *
 * CPU restatement of the reference's svd_gpu() path, phase by phase.  It is a
 * restatement, not a copy: the arithmetic (expression shapes, summation order,
 * thresholds, iteration limits) follows the reference so that the two agree to
 * rounding, but the code is organised differently (shared helpers, explicit
 * permutations, heap instead of C99 VLAs).  Every function cites the reference
 * lines it follows.
 */
#include "svd_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *xmalloc(size_t bytes)
{
    void *p = malloc(bytes ? bytes : 1);
    if (!p) { fprintf(stderr, "svd_oracle: out of memory (%zu bytes)\n", bytes); abort(); }
    return p;
}

/* ------------------------------------------------------------------ inputs */

void orc_fill_rand(double *A, long count, double lo, double hi, int seed)
{
    /* test-whole-svd.c:18-24: c = rand()/(RAND_MAX+1.0); a + c*(b-a). */
    if (seed >= 0) srand((unsigned)seed);
    for (long i = 0; i < count; ++i) {
        double c = rand() / (1.0 * RAND_MAX + 1);
        A[i] = lo + c * (hi - lo);
    }
}

/* ------------------------------------------------- phase 1: bidiagonalization */

/* The reflector recipe shared by the column and the row case
 * (bidiag.c:76-94 / :100-122, update_scale_matcol.cl:45-82):
 * returns -sign(x0)*||x||, leaves the unit-norm Householder vector in x. */
static double make_reflector(double *x, long stride, int len)
{
    double x0 = x[0];
    int sgn = (x0 < 0) ? -1 : 1;
    double ss = 0.0;
    for (int k = 0; k < len; ++k) ss += x[k * stride] * x[k * stride];
    double nu = sqrt(ss);
    x[0] += sgn * nu;
    double scale = sqrt(2.0) * sqrt(nu * nu + fabs(nu * x0));
    double inv = 1.0 / scale;
    for (int k = 0; k < len; ++k) x[k * stride] *= inv;
    return -sgn * nu;
}

void orc_bidiag(int m, int n, double *A, double *alpha, double *beta)
{
    const long ld = m;
    const int mn = (m < n) ? m : n;
    double *acc = (double *)xmalloc(sizeof(double) * (size_t)m);

    for (int i = 0; i < mn - 1; ++i) {
        /* column reflector, then H applied to the columns on its right (bidiag.c:188-218) */
        double *v = A + i + i * ld;
        const int L = m - i;
        alpha[i] = make_reflector(v, 1, L);
        for (int j = i + 1; j < n; ++j) {
            double *col = A + i + j * ld;
            double ip = 0.0;
            for (int r = 0; r < L; ++r) ip += v[r] * col[r];
            for (int r = 0; r < L; ++r) col[r] -= 2 * v[r] * ip;
        }
        if (i < n - 2) {
            /* row reflector, then G applied to the rows below (bidiag.c:220-250).  The row
             * dot products are accumulated column by column, which visits the terms of each
             * row's sum in the same order as the reference's row-by-row loop. */
            double *u = A + i + (i + 1) * ld;
            const int R = n - i - 1, Lb = m - i - 1;
            beta[i] = make_reflector(u, ld, R);
            double *blk = A + (i + 1) + (i + 1) * ld;
            for (int r = 0; r < Lb; ++r) acc[r] = 0.0;
            for (int j = 0; j < R; ++j) {
                const double uj = u[j * ld];
                const double *col = blk + j * ld;
                for (int r = 0; r < Lb; ++r) acc[r] += uj * col[r];
            }
            for (int j = 0; j < R; ++j) {
                const double uj = u[j * ld];
                double *col = blk + j * ld;
                for (int r = 0; r < Lb; ++r) col[r] -= 2 * uj * acc[r];
            }
        } else {
            beta[i] = A[i + (i + 1) * ld];
            A[i + (i + 1) * ld] = 0;
        }
    }
    if (n >= m + 1) {
        /* wide: last row reflector, diagonal entry moved out (bidiag.c:137-163) */
        beta[mn - 1] = make_reflector(A + (mn - 1) + mn * ld, ld, n - mn);
        alpha[mn - 1] = A[(mn - 1) + (mn - 1) * ld];
        A[(mn - 1) + (mn - 1) * ld] = 0;
    } else {
        /* tall / square: last column reflector (bidiag.c:165-183) */
        alpha[mn - 1] = make_reflector(A + (mn - 1) + (mn - 1) * ld, 1, m - mn + 1);
    }
    free(acc);
}

/* ------------------------------------------- phase 2: dDC singular values */

/* Merge order of d = [0 | run1 (K values) | run2 (N-K-1 values)]
 * (Calculations-Parallel.c:65-104 and :371-448): src[k] is the position in d
 * of the k-th smallest; ties take the first run. */
static void merge_order(int K, int N, const double *d, int *src)
{
    int i = 1, j = K + 1, k = 1;
    src[0] = 0;
    while (i <= K && j < N) src[k++] = (d[i] <= d[j]) ? i++ : j++;
    while (i <= K) src[k++] = i++;
    while (j < N) src[k++] = j++;
}

/* One root of 1 + sum_j c_j/(t_j - t_i - 1/gamma) = 0 in the reference's gamma
 * variable (Calculations-Parallel.c:111-344).  t, c sorted ascending in t. */
static double secular_root(int N, int i, const double *t, const double *c, double *d2)
{
    const double eps = 1e-12;
    double prev = 9999.0, result = 0.0, iter = 0, maxiter = 30;
    double gamma, g, gp, hold;
    int k;

    for (k = 0; k < N; ++k) d2[k] = t[k] - t[i];
    if (c[i] < 1e-20) return sqrt(t[i]) + 1e-14;               /* :143-144 */

    if (i < N - 1) {                                              /* :145-255 */
        const double delta = d2[i + 1];
        const double a1 = -c[i];
        const double a2_start = 1.0 + (c[i] + c[i + 1]) / delta;
        const double a3 = -1.0 / delta;
        double a2 = a2_start, bp;
        gamma = (-a2 - sqrt(a2 * a2 - 4.0 * a1 * a3)) / (2.0 * a1);
        g = 0.0; gp = 1.0;
        while ((gp > 0.0 || fabs(g) > 1e-5 ||
                fabs(prev - result) / (fabs(prev) + fabs(result)) > eps) && iter < maxiter) {
            prev = result;
            a2 = a2_start; bp = 0.0;
            for (k = 0; k < N; ++k) {
                if (k == i || k == i + 1) continue;
                double den = d2[k] * gamma - 1.0;
                a2 += c[k] * (gamma + a3) / den;
                bp += -c[k] * (d2[k] * a3 + 1.0) / (den * den);
            }
            g = a3 + a2 * gamma + a1 * gamma * gamma;
            gp = 2.0 * a1 * gamma + a2 + gamma * bp;
            hold = gamma - (g / gp);
            if (hold < 1.0 / delta) {
                gamma = gamma / 2.0 + (1.0 / (2.0 * delta));
            } else if (g > 0.0 && gp > 0) {
                if (iter < 5) gamma = ((gamma > 1.0) ? gamma : 1.0) * 10.0;
                else gamma = (-a2 - sqrt(a2 * a2 - 4.0 * a1 * a3)) / (2.0 * a1);
            } else if (g > 0.0 && gamma > 10e24) {
                prev = result; gp = 0.0; g = 0.0;
            } else {
                gamma = hold;
            }
            iter++;
            result = gamma;
        }
        if (iter == maxiter) {                                    /* bisection, :200-254 */
            double lo = 1 / delta, hi = -1.0;
            gamma = lo;
            while (hi < 0) {
                gamma *= 10.0;
                a2 = a2_start;
                for (k = 0; k < N; ++k)
                    if (k != i && k != i + 1) a2 += c[k] * (gamma + a3) / (d2[k] * gamma - 1.0);
                g = a3 + a2 * gamma + a1 * gamma * gamma;
                if (g < 0.0) hi = gamma; else lo = gamma;
            }
            gamma = (lo + hi) / 2.0;
            while (fabs(lo - hi) / (fabs(lo) + fabs(hi)) > eps) {
                a2 = a2_start;
                for (k = 0; k < N; ++k)
                    if (k != i && k != i + 1) a2 += c[k] * (gamma + a3) / (d2[k] * gamma - 1.0);
                g = a3 + a2 * gamma + a1 * gamma * gamma;
                if (g > 0.0) lo = gamma; else hi = gamma;
                gamma = (lo + hi) / 2.0;
            }
        }
        return sqrt(t[i] + 1.0 / gamma);
    }

    /* last root, above every pole (:256-343) */
    gamma = (c[i] > 1e-14 ? 1.0 / c[i] : 1);
    g = 0.0; gp = 1.0;
    while ((gp > 0 || fabs(g) > 1e-5 ||
            fabs(prev - result) / (fabs(prev) + fabs(result)) > eps) && iter < maxiter) {
        prev = result;
        g = 1 - c[i] * gamma;
        gp = -c[i];
        for (k = 0; k < N - 1; ++k) {
            double den = d2[k] * gamma - 1.0;
            g += c[k] * gamma / den;
            gp -= c[k] / (den * den);
        }
        hold = gamma - g / gp;
        if (hold < 0.0) gamma = gamma / 2.0;
        else if (g > 0.0 && gamma > 10e24) { prev = result; gp = 0.0; g = 0.0; }
        else gamma = hold;
        result = gamma;
        iter++;
    }
    if (iter == maxiter) {
        double lo = 0.0, hi = -1.0;
        gamma = 0.1;
        while (hi < 0.0) {
            gamma *= 10.0;
            g = 1 - c[i] * gamma;
            for (k = 0; k < N - 1; ++k) g += c[k] * gamma / (d2[k] * gamma - 1);
            if (g < 0.0) hi = gamma; else lo = gamma;
        }
        gamma = (lo + hi) / 2.0;
        while (fabs(lo - hi) / (fabs(lo) + fabs(hi)) > eps) {
            g = 1 - c[i] * gamma;
            for (k = 0; k < N - 1; ++k) g += c[k] * gamma / (d2[k] * gamma - 1);
            if (g > 0.0) lo = gamma; else hi = gamma;
            gamma = (lo + hi) / 2.0;
        }
    }
    return sqrt(t[i] + 1.0 / gamma);
}

/* Calculations-Parallel.c:48-348 */
static void secular_all(int K, int N, const double *d, const double *z, double *sigma)
{
    int *src = (int *)xmalloc(sizeof(int) * (size_t)N);
    double *t = (double *)xmalloc(sizeof(double) * 2 * (size_t)N), *c = t + N;
    merge_order(K, N, d, src);
    for (int k = 0; k < N; ++k) {
        t[k] = (k == 0) ? 0.0 : d[src[k]] * d[src[k]];
        c[k] = z[src[k]] * z[src[k]];
    }
#pragma omp parallel
    {
        double *d2 = (double *)xmalloc(sizeof(double) * (size_t)N);
#pragma omp for
        for (int i = 0; i < N; ++i) sigma[i] = secular_root(N, i, t, c, d2);
        free(d2);
    }
    free(t); free(src);
}

/* Calculations-Parallel.c:351-588: first/last rows of the merged V from the
 * arrow-matrix eigenvectors, with the Loewner-recomputed z evaluated in logs. */
static void merged_rows(int K, int N, const double *d, const double *sigma, const double *z,
                        double *first_row, double *last_row, int want_first, int want_last)
{
    int *src = (int *)xmalloc(sizeof(int) * (size_t)N);
    double *w = (double *)xmalloc(sizeof(double) * 9 * (size_t)N);
    double *t = w, *zo = w + N, *s2 = w + 2 * N, *hf = w + 3 * N, *hl = w + 4 * N;
    double *zh = w + 5 * N, *zsgn = w + 6 * N, *nv = w + 7 * N, *nu = w + 8 * N;

    merge_order(K, N, d, src);
    for (int k = 0; k < N; ++k) {
        t[k] = (k == 0) ? 0.0 : d[src[k]] * d[src[k]];
        zo[k] = z[src[k]];
        if (want_first) hf[k] = first_row[src[k]];
        if (want_last) hl[k] = last_row[src[k]];
        s2[k] = sigma[k] * sigma[k];
    }
    /* log of the recomputed z (:462-481) */
#pragma omp parallel for
    for (int i = 0; i < N; ++i) {
        if (fabs(t[i] - s2[i]) < 1e-14 || (i > 0 && fabs(t[i] - s2[i - 1]) < 1e-14)) {
            zh[i] = 0.0;
        } else {
            double a = log(s2[N - 1] - t[i]);
            for (int j = 0; j < i; ++j) a += log(t[i] - s2[j]) - log(t[i] - t[j]);
            for (int j = i; j < N - 1; ++j) a += log(s2[j] - t[i]) - log(t[j + 1] - t[i]);
            zh[i] = a / 2;
        }
    }
    /* vector norms (:489-511) */
#pragma omp parallel for
    for (int i = 0; i < N; ++i) {
        if (zh[i] == 0.0) { nv[i] = 1.0; nu[i] = 1.0; continue; }
        double sv = 0.0, su = 1.0;
        for (int j = 0; j < N; ++j) {
            double term = exp(zh[j] - log(fabs(t[j] - s2[i])));
            term = term * term;
            sv += term;
            su += term * t[j];
        }
        nv[i] = sqrt(sv); nu[i] = sqrt(su);
    }
    /* signs (:515-536) */
#pragma omp parallel for
    for (int i = 0; i < N; ++i) {
        if (zh[i] == 0.0) {
            if (fabs(t[i] - s2[i]) < 1e-14) zsgn[i] = (zo[i] > 0 ? 1 : -1);
            else zsgn[i] = (zo[i] > 0 ? -1 : 1);
        } else {
            double term = 0.0;
            for (int k = 0; k < N; ++k) term += sigma[k] / nv[k] / nu[k] / (t[i] - s2[k]);
            zsgn[i] = -1 * (zo[i] > 0 ? 1 : -1) * (term > 0 ? 1 : -1);
        }
    }
    /* rotate the rows (:541-584) */
#pragma omp parallel
    {
        double *v = (double *)xmalloc(sizeof(double) * (size_t)N);
#pragma omp for
        for (int i = 0; i < N; ++i) {
            for (int j = 0; j < N; ++j) {
                if (zh[i] == 0.0) v[j] = (j == i ? 1.0 : 0.0);
                else if (zh[j] == 0.0) v[j] = 0.0;
                else {
                    v[j] = exp(zh[j] - log((j > i ? t[j] - s2[i] : s2[i] - t[j])));
                    v[j] *= (j > i ? 1 : -1);
                }
                v[j] *= zsgn[j];
            }
            double f = 0.0, l = 0.0;
            for (int j = 0; j < N; ++j) {
                v[j] /= nv[i];
                if (want_first) f += v[j] * hf[j];
                if (want_last) l += v[j] * hl[j];
            }
            if (want_first) first_row[i] = f;
            if (want_last) last_row[i] = l;
        }
        free(v);
    }
    free(w); free(src);
}

/* Calculations-Parallel.c:590-703: closed forms for 1x2 and 2x3 blocks. */
static void ddc_leaf(int N, const double *b1, const double *b2, double *sigma, double *first,
                     double *last, double *phi, double *psi, int need)
{
    const int wf = (need == 1 || need == 3), wl = (need == 2 || need == 3);
    if (N == 1) {
        sigma[0] = sqrt(b1[0] * b1[0] + b2[0] * b2[0]);
        if (wf) first[0] = b1[0] / sigma[0];
        if (wl) last[0] = b2[0] / sigma[0];
        *phi = b2[0] / sigma[0];
        *psi = -b1[0] / sigma[0];
        return;
    }
    double a = b1[0] * b1[0] + b2[0] * b2[0];
    double d = b1[1] * b1[1] + b2[1] * b2[1];
    double z0 = (a + d) / 2.0, z1 = 0.0, z2 = (a - d) / 2.0, z3 = b2[0] * b1[1];
    double e = z1 * z1 + z2 * z2 + z3 * z3;
    double n1 = z0 * z0 + e, n2 = z0 * z0 - e;
    sigma[0] = sqrt(sqrt(n1 - sqrt(n1 * n1 - n2 * n2)));
    sigma[1] = sqrt(sqrt(n1 + sqrt(n1 * n1 - n2 * n2)));
    *phi = (b2[1] == 0.0 ? 0.0 : (b1[0] == 0.0 ? 1.0 : -b2[0] / b1[0]));
    *psi = (b2[1] == 0.0 ? 1.0 : (b1[0] == 0.0 ? 0.0 : -b1[1] / b2[1]));
    double nrm = ((b2[1] == 0.0 || b1[0] == 0.0) ? 1.0 : sqrt(1 + *phi * *phi + *psi * *psi));
    *phi /= nrm; *psi /= nrm;
    for (int k = 0; k < 2; ++k) {
        double v1 = -b2[0] * b1[0] / (b1[0] * b1[0] - sigma[k] * sigma[k]);
        double v3 = -b2[1] * b1[1] / (b2[1] * b2[1] - sigma[k] * sigma[k]);
        nrm = sqrt(1 + v1 * v1 + v3 * v3);
        if (wf) first[k] = v1 / nrm;
        if (wl) last[k] = v3 / nrm;
    }
}

/* Calculations-Parallel.c:706-850.  sigma and scratch swap roles at every level. */
static void ddc_node(int N, const double *b1, const double *b2, double *sigma, double *scratch,
                     double *first_row, double *last_row, double *phi, double *psi, int need)
{
    if (N <= 2) { ddc_leaf(N, b1, b2, sigma, first_row, last_row, phi, psi, need); return; }

    const int K = N / 2, N2 = N - K - 1;
    const int wf = (need == 1 || need == 3), wl = (need == 2 || need == 3);
    double phi_c[2], psi_c[2];
    double *last1 = (double *)xmalloc(sizeof(double) * (size_t)(K + N2 + N));
    double *first2 = last1 + K, *z = first2 + N2;

    scratch[0] = 0;
    ddc_node(K, b1, b2, scratch + 1, sigma + 1, wf ? first_row + 1 : NULL, last1,
             &phi_c[0], &psi_c[0], wf ? 3 : 2);
    ddc_node(N2, b1 + K + 1, b2 + K + 1, scratch + K + 1, sigma + K + 1, first2,
             wl ? last_row + K + 1 : NULL, &phi_c[1], &psi_c[1], wl ? 3 : 1);

    double p = b1[K] * psi_c[0], q = b2[K] * phi_c[1];
    double r0 = sqrt(p * p + q * q), c0 = p / r0, s0 = q / r0;
    z[0] = r0;
    for (int l = 1; l <= K; ++l) z[l] = b1[K] * last1[l - 1];
    for (int l = K + 1; l < N; ++l) z[l] = b2[K] * first2[l - K - 1];

    secular_all(K, N, scratch, z, sigma);

    if (wf) { first_row[0] = c0 * phi_c[0]; for (int l = K + 1; l < N; ++l) first_row[l] = 0.0; }
    if (wl) { last_row[0] = s0 * psi_c[1]; for (int l = 1; l <= K; ++l) last_row[l] = 0.0; }
    if (need != 0) {
        merged_rows(K, N, scratch, sigma, z, first_row, last_row, wf, wl);
        *phi = -s0 * phi_c[0];
        *psi = c0 * psi_c[1];
    }
    if (wf) {                                                     /* :812-827 */
        double s = 0.0;
        for (int l = 0; l < N; ++l) s += first_row[l] * first_row[l];
        s = sqrt(s + *phi * *phi);
        for (int l = 0; l < N; ++l) first_row[l] /= s;
        *phi /= s;
    }
    if (wl) {                                                     /* :829-845 */
        double s = 0.0;
        for (int l = 0; l < N; ++l) s += last_row[l] * last_row[l];
        s = sqrt(s + *psi * *psi);
        for (int l = 0; l < N; ++l) last_row[l] /= s;
        *psi /= s;
    }
    free(last1);
}

void orc_ddc_values(int N, const double *b1, const double *b2, double *sigma)
{
    double *scratch = (double *)xmalloc(sizeof(double) * (size_t)N);
    double phi, psi;
    ddc_node(N, b1, b2, sigma, scratch, NULL, NULL, &phi, &psi, 0);
    free(scratch);
}

/* --------------------------------------------- phase 3: twisted vectors */

void orc_right_vectors(int n, int m, const double *a, const double *b, const double *sigma,
                       double *X)
{
    /* T = B^T B (parallel-twisted.c:290-316): diagonal td[m], off-diagonal te[m-1] */
    double *td = (double *)xmalloc(sizeof(double) * 2 * (size_t)m), *te = td + m;
    td[0] = a[0] * a[0];
    for (int i = 0; i < n - 1; ++i) { te[i] = a[i] * b[i]; td[i + 1] = a[i + 1] * a[i + 1] + b[i] * b[i]; }
    if (m > n) { te[n - 1] = a[n - 1] * b[n - 1]; td[n] = b[n - 1] * b[n - 1]; }

#pragma omp parallel
    {
        double *w = (double *)xmalloc(sizeof(double) * 6 * (size_t)m);
        double *D1 = w, *D2 = w + m, *P = w + 2 * m, *Q = w + 3 * m, *gam = w + 4 * m, *xt = w + 5 * m;
#pragma omp for
        for (int k = 0; k < n; ++k) {
            const double s2 = sigma[k] * sigma[k];
            double *x = X + (size_t)k * m;
            /* LDL^T forward and UDU^T backward of T - sigma^2 I (:339-360) */
            D1[0] = td[0] - s2;
            D2[m - 1] = td[m - 1] - s2;
            for (int i = 0; i < m - 1; ++i) {
                P[i] = te[i] / D1[i];
                D1[i + 1] = td[i + 1] - s2 - P[i] * P[i] * D1[i];
                Q[m - 2 - i] = te[m - 2 - i] / D2[m - 1 - i];
                D2[m - 2 - i] = td[m - 2 - i] - s2 - D2[m - 1 - i] * Q[m - 2 - i] * Q[m - 2 - i];
            }
            /* gamma and its smallest-magnitude entry, later index on ties (:458-488, :240-288) */
            gam[0] = D1[0] + D2[0] - (a[0] * a[0] - s2);
            for (int j = 1; j < n; ++j)
                gam[j] = D1[j] + D2[j] - (a[j] * a[j] + b[j - 1] * b[j - 1] - s2);
            if (m > n) gam[m - 1] = D1[m - 1] + D2[m - 1] - (b[n - 1] * b[n - 1] - s2);
            int kk = 0;
            for (int j = 1; j < m; ++j) if (!(fabs(gam[kk]) < fabs(gam[j]))) kk = j;
            /* solve N_k x = e_k outward from the twist (:495-521) */
            x[kk] = 1.0;
            for (int j = kk + 1; j < m; ++j) x[j] = -1.0 * Q[j - 1] * x[j - 1];
            for (int j = kk - 1; j >= 0; --j) x[j] = -1.0 * P[j] * x[j + 1];
            /* one more solve with the factorization twisted at m/2 (:392-424) */
            const int h = m / 2;
            xt[0] = x[0];
            xt[m - 1] = x[m - 1];
            for (int j = 1; j < h; ++j) xt[j] = x[j] - P[j - 1] * xt[j - 1];
            for (int j = m - 2; j > h; --j) xt[j] = x[j] - Q[j] * xt[j + 1];
            xt[h] = x[h] - Q[h] * xt[h + 1] - P[h - 1] * xt[h - 1];
            x[h] = xt[h] / gam[h];
            for (int j = h + 1; j < m; ++j) x[j] = (xt[j] - D2[j] * Q[j - 1] * x[j - 1]) / D2[j];
            for (int j = h - 1; j >= 0; --j) x[j] = (xt[j] - D1[j] * P[j] * x[j + 1]) / D1[j];
            /* normalise (:106-120) */
            double ss = 0.0;
            for (int j = 0; j < m; ++j) ss = ss + x[j] * x[j];
            ss = sqrt(ss);
            for (int j = 0; j < m; ++j) x[j] = x[j] / ss;
        }
        free(w);
    }
    free(td);
}

void orc_left_vectors(int n, int m, const double *a, const double *b, const double *sigma,
                      const double *X, double *Y)
{
    /* parallel-twisted.c:58-86 and :530-551 */
#pragma omp parallel for
    for (int i = 0; i < n; ++i) {
        const double *x = X + (size_t)i * m;
        double *y = Y + (size_t)i * n;
        for (int j = 0; j < n - 1; ++j) y[j] = a[j] * x[j] + b[j] * x[j + 1];
        y[n - 1] = (n == m) ? a[n - 1] * x[n - 1] : a[n - 1] * x[n - 1] + b[n - 1] * x[n];
        for (int j = 0; j < n; ++j) y[j] = y[j] / sigma[i];
    }
}

/* ---------------------------------------------- phase 4: back-transform */

void orc_apply_left(int m, int n, int vec, const double *A_mod, const double *Y, double *out)
{
    /* bidiag_par.c:1046-1095 */
    const int mn = (m < n) ? m : n;
    const int last = (m < n) ? mn : mn - 1;
    for (int i = 0; i < mn; ++i) out[i] = Y[(size_t)vec * mn + i];
    for (int i = mn; i < m; ++i) out[i] = 0.0;
    for (int j = last; j >= 0; --j) {
        const double *v = A_mod + j + (size_t)j * m;
        double ip = 0.0;
        for (int k = 0; k < m - j; ++k) ip += out[j + k] * v[k];
        for (int k = j; k < m; ++k) out[k] -= 2 * A_mod[k + (size_t)j * m] * ip;
    }
}

void orc_apply_right(int m, int n, int vec, const double *AT, const double *X, double *out)
{
    /* bidiag_par.c:990-1043; AT is n x m with leading dimension n (the reference indexes it
     * with m, which coincides only for square inputs — SURVEY.md fact 2).  X holds vectors of
     * length len_beta+1 (the reference strides it by mn, right only when m >= n). */
    const int mn = (m < n) ? m : n;
    const int nref = (m < n) ? mn : mn - 1;
    const int xl = nref + 1;
    for (int i = 0; i < xl && i < n; ++i) out[i] = X[(size_t)vec * xl + i];
    for (int i = xl; i < n; ++i) out[i] = 0.0;
    for (int j = (nref + 1 < m ? nref + 1 : m - 1); j >= 0; --j) {
        const double *r = AT + (j + 1) + (size_t)j * n;
        double ip = 0.0;
        for (int k = 0; k < n - j - 1; ++k) ip += r[k] * out[j + 1 + k];
        for (int k = j + 1; k < n; ++k) out[k] -= 2 * AT[(size_t)j * n + k] * ip;
    }
}

/* explicit orthogonal factors (bidiag.c:252-359 form_u / form_v == bidiag_par.c:877-988):
 * column i of U is Q_L e_i, column i of V is Q_R e_i, reflectors applied last to first;
 * reflectors with an index above i leave e_i unchanged, so they are skipped as in the reference */
void orc_form_u(int m, int n, const double *A_mod, double *U)
{
    const int mn = (m < n) ? m : n;
    for (int i = 0; i < m; ++i) {
        double *u = U + (size_t)i * m;
        for (int r = 0; r < m; ++r) u[r] = 0;
        u[i] = 1;
        for (int j = (i < mn - 1 ? i : mn - 1); j >= 0; --j) {
            const double *v = A_mod + j + (size_t)j * m;
            double ip = 0.0;
            for (int k = 0; k < m - j; ++k) ip += u[j + k] * v[k];
            for (int k = j; k < m; ++k) u[k] -= 2 * A_mod[k + (size_t)j * m] * ip;
        }
    }
}

void orc_form_v(int m, int n, const double *A_mod, double *V)
{
    const int mn = (m < n) ? m : n;
    const int nref = (m < n) ? mn : mn - 1;
    for (int i = 0; i < n; ++i) {
        double *v = V + (size_t)i * n;
        for (int r = 0; r < n; ++r) v[r] = 0;
        v[i] = 1;
        for (int j = (i < nref ? i - 1 : nref - 1); j >= 0; --j) {
            const double *row = A_mod + j + (size_t)(j + 1) * m;      /* row reflector j, stride m */
            double ip = 0.0;
            for (int k = 0; k < n - j - 1; ++k) ip += row[(size_t)k * m] * v[j + 1 + k];
            for (int k = j + 1; k < n; ++k) v[k] -= 2 * A_mod[j + (size_t)k * m] * ip;
        }
    }
}

/* ------------------------------------------------------ the whole path */

static double g_t[6];
void orc_last_timings(double t[6]) { memcpy(t, g_t, sizeof g_t); }

void orc_svd(int m, int n, double *A, double *sigma, double *U, double *V)
{
    /* svd_gpu.c:53-131 */
    const int mn = (m >= n) ? n : m;
    const int len_beta = (m >= n) ? n - 1 : m;
    const int xl = len_beta + 1;
    double *AT = (double *)xmalloc(sizeof(double) * (size_t)m * n);
    double *alpha = (double *)xmalloc(sizeof(double) * (size_t)mn);
    double *beta = (double *)xmalloc(sizeof(double) * (size_t)(mn + 1));
    double *X = (double *)xmalloc(sizeof(double) * (size_t)mn * xl);
    double *Y = (double *)xmalloc(sizeof(double) * (size_t)mn * mn);
    double t0;

    t0 = now_s();
    orc_bidiag(m, n, A, alpha, beta);
    if (len_beta < mn) beta[mn - 1] = 0.0;      /* the reference reads this slot unset (fact 4) */
    g_t[0] = now_s() - t0; t0 = now_s();
    for (int i = 0; i < n; ++i)                   /* matrix_helper.c:166-174 */
        for (int j = 0; j < m; ++j) AT[i + (size_t)j * n] = A[j + (size_t)i * m];
    g_t[1] = now_s() - t0; t0 = now_s();
    orc_ddc_values(mn, alpha, beta, sigma);
    g_t[2] = now_s() - t0; t0 = now_s();
    orc_right_vectors(mn, xl, alpha, beta, sigma, X);
    g_t[3] = now_s() - t0; t0 = now_s();
    orc_left_vectors(mn, xl, alpha, beta, sigma, X, Y);
    g_t[4] = now_s() - t0; t0 = now_s();
#pragma omp parallel for
    for (int i = 0; i < mn; ++i) {
        orc_apply_left(m, n, i, A, Y, U + (size_t)i * m);
        orc_apply_right(m, n, i, AT, X, V + (size_t)i * n);
    }
    g_t[5] = now_s() - t0;
    free(AT); free(alpha); free(beta); free(X); free(Y);
}
