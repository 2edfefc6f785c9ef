/* svd_gpu.c — the C host orchestrator behind the drop-in entry point svd_gpu().
 *
 * Same phase sequence as the reference's svd_gpu.c:100-121
 *     bidiag_par -> (transpose) -> GetSingularValues_Parallel -> CalcRightSingularVectors
 *     -> RighttoLeftSingularVectors -> multU / multV per vector
 * but every phase runs on the GPU through the C-ABI layer (include/cuda-helper.h); the
 * matrix crosses PCIe once in each direction, the host transpose disappears (the row
 * reflectors are gathered on the device) and the n x n intermediates X, Y are written
 * straight into the device images of V and U.  Host code is plain C, as in the reference.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../../include/svd_gpu_b200.h"
#include "../../include/cuda-helper.h"

typedef struct {
    int inited;
    void *stream, *copy_stream;
    void *ev[6];
    void *ev_copy;
    void *ev_first;           /* the first of U / V to be final (the other one is still being back-transformed) */
    int first_is_u, first_recorded;
    void *arena;
    size_t arena_bytes;
    int nb, rqi;
    int qr_first;             /* 1: m >= qr_ratio10/10 * n goes through QR first */
    int qr_ratio10;
    int wide_transpose;       /* 1: m < n is solved as the SVD of A^T (U and V swapped) */
    float ms[7];
    int ms_pending;           /* events recorded but not yet read */
} svd_ctx;

static svd_ctx g;

static size_t up256(size_t b) { return (b + 255) / 256 * 256; }
static size_t maxz(size_t a, size_t b) { return a > b ? a : b; }
static double wall_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

static void ctx_init(void)
{
    if (g.inited) return;
    const char *e;
    (void)svdgpu_device_count();                 /* aborts loudly when there is no GPU */
    if ((e = getenv("SVD_GPU_DEVICE")) != NULL) svdgpu_set_device(atoi(e));
    g.nb = 32; g.rqi = 1; g.qr_first = 1; g.qr_ratio10 = 25; g.wide_transpose = 1;
    if ((e = getenv("SVD_GPU_QR_FIRST")) != NULL) g.qr_first = atoi(e);
    if ((e = getenv("SVD_GPU_QR_RATIO10")) != NULL) g.qr_ratio10 = atoi(e);
    if ((e = getenv("SVD_GPU_WIDE_TRANSPOSE")) != NULL) g.wide_transpose = atoi(e);
    if ((e = getenv("SVD_GPU_NB")) != NULL) g.nb = atoi(e);
    if ((e = getenv("SVD_GPU_RQI")) != NULL) g.rqi = atoi(e);
    g.stream = svdgpu_stream_create();
    g.copy_stream = svdgpu_stream_create();
    for (int i = 0; i < 6; ++i) g.ev[i] = svdgpu_event_create();
    g.ev_copy = svdgpu_event_create();
    g.ev_first = svdgpu_event_create();
    g.inited = 1;
}

void svd_gpu_set_option(const char *name, int value)
{
    ctx_init();
    if (!strcmp(name, "nb")) g.nb = value;
    else if (!strcmp(name, "rqi")) g.rqi = value;
    else if (!strcmp(name, "qr_first")) g.qr_first = value;
    else if (!strcmp(name, "qr_ratio10")) g.qr_ratio10 = value;
    else if (!strcmp(name, "wide_transpose")) g.wide_transpose = value;
    else if (!strcmp(name, "release")) {          /* drop the cached device arena */
        svdgpu_free(g.arena); g.arena = NULL; g.arena_bytes = 0;
    } else { fprintf(stderr, "svd_gpu_set_option: unknown option '%s'\n", name); abort(); }
}

static char *arena_get(size_t bytes)
{
    if (bytes > g.arena_bytes) {
        svdgpu_free(g.arena);
        g.arena = svdgpu_malloc(bytes);
        g.arena_bytes = bytes;
    }
    return (char *)g.arena;
}

/* number of Householder reflectors bidiag leaves on each side (bidiag_par.c:1046-1060, :1014-1023):
 * the extra ones counted here are zero vectors (H = I), harmless to include */
static int n_left(int m, int n) { return m < n ? m : n; }
static int n_right(int m, int n) { return m >= n ? (n >= 2 ? n - 2 : 0) : m; }

static size_t phase_work_bytes(int m, int n, int ns, long lda)
{
    const int mn = m < n ? m : n, len_beta = (m >= n) ? n - 1 : m;
    size_t w = svdgpu_bidiag_workspace(m, n, lda);
    w = maxz(w, svdgpu_ddc_workspace(mn));
    if (ns > 0) {
        w = maxz(w, svdgpu_twisted_workspace(mn, len_beta + 1, ns));
        w = maxz(w, svdgpu_backtransform_workspace(m, n_left(m, n), ns));
        w = maxz(w, svdgpu_backtransform_workspace(n, n_right(m, n), ns));
    }
    return up256(w);
}

/* QR first: worthwhile once the m x n bidiagonalization (BLAS2, ~4 m n^2) costs well more than a
 * tensor-core QR (2 m n^2 on the DMMA GEMM) plus the n x n problem */
static int use_qr_first(int m, int n)
{
    return g.qr_first && n >= 2 && (long)m * 10 >= (long)n * g.qr_ratio10 && m > n;
}
static size_t qr_r_bytes(int n) { return up256(sizeof(double) * (size_t)((n + 1) / 2 * 2) * n); }
static size_t qr_extra_bytes(int m, int n, int ns)
{
    if (!use_qr_first(m, n)) return 0;
    const long ldr = (n + 1) / 2 * 2;
    return qr_r_bytes(n) + maxz(up256(svdgpu_qr_workspace(m, n)),
                                maxz(phase_work_bytes(n, n, ns, ldr), phase_work_bytes(m, n, ns, (m + 1) / 2 * 2)));
}

static void vectors_core(int m, int n, const double *dA, long lda, const double *dalpha, const double *dbeta,
                         const double *dsig_all, int i0, int ns, double *dU, long ldu, double *dV, long ldv,
                         double *dsig_out, void *work, void *stream, void *ev_mid)
{
    const int mn = m < n ? m : n, len_beta = (m >= n) ? n - 1 : m, mb = len_beta + 1;
    svdgpu_memset(dU, 0, sizeof(double) * (size_t)ldu * ns, stream);
    svdgpu_memset(dV, 0, sizeof(double) * (size_t)ldv * ns, stream);
    /* x_i straight into V(:,i) (top mb entries), y_i = B x_i / sigma_i straight into U(:,i) */
    svdgpu_twisted_vectors(mn, mb, dalpha, dbeta, dsig_all, mn, i0, ns, dV, ldv, dU, ldu, dsig_out,
                           g.rqi, work, stream);
    if (ev_mid) svdgpu_event_record(ev_mid, stream);
    svdgpu_wy_apply(1, m, n_left(m, n), dA, lda, dU, ldu, ns, work, stream);
    svdgpu_event_record(g.ev_first, stream); g.first_is_u = 1; g.first_recorded = 1;
    svdgpu_wy_apply(0, n, n_right(m, n), dA, lda, dV, ldv, ns, work, stream);
}

void svd_gpu_values_dev(int m, int n, double *dA, long lda, double *dalpha, double *dbeta, double *dsigma,
                        void *stream)
{
    ctx_init();
    const int mn = m < n ? m : n;
    char *work = arena_get(phase_work_bytes(m, n, 0, lda));
    svdgpu_memset(dbeta, 0, sizeof(double) * (size_t)mn, stream);     /* beta[mn-1] = 0 when B is square */
    svdgpu_bidiag(m, n, dA, lda, dalpha, dbeta, work, g.nb, stream);
    svdgpu_ddc_values(mn, dalpha, dbeta, dsigma, work, stream);
}

void svd_gpu_vectors_dev(int m, int n, const double *dA_mod, long lda, const double *dalpha,
                         const double *dbeta, const double *dsigma_all, int i0, int ns, double *dUblk,
                         long ldu, double *dVblk, long ldv, double *dsig_out, void *stream)
{
    ctx_init();
    if (ns <= 0) return;
    char *work = arena_get(phase_work_bytes(m, n, ns, lda));
    vectors_core(m, n, dA_mod, lda, dalpha, dbeta, dsigma_all, i0, ns, dUblk, ldu, dVblk, ldv, dsig_out, work,
                 stream, NULL);
}

static void svd_dev_direct(int m, int n, double *dA, long lda, double *dsigma, double *dU, long ldu, double *dV,
                           long ldv, char *scratch, void *stream, void *ev_after_bidiag_for_copy, int record_start)
{
    const int mn = m < n ? m : n;
    const int want_vec = (dU != NULL && dV != NULL);
    double *dalpha = (double *)scratch;             scratch += up256(sizeof(double) * (size_t)mn);
    double *dbeta = (double *)scratch;              scratch += up256(sizeof(double) * ((size_t)mn + 1));
    double *dsig = (double *)scratch;               scratch += up256(sizeof(double) * (size_t)mn);
    double *dscale = (double *)scratch;             scratch += 256;
    void *work = scratch;

    if (record_start) svdgpu_event_record(g.ev[0], stream);
    /* range guard: exact power-of-two scaling when max|A| is far from 1 (sigma is scaled back below) */
    svdgpu_scale_matrix(m, n, dA, lda, dscale, (double *)work, stream);
    svdgpu_memset(dbeta, 0, sizeof(double) * ((size_t)mn + 1), stream);
    if (use_qr_first(m, n)) {
        /* A = Q R, then the square problem on R; dA keeps Q's reflectors (and R above the diagonal),
         * not bidiagonalization reflectors */
        const long ldr = (n + 1) / 2 * 2;
        double *dR = (double *)work;
        work = (char *)work + qr_r_bytes(n);
        svdgpu_qr(m, n, dA, lda, dR, ldr, work, stream);
        svdgpu_bidiag(n, n, dR, ldr, dalpha, dbeta, work, g.nb, stream);
        svdgpu_event_record(g.ev[1], stream);
        if (ev_after_bidiag_for_copy) svdgpu_event_record(ev_after_bidiag_for_copy, stream);
        svdgpu_ddc_values(n, dalpha, dbeta, want_vec ? dsig : dsigma, work, stream);
        svdgpu_event_record(g.ev[2], stream);
        if (want_vec) {
            svdgpu_memset(dU, 0, sizeof(double) * (size_t)ldu * n, stream);
            svdgpu_memset(dV, 0, sizeof(double) * (size_t)ldv * n, stream);
            svdgpu_twisted_vectors(n, n, dalpha, dbeta, dsig, n, 0, n, dV, ldv, dU, ldu, dsigma, g.rqi, work, stream);
            svdgpu_event_record(g.ev[3], stream);
            svdgpu_wy_apply(1, n, n_left(n, n), dR, ldr, dU, ldu, n, work, stream);
            svdgpu_wy_apply(0, n, n_right(n, n), dR, ldr, dV, ldv, n, work, stream);
            svdgpu_event_record(g.ev_first, stream); g.first_is_u = 0; g.first_recorded = 1;
            svdgpu_wy_apply(1, m, n, dA, lda, dU, ldu, n, work, stream);      /* U = Q [U_R; 0] */
        } else {
            svdgpu_event_record(g.ev[3], stream);
        }
        svdgpu_scale_vector(mn, dsigma, dscale + 1, stream);
        svdgpu_event_record(g.ev[4], stream);
        g.ms_pending = 1;
        return;
    }
    svdgpu_bidiag(m, n, dA, lda, dalpha, dbeta, work, g.nb, stream);
    svdgpu_event_record(g.ev[1], stream);
    if (ev_after_bidiag_for_copy) svdgpu_event_record(ev_after_bidiag_for_copy, stream);
    svdgpu_ddc_values(mn, dalpha, dbeta, want_vec ? dsig : dsigma, work, stream);
    svdgpu_event_record(g.ev[2], stream);
    if (want_vec) {
        vectors_core(m, n, dA, lda, dalpha, dbeta, dsig, 0, mn, dU, ldu, dV, ldv, dsigma, work, stream, g.ev[3]);
    } else {
        svdgpu_event_record(g.ev[3], stream);
    }
    svdgpu_scale_vector(mn, dsigma, dscale + 1, stream);
    svdgpu_event_record(g.ev[4], stream);
    g.ms_pending = 1;
}

static size_t small_bytes(int mn)
{
    return 2 * up256(sizeof(double) * (size_t)mn) + up256(sizeof(double) * ((size_t)mn + 1)) + 256;
}
static size_t direct_scratch_bytes(int m, int n, int ns, long lda)
{
    return small_bytes(m < n ? m : n) + maxz(phase_work_bytes(m, n, ns, lda), qr_extra_bytes(m, n, ns));
}
/* wide inputs: A^T = V S U^T is a tall problem, which gets the left vectors from their own twisted
 * factorization (and the QR-first route when n >> m) */
static int use_wide_transpose(int m, int n) { return g.wide_transpose && m < n; }
static size_t scratch_bytes(int m, int n, int ns, long lda)
{
    if (!use_wide_transpose(m, n)) return direct_scratch_bytes(m, n, ns, lda);
    const long ldt = (n + 1) / 2 * 2;
    return up256(sizeof(double) * (size_t)ldt * m) + direct_scratch_bytes(n, m, ns, ldt);
}

static void svd_dev_inner(int m, int n, double *dA, long lda, double *dsigma, double *dU, long ldu, double *dV,
                          long ldv, char *scratch, void *stream, void *ev_after_bidiag_for_copy)
{
    if ((dU == NULL) != (dV == NULL)) {
        fprintf(stderr, "svd_gpu: U and V must both be given or both be NULL (values only)\n");
        abort();
    }
    if (!use_wide_transpose(m, n)) {
        svd_dev_direct(m, n, dA, lda, dsigma, dU, ldu, dV, ldv, scratch, stream, ev_after_bidiag_for_copy, 1);
        return;
    }
    const long ldt = (n + 1) / 2 * 2;
    double *dAt = (double *)scratch;
    scratch += up256(sizeof(double) * (size_t)ldt * m);
    svdgpu_event_record(g.ev[0], stream);
    if (ldt != n) svdgpu_memset(dAt, 0, sizeof(double) * (size_t)ldt * m, stream);
    svdgpu_transpose(m, n, dA, lda, dAt, ldt, stream);
    svd_dev_direct(n, m, dAt, ldt, dsigma, dV, ldv, dU, ldu, scratch, stream, NULL, 0);
    g.first_is_u = !g.first_is_u;
    /* A leaves as the transpose of the tall problem's reflector storage */
    svdgpu_transpose(n, m, dAt, ldt, dA, lda, stream);
    if (ev_after_bidiag_for_copy) svdgpu_event_record(ev_after_bidiag_for_copy, stream);
}

void svd_gpu_dev(int m, int n, double *dA, long lda, double *dsigma, double *dU, long ldu, double *dV, long ldv,
                 void *stream)
{
    ctx_init();
    const int mn = m < n ? m : n;
    const int ns = (dU && dV) ? mn : 0;
    char *scratch = arena_get(scratch_bytes(m, n, ns, lda));
    g.ms[0] = g.ms[5] = 0.f;
    svd_dev_inner(m, n, dA, lda, dsigma, dU, ldu, dV, ldv, scratch, stream, NULL);
}

void svd_gpu_last_phase_ms(float ms[7])
{
    if (g.ms_pending) {
        g.ms[1] = svdgpu_event_elapsed_ms(g.ev[0], g.ev[1]);
        g.ms[2] = svdgpu_event_elapsed_ms(g.ev[1], g.ev[2]);
        g.ms[3] = svdgpu_event_elapsed_ms(g.ev[2], g.ev[3]);
        g.ms[4] = svdgpu_event_elapsed_ms(g.ev[3], g.ev[4]);
        if (g.ms[0] == 0.f && g.ms[5] == 0.f) g.ms[6] = g.ms[1] + g.ms[2] + g.ms[3] + g.ms[4];
        g.ms_pending = 0;
    }
    memcpy(ms, g.ms, sizeof g.ms);
}

/* ---- the drop-in entry point (svd_gpu.h:5, svd_gpu.c:53) ---------------------------- */
void svd_gpu(int m, int n, double *A, double *sigma, double *U, double *V)
{
    if (m <= 0 || n <= 0 || !A || !sigma) {
        fprintf(stderr, "svd_gpu: bad arguments (m=%d n=%d)\n", m, n);
        abort();
    }
    ctx_init();
    const double t_start = wall_ms();
    const int mn = m < n ? m : n;
    const int want_vec = (U != NULL && V != NULL);
    const long lda = (m + 1) / 2 * 2;
    const size_t bytesA = up256(sizeof(double) * (size_t)lda * n);
    const size_t bytesU = want_vec ? up256(sizeof(double) * (size_t)m * mn) : 0;
    const size_t bytesV = want_vec ? up256(sizeof(double) * (size_t)n * mn) : 0;
    char *base = arena_get(bytesA + bytesU + bytesV + up256(sizeof(double) * (size_t)mn) +
                           scratch_bytes(m, n, want_vec ? mn : 0, lda));
    double *dA = (double *)base;
    double *dU = want_vec ? (double *)(base + bytesA) : NULL;
    double *dV = want_vec ? (double *)(base + bytesA + bytesU) : NULL;
    double *dsig_final = (double *)(base + bytesA + bytesU + bytesV);
    char *scratch = base + bytesA + bytesU + bytesV + up256(sizeof(double) * (size_t)mn);

    /* host -> device (the reference's blocking clEnqueueWriteBuffer, bidiag_par.c:298-301) */
    const double t_h2d0 = wall_ms();
    if (lda == m) {
        svdgpu_h2d(dA, A, sizeof(double) * (size_t)m * n, g.stream);
    } else {
        svdgpu_memset(dA, 0, sizeof(double) * (size_t)lda * n, g.stream);
        svdgpu_h2d_2d(dA, sizeof(double) * lda, A, sizeof(double) * m, sizeof(double) * m, n, g.stream);
    }
    svdgpu_stream_sync(g.stream);
    g.ms[0] = (float)(wall_ms() - t_h2d0);

    g.first_recorded = 0;
    svd_dev_inner(m, n, dA, lda, dsig_final, dU, m, dV, n, scratch, g.stream, g.ev_copy);

    /* device -> host.  The reflector matrix goes back on the copy stream as soon as the
     * bidiagonalization is done (nothing modifies dA afterwards), overlapping the later phases. */
    svdgpu_stream_wait_event(g.copy_stream, g.ev_copy);
    if (lda == m) svdgpu_d2h(A, dA, sizeof(double) * (size_t)m * n, g.copy_stream);
    else svdgpu_d2h_2d(A, sizeof(double) * m, dA, sizeof(double) * lda, sizeof(double) * m, n, g.copy_stream);
    /* ... and so does whichever of U / V is final first, while the other is still being back-transformed.
     * Only the first min(m,n) columns are written (svd_gpu.c:118-121). */
    int early = -1;                                   /* 1: U went on the copy stream, 0: V */
    if (want_vec && g.first_recorded) {
        early = g.first_is_u;
        svdgpu_stream_wait_event(g.copy_stream, g.ev_first);
        if (early) svdgpu_d2h(U, dU, sizeof(double) * (size_t)m * mn, g.copy_stream);
        else svdgpu_d2h(V, dV, sizeof(double) * (size_t)n * mn, g.copy_stream);
    }
    svdgpu_stream_sync(g.stream);
    const double t_d2h0 = wall_ms();
    svdgpu_d2h(sigma, dsig_final, sizeof(double) * (size_t)mn, g.stream);
    if (want_vec) {
        if (early != 1) svdgpu_d2h(U, dU, sizeof(double) * (size_t)m * mn, g.stream);
        if (early != 0) svdgpu_d2h(V, dV, sizeof(double) * (size_t)n * mn, g.stream);
    }
    svdgpu_stream_sync(g.stream);
    svdgpu_stream_sync(g.copy_stream);
    g.ms[5] = (float)(wall_ms() - t_d2h0);
    g.ms[6] = (float)(wall_ms() - t_start);
}
