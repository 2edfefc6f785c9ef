/* svd_gpu.c — the C host orchestrator behind the drop-in entry point svd_gpu().
 *
 * Same phase sequence as the reference's svd_gpu.c:100-121
 *     bidiag_par -> (transpose) -> GetSingularValues_Parallel -> CalcRightSingularVectors
 *     -> RighttoLeftSingularVectors -> multU / multV per vector
 * but every phase runs on the GPU through the C-ABI layer (include/cuda-helper.h); the matrix
 * crosses PCIe once in each direction, the host transpose disappears (the row reflectors are
 * gathered on the device) and the n x n intermediates X, Y are written straight into the device
 * images of V and U.  Host code is plain C, as in the reference.
 *
 * One code path serves 1 and N GPUs.  The reference's shardable loop is the
 * "omp parallel for" over singular vectors, svd_gpu.c:117-121; here a GROUP of ranks (one process
 * driving N devices, or one process per device) splits the singular values into contiguous blocks:
 *     rank 0      bidiagonalization (+ QR first for tall inputs) and dDC singular values;
 *                 while the factorization runs, the reflectors that are already final are turned
 *                 into compact-WY panels (V, V T) on a low-priority stream and broadcast, chunk by
 *                 chunk, over NCCL — the other ranks never see the reflector matrix, only panels
 *                 ready to apply, and the transfer hides behind the factorization;
 *     every rank  twisted vectors and back-transform of its own block of singular values;
 *     exchanges   broadcast of panels, broadcast of alpha | beta | sigma ("all-gather the
 *                 bidiagonal"), all-gather of the polished singular values.  U / V blocks stay
 *                 with their ranks (svd_gpu() copies each block to the host from its own GPU).
 * With one rank the same code runs without NCCL; the panel set-up then overlaps the dDC and
 * twisted phases instead of the factorization.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../../include/svd_gpu_b200.h"
#include "../../include/cuda-helper.h"

#define SVD_MAX_DEV 16
#define CHUNK_PANELS 4            /* compact-WY panels per broadcast chunk */

/* ---- per-device context: streams, events, cached arena.  One lock per device: concurrent calls on
 * different devices run in parallel, calls on the same device are serialised. ----------------- */
typedef struct {
    int inited, dev;
    void *s_main, *s_side, *s_comm, *s_copy;
    void *ev[8];                  /* 0 start, 1 factorization, 2 dDC, 3 twisted, 4 back-transform, 5 small arrived */
    void *ev_in, *ev_prog, *ev_ready, *ev_panels, *ev_first, *ev_done, *ev_commdone, *ev_copy;
    void *arena;
    size_t arena_bytes;
    size_t io_reserved;           /* bytes at the start of the arena that belong to the host-pointer wrapper */
    float ms[8];
    int ms_pending;
    int first_is_u, first_recorded;
    int was_root;                 /* the last call ran the factorization on this device */
    pthread_mutex_t mu;
} svd_ctx;

typedef struct {
    int inited;
    int nb, rqi, qr_first, qr_ratio10, wide_transpose, wy_overlap, host_register, ngpus;
} svd_opts;

static svd_ctx g_ctx[SVD_MAX_DEV];
static svd_opts g_opt;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static int g_last_dev = 0;        /* device whose context holds the timings of the last call */

static size_t up256(size_t b) { return (b + 255) / 256 * 256; }
static size_t maxz(size_t a, size_t b) { return a > b ? a : b; }
static double wall_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}
static int env_int(const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; }

static void init_once(void)
{
    for (int d = 0; d < SVD_MAX_DEV; ++d) pthread_mutex_init(&g_ctx[d].mu, NULL);
}
static void opts_init(void)
{
    pthread_once(&g_once, init_once);
    pthread_mutex_lock(&g_mu);
    if (!g_opt.inited) {
        g_opt.nb = env_int("SVD_GPU_NB", 32);
        g_opt.rqi = env_int("SVD_GPU_RQI", 1);
        g_opt.qr_first = env_int("SVD_GPU_QR_FIRST", 1);
        g_opt.qr_ratio10 = env_int("SVD_GPU_QR_RATIO10", 25);
        g_opt.wide_transpose = env_int("SVD_GPU_WIDE_TRANSPOSE", 1);
        g_opt.wy_overlap = env_int("SVD_GPU_WY_OVERLAP", 0);
        g_opt.host_register = env_int("SVD_GPU_HOST_REGISTER", 0);
        g_opt.ngpus = env_int("SVD_GPU_NGPUS", 1);
        g_opt.inited = 1;
    }
    pthread_mutex_unlock(&g_mu);
}

/* lock and return the context of device `dev` (made current) */
static void device_init(void)
{
    static int done = 0;
    pthread_mutex_lock(&g_mu);
    if (!done) {
        (void)svdgpu_device_count();                 /* aborts loudly when there is no GPU: there is no CPU fallback */
        const char *e = getenv("SVD_GPU_DEVICE");
        if (e) svdgpu_set_device(atoi(e));
        done = 1;
    }
    pthread_mutex_unlock(&g_mu);
}

static svd_ctx *ctx_acquire(int dev)
{
    opts_init();
    if (dev < 0 || dev >= SVD_MAX_DEV) { fprintf(stderr, "svd_gpu: device %d out of range\n", dev); abort(); }
    svd_ctx *c = &g_ctx[dev];
    pthread_mutex_lock(&c->mu);
    svdgpu_set_device(dev);
    if (!c->inited) {
        c->dev = dev;
        c->s_main = svdgpu_stream_create_priority(1);
        c->s_side = svdgpu_stream_create_priority(0);
        c->s_comm = svdgpu_stream_create_priority(0);
        c->s_copy = svdgpu_stream_create();
        for (int i = 0; i < 8; ++i) c->ev[i] = svdgpu_event_create();
        c->ev_in = svdgpu_event_create();      c->ev_prog = svdgpu_event_create();
        c->ev_ready = svdgpu_event_create();   c->ev_panels = svdgpu_event_create();
        c->ev_first = svdgpu_event_create();   c->ev_done = svdgpu_event_create();
        c->ev_commdone = svdgpu_event_create(); c->ev_copy = svdgpu_event_create();
        c->inited = 1;
    }
    return c;
}
static void ctx_release(svd_ctx *c) { pthread_mutex_unlock(&c->mu); }

static char *arena_get(svd_ctx *c, size_t bytes)
{
    if (bytes > c->arena_bytes) {
        if (c->arena) { svdgpu_device_sync(); svdgpu_free(c->arena); }
        c->arena = svdgpu_malloc(bytes);
        c->arena_bytes = bytes;
    }
    return (char *)c->arena;
}

void svd_gpu_set_option(const char *name, int value)
{
    opts_init();
    pthread_mutex_lock(&g_mu);
    if (!strcmp(name, "nb")) g_opt.nb = value;
    else if (!strcmp(name, "rqi")) g_opt.rqi = value;
    else if (!strcmp(name, "qr_first")) g_opt.qr_first = value;
    else if (!strcmp(name, "qr_ratio10")) g_opt.qr_ratio10 = value;
    else if (!strcmp(name, "wide_transpose")) g_opt.wide_transpose = value;
    else if (!strcmp(name, "wy_overlap")) g_opt.wy_overlap = value;
    else if (!strcmp(name, "host_register")) g_opt.host_register = value;
    else if (!strcmp(name, "ngpus")) g_opt.ngpus = value;
    else if (!strcmp(name, "release")) {          /* drop the cached device arenas */
        pthread_mutex_unlock(&g_mu);
        const int cur = svdgpu_get_device();
        for (int d = 0; d < SVD_MAX_DEV; ++d) {
            svd_ctx *c = &g_ctx[d];
            pthread_mutex_lock(&c->mu);
            if (c->arena) {
                svdgpu_set_device(d); svdgpu_device_sync(); svdgpu_free(c->arena);
                c->arena = NULL; c->arena_bytes = 0;
            }
            pthread_mutex_unlock(&c->mu);
        }
        svdgpu_set_device(cur);
        return;
    } else { fprintf(stderr, "svd_gpu_set_option: unknown option '%s'\n", name); abort(); }
    pthread_mutex_unlock(&g_mu);
}

/* number of Householder reflectors bidiag leaves on each side (bidiag_par.c:1046-1060, :1014-1023):
 * the extra ones counted here are zero vectors (H = I), harmless to include */
static int n_left(int m, int n) { return m < n ? m : n; }
static int n_right(int m, int n) { return m >= n ? (n >= 2 ? n - 2 : 0) : m; }

/* QR first: worthwhile once the m x n bidiagonalization (BLAS2, ~4 m n^2) costs well more than a
 * tensor-core QR (2 m n^2 on the DMMA GEMM) plus the n x n problem */
static int use_qr_first(int m, int n)
{
    return g_opt.qr_first && n >= 2 && (long)m * 10 >= (long)n * g_opt.qr_ratio10 && m > n;
}
/* wide inputs: A^T = V S U^T is a tall problem, which gets the left vectors from their own twisted
 * factorization (and the QR-first route when n >> m) */
static int use_wide_transpose(int m, int n) { return g_opt.wide_transpose && m < n; }

void svdgpu_shard_range(int mn, int world, int rank, int *blk, int *i0, int *ns)
{
    const int b = (mn + world - 1) / world;
    int s = rank * b;
    if (s > mn) s = mn;
    int c = mn - s;
    if (c > b) c = b;
    if (c < 0) c = 0;
    if (blk) *blk = b;
    if (i0) *i0 = s;
    if (ns) *ns = c;
}

/* =================================================================================================
 * groups
 * ================================================================================================= */
struct svdgpu_group {
    int world, nlocal, rank0;     /* ranks in total, ranks driven by this process, global rank of local rank 0 */
    int dev[SVD_MAX_DEV];
    void *comm[SVD_MAX_DEV];
};

svdgpu_group *svdgpu_group_create_local(int ndev, const int *devices)
{
    opts_init();
    device_init();
    if (ndev < 1 || ndev > SVD_MAX_DEV || ndev > svdgpu_device_count()) {
        fprintf(stderr, "svdgpu_group_create_local: %d devices requested, %d visible\n", ndev, svdgpu_device_count());
        abort();
    }
    svdgpu_group *g = (svdgpu_group *)calloc(1, sizeof *g);
    if (!g) abort();
    g->world = g->nlocal = ndev; g->rank0 = 0;
    for (int i = 0; i < ndev; ++i) g->dev[i] = devices ? devices[i] : i;
    if (ndev > 1) {
        const int cur = svdgpu_get_device();
        svdgpu_nccl_comm_init_all(ndev, g->dev, g->comm);
        svdgpu_set_device(cur);
    }
    return g;
}
svdgpu_group *svdgpu_group_create_rank(int nranks, int rank, const void *id128)
{
    opts_init();
    device_init();
    svdgpu_group *g = (svdgpu_group *)calloc(1, sizeof *g);
    if (!g) abort();
    g->world = nranks; g->nlocal = 1; g->rank0 = rank;
    g->dev[0] = svdgpu_get_device();
    if (nranks > 1) g->comm[0] = svdgpu_nccl_comm_init_rank(nranks, rank, id128);
    return g;
}
void svdgpu_group_destroy(svdgpu_group *g)
{
    if (!g) return;
    const int cur = svdgpu_get_device();
    for (int i = 0; i < g->nlocal; ++i)
        if (g->comm[i]) { svdgpu_set_device(g->dev[i]); svdgpu_nccl_comm_destroy(g->comm[i]); }
    svdgpu_set_device(cur);
    free(g);
}
int svdgpu_group_size(const svdgpu_group *g) { return g->world; }
int svdgpu_group_nlocal(const svdgpu_group *g) { return g->nlocal; }
int svdgpu_group_rank(const svdgpu_group *g, int local) { return g->rank0 + local; }
int svdgpu_group_device(const svdgpu_group *g, int local) { return g->dev[local]; }

/* =================================================================================================
 * one call
 * ================================================================================================= */
typedef struct {                  /* one reflector set, prepared and shipped as compact-WY panels */
    int active, left, rows, nref, np, nchunks;
    const double *A;              /* reflector storage on the root */
    long lda;
    size_t off;                   /* offset of the panel storage in every rank's arena */
    int stage;                    /* 0: produced by the QR, 1: by the bidiagonalization */
} refl_set;

typedef struct {
    svdgpu_group *grp;
    svd_ctx *cx[SVD_MAX_DEV];     /* context of every local rank */
    char *base[SVD_MAX_DEV];      /* arena of every local rank */
    int m, n, mn, mb, world, blk, qr;
    int has_root;                 /* local rank 0 is global rank 0 */
    refl_set set[3];              /* [0] Q of the QR (QR first only), [1] left, [2] right reflectors of the bidiagonalization */
    int order[3 * 4096][2];       /* canonical chunk order: (set, chunk) */
    int norder, issued;           /* chunks issued so far (root: prepared and broadcast) */
    int progressive;              /* prepare panels while the factorization runs */
    /* arena offsets */
    size_t off_small, off_sigblk, off_siggather, off_tw, off_apply, off_fact, off_R, total_root, total_peer;
} svd_call;

static int chunk_count(int np) { return (np + CHUNK_PANELS - 1) / CHUNK_PANELS; }

static void plan_sets(svd_call *c, const double *dA, long lda)
{
    const int m = c->m, n = c->n;
    memset(c->set, 0, sizeof c->set);
    if (c->qr) {
        /* A = Q R: Q's reflectors in dA (m x n), the bidiagonalization runs on R (n x n, in the arena) */
        c->set[0].active = 1; c->set[0].left = 1; c->set[0].rows = m; c->set[0].nref = n; c->set[0].A = dA;
        c->set[0].lda = lda; c->set[0].stage = 0;
        c->set[1].active = 1; c->set[1].left = 1; c->set[1].rows = n; c->set[1].nref = n_left(n, n); c->set[1].stage = 1;
        c->set[2].active = n_right(n, n) > 0; c->set[2].left = 0; c->set[2].rows = n; c->set[2].nref = n_right(n, n);
        c->set[2].stage = 1;
    } else {
        c->set[1].active = 1; c->set[1].left = 1; c->set[1].rows = m; c->set[1].nref = n_left(m, n); c->set[1].A = dA;
        c->set[1].lda = lda; c->set[1].stage = 1;
        c->set[2].active = n_right(m, n) > 0; c->set[2].left = 0; c->set[2].rows = n; c->set[2].nref = n_right(m, n);
        c->set[2].A = dA; c->set[2].lda = lda; c->set[2].stage = 1;
    }
    c->norder = 0;
    for (int s = 0; s < 3; ++s) {
        if (!c->set[s].active) continue;
        c->set[s].np = svdgpu_wy_panel_count(c->set[s].nref);
        c->set[s].nchunks = chunk_count(c->set[s].np);
    }
    /* canonical order: the QR's chunks, then the bidiagonalization's by (chunk, side) — the order in which
     * they become final, whatever the granularity of the progress callbacks */
    for (int ch = 0; c->set[0].active && ch < c->set[0].nchunks; ++ch) { c->order[c->norder][0] = 0; c->order[c->norder++][1] = ch; }
    const int nc1 = c->set[1].nchunks, nc2 = c->set[2].active ? c->set[2].nchunks : 0;
    for (int ch = 0; ch < (nc1 > nc2 ? nc1 : nc2); ++ch) {
        if (ch < nc1) { c->order[c->norder][0] = 1; c->order[c->norder++][1] = ch; }
        if (ch < nc2) { c->order[c->norder][0] = 2; c->order[c->norder++][1] = ch; }
    }
    if (c->norder > 3 * 4096) { fprintf(stderr, "svd_gpu: too many panel chunks\n"); abort(); }
}

static void plan_arena(svd_call *c, long lda)
{
    const int m = c->m, n = c->n, mn = c->mn;
    size_t o = 0;
    c->off_small = o;      o += up256(sizeof(double) * (3 * (size_t)mn + 16));
    c->off_sigblk = o;     o += up256(sizeof(double) * ((size_t)c->blk + 8));
    c->off_siggather = o;  o += up256(sizeof(double) * ((size_t)c->blk * c->world + 8));
    for (int s = 0; s < 3; ++s)
        if (c->set[s].active) { c->set[s].off = o; o += up256(svdgpu_wy_panels_bytes(c->set[s].rows, c->set[s].nref)); }
    c->off_tw = o;         o += up256(svdgpu_twisted_workspace(mn, c->mb, c->blk));
    c->off_apply = o;      o += up256(svdgpu_wy_apply_workspace(c->blk));
    c->total_peer = o;
    c->off_R = o;
    if (c->qr) o += up256(sizeof(double) * (size_t)((n + 1) / 2 * 2) * n);
    c->off_fact = o;
    {
        size_t w = svdgpu_ddc_workspace(mn);
        if (c->qr) {
            w = maxz(w, svdgpu_qr_workspace(m, n));
            w = maxz(w, svdgpu_bidiag_workspace(n, n, (n + 1) / 2 * 2));
        } else {
            w = maxz(w, svdgpu_bidiag_workspace(m, n, lda));
        }
        o += up256(w);
    }
    c->total_root = o;
}

/* Pure planning (no device): the route svd_gpu() takes for an m x n input and the canonical order in which the
 * compact-WY panel chunks are prepared on rank 0 and broadcast — every rank derives the same list from (m, n)
 * alone, which is what lets the other ranks post their receives before rank 0 has produced anything.
 * route[0] = 1 if the input is solved as the SVD of its transpose, route[1] = 1 for QR first, route[2..3] = the
 * (rows, cols) the core factorizes.  Per chunk: set (0 Q of the QR, 1 left, 2 right reflectors), first panel,
 * end panel, and the number of reflectors of the producing factorization that must be final.  Returns the count. */
int svdgpu_plan_chunks(int m, int n, int world, int route[4], int max_out, int *set_out, int *pb_out, int *pe_out,
                       int *need_out)
{
    opts_init();
    svdgpu_group g;
    memset(&g, 0, sizeof g);
    g.world = world < 1 ? 1 : world; g.nlocal = 1;
    svd_call *c = (svd_call *)malloc(sizeof *c);
    if (!c) abort();
    const int wide = use_wide_transpose(m, n);
    const int tm = wide ? n : m, tn = wide ? m : n;
    memset(c, 0, sizeof *c);
    c->grp = &g; c->m = tm; c->n = tn; c->mn = tm < tn ? tm : tn; c->world = g.world;
    c->qr = use_qr_first(tm, tn);
    plan_sets(c, NULL, 0);
    if (route) { route[0] = wide; route[1] = c->qr; route[2] = tm; route[3] = tn; }
    const int nbw = 128;
    int cnt = 0;
    for (int i = 0; i < c->norder; ++i, ++cnt) {
        if (cnt >= max_out) continue;
        const refl_set *s = &c->set[c->order[i][0]];
        const int ch = c->order[i][1];
        const int pb = ch * CHUNK_PANELS, pe = (pb + CHUNK_PANELS < s->np) ? pb + CHUNK_PANELS : s->np;
        set_out[cnt] = c->order[i][0]; pb_out[cnt] = pb; pe_out[cnt] = pe;
        need_out[cnt] = pe * nbw < s->nref ? pe * nbw : s->nref;
    }
    free(c);
    return cnt;
}

/* ---- broadcast of one chunk of prepared panels over the local ranks (V, then V T) ----------------- */
static void bcast_chunk(svd_call *c, int idx)
{
    const refl_set *s = &c->set[c->order[idx][0]];
    const int ch = c->order[idx][1];
    const int pb = ch * CHUNK_PANELS, pe = (pb + CHUNK_PANELS < s->np) ? pb + CHUNK_PANELS : s->np;
    svdgpu_group *g = c->grp;
    for (int pass = 0; pass < 2; ++pass) {
        svdgpu_nccl_group_start();
        for (int lr = 0; lr < g->nlocal; ++lr) {
            double *V, *VT; size_t cnt;
            svdgpu_set_device(c->cx[lr]->dev);
            svdgpu_wy_panel_slices(c->base[lr] + s->off, s->rows, s->nref, pb, pe, &V, &VT, &cnt);
            svdgpu_nccl_bcast(g->comm[lr], pass ? VT : V, cnt, 0, c->cx[lr]->s_comm);
        }
        svdgpu_nccl_group_end();
    }
}

/* progress callback of the factorizations (rank 0 only): reflectors [0, done) of `stage` are final.
 * Everything that became final since the last call is prepared in ONE batched set-up per reflector set on
 * the low-priority stream (short, wide kernels: they slip between the factorization's launches instead of
 * holding SMs for long), then broadcast chunk by chunk in the canonical order. */
typedef struct { svd_call *c; int stage; } prog_user;
static void on_progress(void *user, int done, void *stream)
{
    prog_user *u = (prog_user *)user;
    svd_call *c = u->c;
    svd_ctx *r = c->cx[0];
    const int nbw = svdgpu_wy_panel_width();
    int upto = c->issued;
    while (upto < c->norder) {
        const refl_set *s = &c->set[c->order[upto][0]];
        const int ch = c->order[upto][1];
        if (s->stage != u->stage) break;
        const int pe = (ch + 1) * CHUNK_PANELS < s->np ? (ch + 1) * CHUNK_PANELS : s->np;
        const int need = pe * nbw < s->nref ? pe * nbw : s->nref;
        const int final = done >= s->nref;
        if (done < need && !final) break;
        if (!c->progressive && !final) break;         /* one rank: everything at once, after the factorization */
        ++upto;
    }
    if (upto == c->issued) return;
    svdgpu_set_device(r->dev);
    /* wy_overlap = 2 (one rank only): no side stream, the set-up simply follows the factorization */
    void *s_setup = (c->world == 1 && g_opt.wy_overlap == 2) ? r->s_main : r->s_side;
    svdgpu_event_record(r->ev_prog, stream);
    svdgpu_stream_wait_event(s_setup, r->ev_prog);
    for (int q = 0; q < 3; ++q) {
        int pb = -1, pe = -1;
        for (int i = c->issued; i < upto; ++i) {
            if (c->order[i][0] != q) continue;
            const int ch = c->order[i][1];
            if (pb < 0) pb = ch * CHUNK_PANELS;
            pe = (ch + 1) * CHUNK_PANELS < c->set[q].np ? (ch + 1) * CHUNK_PANELS : c->set[q].np;
        }
        if (pb >= 0)
            svdgpu_wy_setup(c->set[q].left, c->set[q].rows, c->set[q].nref, c->set[q].A, c->set[q].lda,
                            c->base[0] + c->set[q].off, pb, pe, s_setup);
    }
    svdgpu_event_record(r->ev_ready, s_setup);
    if (c->world > 1) {
        svdgpu_stream_wait_event(r->s_comm, r->ev_ready);
        for (int i = c->issued; i < upto; ++i) bcast_chunk(c, i);
        svdgpu_set_device(r->dev);
    }
    c->issued = upto;
}

/* twisted vectors + back-transform of one local rank's block, enqueued on that rank's streams */
typedef struct { svd_call *c; int lr; double *dU, *dV; long ldu, ldv; } rank_job;
static void *rank_vectors(void *arg)
{
    rank_job *j = (rank_job *)arg;
    svd_call *c = j->c;
    svdgpu_group *g = c->grp;
    const int lr = j->lr, mn = c->mn, world = c->world;
    svd_ctx *x = c->cx[lr];
    svdgpu_set_device(x->dev);
    int i0, ns;
    svdgpu_shard_range(mn, world, g->rank0 + lr, NULL, &i0, &ns);
    double *small = (double *)(c->base[lr] + c->off_small);
    double *dalpha = small, *dbeta = small + mn, *dsig = small + 2 * (size_t)mn + 1;
    double *sigblk = (double *)(c->base[lr] + c->off_sigblk);
    void *aws = c->base[lr] + c->off_apply;
    if (world > 1) {
        svdgpu_event_record(x->ev_commdone, x->s_comm);       /* panels and the bidiagonal have arrived */
        svdgpu_stream_wait_event(x->s_main, x->ev_commdone);
    }
    svdgpu_event_record(x->ev[5], x->s_main);
    svdgpu_range_push("svd_gpu:twisted");
    svdgpu_memset(sigblk, 0, sizeof(double) * (size_t)c->blk, x->s_main);
    if (ns > 0) {
        /* x_i straight into V(:,i) (top mb entries), y_i straight into U(:,i) */
        svdgpu_memset(j->dU, 0, sizeof(double) * (size_t)j->ldu * ns, x->s_main);
        svdgpu_memset(j->dV, 0, sizeof(double) * (size_t)j->ldv * ns, x->s_main);
        svdgpu_twisted_vectors(mn, c->mb, dalpha, dbeta, dsig, mn, i0, ns, j->dV, j->ldv, j->dU, j->ldu, sigblk,
                               g_opt.rqi, c->base[lr] + c->off_tw, x->s_main);
    }
    svdgpu_event_record(x->ev[3], x->s_main);
    svdgpu_range_pop();
    svdgpu_range_push("svd_gpu:back-transform");
    if (world == 1) svdgpu_stream_wait_event(x->s_main, x->ev_ready);    /* panel set-up (s_side) */
    if (ns > 0) {
        const refl_set *q = &c->set[0], *l = &c->set[1], *rr = &c->set[2];
        svdgpu_wy_apply_prepared(1, l->rows, l->nref, c->base[lr] + l->off, j->dU, j->ldu, ns, aws, x->s_main);
        if (!c->qr) { svdgpu_event_record(x->ev_first, x->s_main); x->first_is_u = 1; x->first_recorded = 1; }
        if (rr->active)
            svdgpu_wy_apply_prepared(0, rr->rows, rr->nref, c->base[lr] + rr->off, j->dV, j->ldv, ns, aws, x->s_main);
        if (c->qr) {
            svdgpu_event_record(x->ev_first, x->s_main); x->first_is_u = 0; x->first_recorded = 1;
            svdgpu_wy_apply_prepared(1, q->rows, q->nref, c->base[lr] + q->off, j->dU, j->ldu, ns, aws, x->s_main);   /* U = Q [U_R; 0] */
        }
    }
    svdgpu_range_pop();
    return NULL;
}

/* The tall / square / direct-wide core: dA (m x n, root only) -> sigma on every rank, blocks of U and V.
 * dsigma[lr], dU[lr] (m x blk, ldu), dV[lr] (n x blk, ldv): per local rank; values only when dU == NULL. */
static void svd_core(svd_call *c, double *dA, long lda, double *const *dsigma, double *const *dU, long ldu,
                     double *const *dV, long ldv, void *ev_after_fact)
{
    svdgpu_group *g = c->grp;
    const int m = c->m, n = c->n, mn = c->mn, world = c->world;
    const int want_vec = (dU != NULL);
    const long nsmall = 3 * (long)mn + 8;             /* alpha[mn] | beta[mn+1] | sigma[mn] | scale[2] (+pad) */
    /* ------------------------------------------------------------ rank 0: factorization + singular values */
    if (c->has_root) {
        svd_ctx *r = c->cx[0];
        svdgpu_set_device(r->dev);
        double *small = (double *)(c->base[0] + c->off_small);
        double *dalpha = small, *dbeta = small + mn, *dsig = small + 2 * (size_t)mn + 1, *dscale = dsig + mn;
        void *work = c->base[0] + c->off_fact;
        prog_user pu0 = {c, 0}, pu1 = {c, 1};
        svdgpu_progress p0 = {on_progress, &pu0, svdgpu_wy_panel_width()};
        svdgpu_progress p1 = {on_progress, &pu1, svdgpu_wy_panel_width()};
        svdgpu_range_push("svd_gpu:factorization");
        /* range guard: exact power-of-two scaling when max|A| is far from 1 (sigma is scaled back) */
        svdgpu_scale_matrix(m, n, dA, lda, dscale, (double *)work, r->s_main);
        svdgpu_memset(dbeta, 0, sizeof(double) * ((size_t)mn + 1), r->s_main);
        if (c->qr) {
            const long ldr = (n + 1) / 2 * 2;
            double *dR = (double *)(c->base[0] + c->off_R);
            c->set[1].A = dR; c->set[1].lda = ldr; c->set[2].A = dR; c->set[2].lda = ldr;
            svdgpu_qr_progress(m, n, dA, lda, dR, ldr, work, want_vec ? &p0 : NULL, r->s_main);
            svdgpu_bidiag_progress(n, n, dR, ldr, dalpha, dbeta, work, g_opt.nb, want_vec ? &p1 : NULL, r->s_main);
        } else {
            svdgpu_bidiag_progress(m, n, dA, lda, dalpha, dbeta, work, g_opt.nb, want_vec ? &p1 : NULL, r->s_main);
        }
        svdgpu_event_record(r->ev[1], r->s_main);
        if (ev_after_fact) svdgpu_event_record(ev_after_fact, r->s_main);
        svdgpu_range_pop();
        svdgpu_range_push("svd_gpu:dDC");
        svdgpu_ddc_values(mn, dalpha, dbeta, dsig, work, r->s_main);
        svdgpu_event_record(r->ev[2], r->s_main);
        svdgpu_range_pop();
        if (!want_vec) {
            svdgpu_d2d(dsigma[0], dsig, sizeof(double) * (size_t)mn, r->s_main);
            svdgpu_scale_vector(mn, dsigma[0], dscale + 1, r->s_main);
            for (int e = 3; e <= 5; ++e) svdgpu_event_record(r->ev[e], r->s_main);
            return;
        }
        if (world > 1) svdgpu_stream_wait_event(r->s_comm, r->ev[2]);
    } else if (want_vec) {
        /* a process without rank 0: post the receives of every chunk, in the canonical order */
        for (int i = 0; i < c->norder; ++i) bcast_chunk(c, i);
        c->issued = c->norder;
    }
    if (!want_vec) return;
    if (c->issued != c->norder) { fprintf(stderr, "svd_gpu: internal error, %d of %d panel chunks issued\n", c->issued, c->norder); abort(); }
    /* ------------------------------------------------------------ alpha | beta | sigma | scale to everyone */
    if (world > 1) {
        svdgpu_nccl_group_start();
        for (int lr = 0; lr < g->nlocal; ++lr) {
            svdgpu_set_device(c->cx[lr]->dev);
            svdgpu_nccl_bcast(g->comm[lr], c->base[lr] + c->off_small, (size_t)nsmall, 0, c->cx[lr]->s_comm);
        }
        svdgpu_nccl_group_end();
    }
    /* ------------------------------------------------------------ every rank: its block of vectors.
     * One process driving several GPUs enqueues each rank's twisted + back-transform work from its own host
     * thread: a back-transform is ~2000 launches, and enqueued rank after rank from one thread the last GPU
     * would start ~N x 0.1 s late (measured on 8 GPUs at 32768^2 before this: 0.63 s on rank 0, 1.61 s on rank 7). */
    rank_job jobs[SVD_MAX_DEV];
    pthread_t th[SVD_MAX_DEV];
    for (int lr = 0; lr < g->nlocal; ++lr) {
        jobs[lr].c = c; jobs[lr].lr = lr; jobs[lr].dU = dU[lr]; jobs[lr].dV = dV[lr]; jobs[lr].ldu = ldu; jobs[lr].ldv = ldv;
    }
    if (g->nlocal == 1) {
        rank_vectors(&jobs[0]);
    } else {
        for (int lr = 0; lr < g->nlocal; ++lr)
            if (pthread_create(&th[lr], NULL, rank_vectors, &jobs[lr]) != 0) { fprintf(stderr, "svd_gpu: pthread_create failed\n"); abort(); }
        for (int lr = 0; lr < g->nlocal; ++lr) pthread_join(th[lr], NULL);
    }
    /* polished singular values of all blocks to everyone (after the vector work in every stream: nothing but the
     * returned sigma depends on it) */
    if (world > 1) {
        svdgpu_nccl_group_start();
        for (int lr = 0; lr < g->nlocal; ++lr) {
            svd_ctx *x = c->cx[lr];
            svdgpu_set_device(x->dev);
            svdgpu_stream_wait_event(x->s_comm, x->ev[3]);
            svdgpu_nccl_allgather(g->comm[lr], c->base[lr] + c->off_sigblk, c->base[lr] + c->off_siggather,
                                  (size_t)c->blk, x->s_comm);
        }
        svdgpu_nccl_group_end();
    }
    for (int lr = 0; lr < g->nlocal; ++lr) {
        svd_ctx *x = c->cx[lr];
        svdgpu_set_device(x->dev);
        double *dscale = (double *)(c->base[lr] + c->off_small) + 3 * (size_t)mn + 1;
        /* sigma: all polished blocks (i0 = rank * blk, so the gathered array is already in order), scaled back */
        if (world > 1) {
            svdgpu_event_record(x->ev_commdone, x->s_comm);
            svdgpu_stream_wait_event(x->s_main, x->ev_commdone);
            svdgpu_d2d(dsigma[lr], c->base[lr] + c->off_siggather, sizeof(double) * (size_t)mn, x->s_main);
        } else {
            svdgpu_d2d(dsigma[lr], c->base[lr] + c->off_sigblk, sizeof(double) * (size_t)mn, x->s_main);
        }
        svdgpu_scale_vector(mn, dsigma[lr], dscale + 1, x->s_main);
        svdgpu_event_record(x->ev[4], x->s_main);
    }
}

/* Whole path on a group.  streams[lr]: the caller's stream on local rank lr (work is ordered after what is
 * already enqueued there, and the stream waits for the results). */
/* everything about a call that depends only on (group, m, n): route, reflector sets, arena layout */
static void call_plan(svd_call *c, svdgpu_group *g, int m, int n, long lda, int *wide_out, long *ldt_out, size_t *bytesAt_out)
{
    const int wide = use_wide_transpose(m, n);
    const int tm = wide ? n : m, tn = wide ? m : n;           /* the problem the core solves */
    memset(c, 0, sizeof *c);
    c->grp = g; c->m = tm; c->n = tn; c->mn = tm < tn ? tm : tn; c->world = g->world;
    c->mb = ((tm >= tn) ? tn - 1 : tm) + 1;
    c->qr = use_qr_first(tm, tn);
    c->has_root = (g->rank0 == 0);
    c->progressive = (g->world > 1) || g_opt.wy_overlap == 1;
    svdgpu_shard_range(c->mn, g->world, 0, &c->blk, NULL, NULL);
    const long ldt = wide ? (n + 1) / 2 * 2 : lda;
    plan_sets(c, NULL, ldt);
    plan_arena(c, ldt);
    *wide_out = wide; *ldt_out = ldt;
    *bytesAt_out = wide ? up256(sizeof(double) * (size_t)ldt * m) : 0;
}
/* device bytes the core needs on local rank lr */
static size_t call_arena_bytes(const svd_call *c, int lr, size_t bytesAt)
{
    return (c->grp->rank0 + lr == 0) ? c->total_root + bytesAt : c->total_peer;
}

static void svd_group_dev(svdgpu_group *g, int m, int n, double *dA, long lda, double *const *dsigma,
                          double *const *dU, long ldu, double *const *dV, long ldv, void *const *streams,
                          void *ev_after_fact, int locked)
{
    svd_call *c = (svd_call *)malloc(sizeof *c);
    if (!c) abort();
    if ((dU == NULL) != (dV == NULL)) {
        fprintf(stderr, "svd_gpu: U and V must both be given or both be NULL (values only)\n");
        abort();
    }
    int wide; long ldt; size_t bytesAt;
    call_plan(c, g, m, n, lda, &wide, &ldt, &bytesAt);
    const int cur = svdgpu_get_device();
    for (int lr = 0; lr < g->nlocal; ++lr) {
        c->cx[lr] = locked ? &g_ctx[g->dev[lr]] : ctx_acquire(g->dev[lr]);
        c->cx[lr]->first_recorded = 0;
        c->cx[lr]->was_root = (g->rank0 + lr == 0);
    }
    for (int lr = 0; lr < g->nlocal; ++lr) {
        svd_ctx *x = c->cx[lr];
        svdgpu_set_device(x->dev);
        c->base[lr] = arena_get(x, x->io_reserved + call_arena_bytes(c, lr, bytesAt)) + x->io_reserved;
        /* fork: the library's streams start after what the caller has enqueued */
        svdgpu_event_record(x->ev_in, streams ? streams[lr] : NULL);
        svdgpu_stream_wait_event(x->s_main, x->ev_in);
        svdgpu_stream_wait_event(x->s_side, x->ev_in);
        svdgpu_stream_wait_event(x->s_comm, x->ev_in);
        svdgpu_event_record(x->ev[0], x->s_main);
        memset(x->ms, 0, sizeof x->ms);
    }
    double *dwork = dA;
    if (c->has_root) {
        svd_ctx *r = c->cx[0];
        svdgpu_set_device(r->dev);
        if (wide) {
            dwork = (double *)(c->base[0] + c->total_root);
            if (ldt != n) svdgpu_memset(dwork, 0, sizeof(double) * (size_t)ldt * m, r->s_main);
            svdgpu_transpose(m, n, dA, lda, dwork, ldt, r->s_main);
        }
        for (int q = 0; q < 3; ++q) { c->set[q].A = c->qr && q > 0 ? NULL : dwork; c->set[q].lda = ldt; }
    }
    svd_core(c, dwork, ldt, dsigma, wide ? dV : dU, wide ? ldv : ldu, wide ? dU : dV, wide ? ldu : ldv,
             wide ? NULL : ev_after_fact);
    if (c->has_root && wide) {
        svd_ctx *r = c->cx[0];
        svdgpu_set_device(r->dev);
        /* A leaves as the transpose of the tall problem's reflector storage */
        svdgpu_transpose(n, m, dwork, ldt, dA, lda, r->s_main);
        if (ev_after_fact) svdgpu_event_record(ev_after_fact, r->s_main);
    }
    /* join: the caller's streams wait for the results */
    for (int lr = 0; lr < g->nlocal; ++lr) {
        svd_ctx *x = c->cx[lr];
        svdgpu_set_device(x->dev);
        if (wide) x->first_is_u = !x->first_is_u;
        svdgpu_event_record(x->ev_done, x->s_main);
        svdgpu_stream_wait_event(streams ? streams[lr] : NULL, x->ev_done);
        svdgpu_event_record(x->ev_commdone, x->s_comm);
        svdgpu_stream_wait_event(streams ? streams[lr] : NULL, x->ev_commdone);
        svdgpu_event_record(x->ev_panels, x->s_side);
        svdgpu_stream_wait_event(streams ? streams[lr] : NULL, x->ev_panels);
        x->ms_pending = 1;
    }
    g_last_dev = g->dev[0];
    if (!locked) for (int lr = g->nlocal - 1; lr >= 0; --lr) ctx_release(c->cx[lr]);
    svdgpu_set_device(cur);
    free(c);
}

static void read_phase_ms(svd_ctx *c)
{
    if (!c->ms_pending) return;
    svdgpu_set_device(c->dev);
    if (c->was_root) {
        c->ms[1] = svdgpu_event_elapsed_ms(c->ev[0], c->ev[1]);
        c->ms[2] = svdgpu_event_elapsed_ms(c->ev[1], c->ev[2]);
        c->ms[7] = svdgpu_event_elapsed_ms(c->ev[2], c->ev[5]);   /* exposed wait for the last panels (N > 1) */
    } else {
        /* a rank without the factorization: only the vector phases are its own */
        c->ms[1] = c->ms[2] = 0.f;
        c->ms[7] = svdgpu_event_elapsed_ms(c->ev[0], c->ev[5]);
    }
    c->ms[3] = svdgpu_event_elapsed_ms(c->ev[5], c->ev[3]);
    c->ms[4] = svdgpu_event_elapsed_ms(c->ev[3], c->ev[4]);
    if (c->ms[0] == 0.f && c->ms[5] == 0.f) c->ms[6] = svdgpu_event_elapsed_ms(c->ev[0], c->ev[4]);
    c->ms_pending = 0;
}

void svd_gpu_last_phase_ms(float ms[7])
{
    opts_init();
    const int cur = svdgpu_get_device();
    svd_ctx *c = &g_ctx[g_last_dev];
    pthread_mutex_lock(&c->mu);
    if (c->inited) read_phase_ms(c);
    memcpy(ms, c->ms, 7 * sizeof(float));
    pthread_mutex_unlock(&c->mu);
    svdgpu_set_device(cur);
}

void svd_gpu_group_phase_ms(svdgpu_group *g, int local, float ms[8])
{
    const int cur = svdgpu_get_device();
    svd_ctx *c = &g_ctx[g->dev[local]];
    pthread_mutex_lock(&c->mu);
    if (c->inited) read_phase_ms(c);
    memcpy(ms, c->ms, 8 * sizeof(float));
    pthread_mutex_unlock(&c->mu);
    svdgpu_set_device(cur);
}

/* ---- device-resident entry points ---------------------------------------------------------------- */
void svd_gpu_dev(int m, int n, double *dA, long lda, double *dsigma, double *dU, long ldu, double *dV, long ldv,
                 void *stream)
{
    opts_init();
    device_init();
    svdgpu_group g;
    memset(&g, 0, sizeof g);
    g.world = g.nlocal = 1; g.rank0 = 0; g.dev[0] = svdgpu_get_device();
    double *sg[1] = {dsigma}, *u[1] = {dU}, *v[1] = {dV};
    void *st[1] = {stream};
    svd_group_dev(&g, m, n, dA, lda, sg, (dU && dV) ? u : NULL, ldu, (dU && dV) ? v : NULL, ldv, st, NULL, 0);
}

void svd_gpu_sharded_dev(svdgpu_group *g, int m, int n, double *dA_root, long lda, double *const *dsigma,
                         double *const *dUblk, long ldu, double *const *dVblk, long ldv, void *const *streams)
{
    opts_init();
    if (!dUblk || !dVblk) { fprintf(stderr, "svd_gpu_sharded_dev: values only runs on one rank, use svd_gpu_dev\n"); abort(); }
    svd_group_dev(g, m, n, dA_root, lda, dsigma, dUblk, ldu, dVblk, ldv, streams, NULL, 0);
}

/* low-level building blocks kept from the first multi-GPU version (direct route only: no range guard,
 * no QR first, no transpose): bidiagonalization + dDC, and the vector phases of one block */
void svd_gpu_values_dev(int m, int n, double *dA, long lda, double *dalpha, double *dbeta, double *dsigma,
                        void *stream)
{
    device_init();
    svd_ctx *c = ctx_acquire(svdgpu_get_device());
    const int mn = m < n ? m : n;
    char *work = arena_get(c, up256(maxz(svdgpu_bidiag_workspace(m, n, lda), svdgpu_ddc_workspace(mn))));
    svdgpu_memset(dbeta, 0, sizeof(double) * (size_t)mn, stream);     /* beta[mn-1] = 0 when B is square */
    svdgpu_bidiag(m, n, dA, lda, dalpha, dbeta, work, g_opt.nb, stream);
    svdgpu_ddc_values(mn, dalpha, dbeta, dsigma, work, stream);
    ctx_release(c);
}

void svd_gpu_vectors_dev(int m, int n, const double *dA_mod, long lda, const double *dalpha,
                         const double *dbeta, const double *dsigma_all, int i0, int ns, double *dUblk,
                         long ldu, double *dVblk, long ldv, double *dsig_out, void *stream)
{
    if (ns <= 0) return;
    device_init();
    svd_ctx *c = ctx_acquire(svdgpu_get_device());
    const int mn = m < n ? m : n, len_beta = (m >= n) ? n - 1 : m, mb = len_beta + 1;
    size_t w = svdgpu_twisted_workspace(mn, mb, ns);
    w = maxz(w, svdgpu_backtransform_workspace(m, n_left(m, n), ns));
    w = maxz(w, svdgpu_backtransform_workspace(n, n_right(m, n), ns));
    char *work = arena_get(c, up256(w));
    svdgpu_memset(dUblk, 0, sizeof(double) * (size_t)ldu * ns, stream);
    svdgpu_memset(dVblk, 0, sizeof(double) * (size_t)ldv * ns, stream);
    svdgpu_twisted_vectors(mn, mb, dalpha, dbeta, dsigma_all, mn, i0, ns, dVblk, ldv, dUblk, ldu, dsig_out,
                           g_opt.rqi, work, stream);
    svdgpu_wy_apply(1, m, n_left(m, n), dA_mod, lda, dUblk, ldu, ns, work, stream);
    svdgpu_wy_apply(0, n, n_right(m, n), dA_mod, lda, dVblk, ldv, ns, work, stream);
    ctx_release(c);
}

/* ---- host-pointer entry points ------------------------------------------------------------------- */
static svdgpu_group *g_host_group = NULL;     /* the group svd_gpu() runs on (SVD_GPU_NGPUS) */

static svdgpu_group *host_group(void)
{
    pthread_mutex_lock(&g_mu);
    int want = g_opt.ngpus < 1 ? 1 : g_opt.ngpus;
    const int have = svdgpu_device_count();
    if (want > have) {
        fprintf(stderr, "svd_gpu: SVD_GPU_NGPUS=%d but only %d devices are visible\n", want, have);
        abort();
    }
    const int first = svdgpu_get_device();
    if (g_host_group && (g_host_group->world != want || g_host_group->dev[0] != first)) {
        svdgpu_group_destroy(g_host_group);
        g_host_group = NULL;
    }
    if (!g_host_group) {
        int devs[SVD_MAX_DEV];
        for (int i = 0; i < want; ++i) devs[i] = (first + i) % have;
        pthread_mutex_unlock(&g_mu);
        svdgpu_group *ng = svdgpu_group_create_local(want, devs);
        pthread_mutex_lock(&g_mu);
        g_host_group = ng;
    }
    svdgpu_group *g = g_host_group;
    pthread_mutex_unlock(&g_mu);
    return g;
}

/* The whole path from host buffers on a group.  A (root), sigma (root): as svd_gpu(); Ublk[lr] / Vblk[lr]:
 * host destination of local rank lr's column block (m x ns, ld m / n x ns, ld n). */
void svd_gpu_sharded(svdgpu_group *g, int m, int n, double *A, double *sigma, double *const *Ublk,
                     double *const *Vblk)
{
    if (m <= 0 || n <= 0 || (g->rank0 == 0 && (!A || !sigma))) {
        fprintf(stderr, "svd_gpu: bad arguments (m=%d n=%d)\n", m, n);
        abort();
    }
    opts_init();
    device_init();
    const double t_start = wall_ms();
    const int mn = m < n ? m : n;
    const int want_vec = (Ublk != NULL && Vblk != NULL);
    const long lda = (m + 1) / 2 * 2;
    const int cur = svdgpu_get_device();
    const int has_root = (g->rank0 == 0);
    int blk;
    svdgpu_shard_range(mn, g->world, 0, &blk, NULL, NULL);
    svd_ctx *cx[SVD_MAX_DEV];
    double *dA = NULL, *dsg[SVD_MAX_DEV], *dU[SVD_MAX_DEV], *dV[SVD_MAX_DEV];
    void *st[SVD_MAX_DEV];
    const size_t bytesA = up256(sizeof(double) * (size_t)lda * n);
    const size_t bytesU = want_vec ? up256(sizeof(double) * (size_t)m * blk) : 0;
    const size_t bytesV = want_vec ? up256(sizeof(double) * (size_t)n * blk) : 0;
    const size_t bytesS = up256(sizeof(double) * (size_t)mn);
    /* the device images of A, U, V, sigma sit at the start of each rank's cached arena, the core's
     * workspace behind them */
    {
        svd_call *pl = (svd_call *)malloc(sizeof *pl);
        if (!pl) abort();
        int wide; long ldt; size_t bytesAt;
        call_plan(pl, g, m, n, lda, &wide, &ldt, &bytesAt);
        for (int lr = 0; lr < g->nlocal; ++lr) {
            cx[lr] = ctx_acquire(g->dev[lr]);
            const int is_root = (g->rank0 + lr == 0);
            cx[lr]->io_reserved = (is_root ? bytesA : 0) + bytesU + bytesV + bytesS;
            char *p = arena_get(cx[lr], cx[lr]->io_reserved + call_arena_bytes(pl, lr, bytesAt));
            if (is_root) { dA = (double *)p; p += bytesA; }
            dU[lr] = want_vec ? (double *)p : NULL; p += bytesU;
            dV[lr] = want_vec ? (double *)p : NULL; p += bytesV;
            dsg[lr] = (double *)p;
            st[lr] = cx[lr]->s_main;
        }
        free(pl);
    }
    /* The reference's callers pass malloc'd memory (test-whole-svd.c:44-66).  Page-locking it for the call
     * ("host_register") is optional and OFF by default: measured on a B200 box (bench.py e2e_pageable),
     * cudaHostRegister + unregister cost more than the pageable copies lose — 16384^2: 4.55 s registered vs
     * 3.77 s plain (3.57 s with buffers the caller page-locked once); 4096^2: 0.277 vs 0.121 (0.108). */
    int regA = 1, regU[SVD_MAX_DEV], regV[SVD_MAX_DEV];
    const double t_h2d0 = wall_ms();
    for (int lr = 0; lr < g->nlocal; ++lr) { regU[lr] = regV[lr] = 1; }
    if (g_opt.host_register) {
        if (has_root) regA = svdgpu_host_register(A, sizeof(double) * (size_t)m * n);
        for (int lr = 0; want_vec && lr < g->nlocal; ++lr) {
            int ns;
            svdgpu_shard_range(mn, g->world, g->rank0 + lr, NULL, NULL, &ns);
            if (ns > 0) {
                regU[lr] = svdgpu_host_register(Ublk[lr], sizeof(double) * (size_t)m * ns);
                regV[lr] = svdgpu_host_register(Vblk[lr], sizeof(double) * (size_t)n * ns);
            }
        }
    }
    /* host -> device (the reference's blocking clEnqueueWriteBuffer, bidiag_par.c:298-301) */
    if (has_root) {
        svd_ctx *r = cx[0];
        svdgpu_set_device(r->dev);
        svdgpu_range_push("svd_gpu:h2d");
        if (lda == m) {
            svdgpu_h2d(dA, A, sizeof(double) * (size_t)m * n, r->s_main);
        } else {
            svdgpu_memset(dA, 0, sizeof(double) * (size_t)lda * n, r->s_main);
            svdgpu_h2d_2d(dA, sizeof(double) * lda, A, sizeof(double) * m, sizeof(double) * m, n, r->s_main);
        }
        svdgpu_stream_sync(r->s_main);
        svdgpu_range_pop();
    }
    const float ms_h2d = (float)(wall_ms() - t_h2d0);

    svd_group_dev(g, m, n, dA, lda, dsg, want_vec ? dU : NULL, m, want_vec ? dV : NULL, n, st,
                  has_root ? cx[0]->ev_copy : NULL, 1);

    /* device -> host.  The reflector matrix goes back on the copy stream as soon as the factorization is
     * done (nothing modifies dA afterwards), overlapping the later phases; so does whichever of U / V is
     * final first, while the other one is still being back-transformed.  Every rank copies its own column
     * block (contiguous in the caller's column-major U, V) over its own PCIe link. */
    int early[SVD_MAX_DEV];
    for (int lr = 0; lr < g->nlocal; ++lr) {
        svd_ctx *x = cx[lr];
        svdgpu_set_device(x->dev);
        int ns;
        svdgpu_shard_range(mn, g->world, g->rank0 + lr, NULL, NULL, &ns);
        early[lr] = -1;
        if (g->rank0 + lr == 0) {
            svdgpu_stream_wait_event(x->s_copy, x->ev_copy);
            if (lda == m) svdgpu_d2h(A, dA, sizeof(double) * (size_t)m * n, x->s_copy);
            else svdgpu_d2h_2d(A, sizeof(double) * m, dA, sizeof(double) * lda, sizeof(double) * m, n, x->s_copy);
        }
        if (want_vec && ns > 0 && x->first_recorded) {
            early[lr] = x->first_is_u;
            svdgpu_stream_wait_event(x->s_copy, x->ev_first);
            if (early[lr]) svdgpu_d2h(Ublk[lr], dU[lr], sizeof(double) * (size_t)m * ns, x->s_copy);
            else svdgpu_d2h(Vblk[lr], dV[lr], sizeof(double) * (size_t)n * ns, x->s_copy);
        }
    }
    for (int lr = 0; lr < g->nlocal; ++lr) { svdgpu_set_device(cx[lr]->dev); svdgpu_stream_sync(cx[lr]->s_main); }
    const double t_d2h0 = wall_ms();
    for (int lr = 0; lr < g->nlocal; ++lr) {
        svd_ctx *x = cx[lr];
        svdgpu_set_device(x->dev);
        int ns;
        svdgpu_shard_range(mn, g->world, g->rank0 + lr, NULL, NULL, &ns);
        if (g->rank0 + lr == 0) svdgpu_d2h(sigma, dsg[lr], sizeof(double) * (size_t)mn, x->s_main);
        if (want_vec && ns > 0) {
            if (early[lr] != 1) svdgpu_d2h(Ublk[lr], dU[lr], sizeof(double) * (size_t)m * ns, x->s_main);
            if (early[lr] != 0) svdgpu_d2h(Vblk[lr], dV[lr], sizeof(double) * (size_t)n * ns, x->s_main);
        }
    }
    for (int lr = 0; lr < g->nlocal; ++lr) {
        svd_ctx *x = cx[lr];
        svdgpu_set_device(x->dev);
        svdgpu_stream_sync(x->s_main);
        svdgpu_stream_sync(x->s_copy);
        svdgpu_stream_sync(x->s_comm);
        svdgpu_stream_sync(x->s_side);
    }
    if (!regA) svdgpu_host_unregister(A);
    for (int lr = 0; lr < g->nlocal; ++lr) {
        if (!regU[lr]) svdgpu_host_unregister(Ublk[lr]);
        if (!regV[lr]) svdgpu_host_unregister(Vblk[lr]);
    }
    for (int lr = 0; lr < g->nlocal; ++lr) {
        svd_ctx *x = cx[lr];
        svdgpu_set_device(x->dev);
        read_phase_ms(x);
        x->ms[0] = ms_h2d;
        x->ms[5] = (float)(wall_ms() - t_d2h0);
        x->ms[6] = (float)(wall_ms() - t_start);
        x->io_reserved = 0;
    }
    for (int lr = g->nlocal - 1; lr >= 0; --lr) ctx_release(cx[lr]);
    svdgpu_set_device(cur);
}

/* ---- the drop-in entry point (svd_gpu.h:5, svd_gpu.c:53) ---------------------------- */
void svd_gpu(int m, int n, double *A, double *sigma, double *U, double *V)
{
    if (m <= 0 || n <= 0 || !A || !sigma) {
        fprintf(stderr, "svd_gpu: bad arguments (m=%d n=%d)\n", m, n);
        abort();
    }
    opts_init();
    device_init();
    const int mn = m < n ? m : n;
    const int want_vec = (U != NULL && V != NULL);
    if ((U == NULL) != (V == NULL)) {
        fprintf(stderr, "svd_gpu: U and V must both be given or both be NULL (values only)\n");
        abort();
    }
    svdgpu_group *g = host_group();
    if (!want_vec) {
        /* values only: one rank */
        svdgpu_group one;
        memset(&one, 0, sizeof one);
        one.world = one.nlocal = 1; one.dev[0] = g->dev[0];
        svd_gpu_sharded(&one, m, n, A, sigma, NULL, NULL);
        return;
    }
    double *ub[SVD_MAX_DEV], *vb[SVD_MAX_DEV];
    for (int lr = 0; lr < g->nlocal; ++lr) {
        int i0;
        svdgpu_shard_range(mn, g->world, lr, NULL, &i0, NULL);
        ub[lr] = U + (size_t)i0 * m;           /* only the first min(m,n) columns are written (svd_gpu.c:118-121) */
        vb[lr] = V + (size_t)i0 * n;
    }
    svd_gpu_sharded(g, m, n, A, sigma, ub, vb);
}

/* ---- the reference driver's check, enabled (test-whole-svd.c:81-96) ------------------------------- */
void svd_gpu_check_dev(int m, int n, const double *dA0, long lda, const double *dsigma, const double *dU, long ldu,
                       const double *dV, long ldv, int nc, double out6[6], void *stream)
{
    opts_init();
    device_init();
    void *work = svdgpu_malloc(svdgpu_check_workspace(m, n, nc) + 64);
    double *dout = (double *)work;
    svdgpu_check(m, n, dA0, lda, dsigma, dU, ldu, dV, ldv, nc, dout, (char *)work + 64, stream);
    svdgpu_d2h(out6, dout, 6 * sizeof(double), stream);
    svdgpu_stream_sync(stream);
    svdgpu_free(work);
}

void svd_gpu_check(int m, int n, const double *A0, const double *sigma, const double *U, const double *V,
                   double out6[6])
{
    opts_init();
    device_init();
    const int mn = m < n ? m : n;
    double *dA = (double *)svdgpu_malloc(sizeof(double) * ((size_t)m * n + (size_t)m * mn + (size_t)n * mn + mn));
    double *dU = dA + (size_t)m * n, *dV = dU + (size_t)m * mn, *ds = dV + (size_t)n * mn;
    svdgpu_h2d(dA, A0, sizeof(double) * (size_t)m * n, NULL);
    svdgpu_h2d(dU, U, sizeof(double) * (size_t)m * mn, NULL);
    svdgpu_h2d(dV, V, sizeof(double) * (size_t)n * mn, NULL);
    svdgpu_h2d(ds, sigma, sizeof(double) * (size_t)mn, NULL);
    svd_gpu_check_dev(m, n, dA, m, ds, dU, m, dV, n, mn, out6, NULL);
    svdgpu_free(dA);
}
