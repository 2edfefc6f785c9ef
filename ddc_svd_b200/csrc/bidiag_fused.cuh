// bidiag_fused.cuh — the fused streaming pass of the bidiagonalization (included by bidiag.cu).
//
// One Golub-Kahan step needs  t1 = A^T c  (column dots, for the left reflector) and then
// t2 = A r  (row dots, for the right reflector), where r depends on ALL of t1 only through the
// entries of its own column:  r_j = f(t1_j, panel data).  So if a column tile of A stays on
// chip between the two uses, both products come out of ONE read of the trailing matrix:
//
//     for every column tile J (full height, CBW columns):
//         sweep 1:  t1_J = A[:,J]^T c                (tile in shared memory)
//         y_J, r_J  from t1_J and the panel rows     (thin, per column)
//         sweep 2:  t2 += A[:,J] r_J                 (same shared-memory tile)
//
// This halves the HBM traffic of the split gemvT/gemvN passes.  sm_100a mapping:
//   * the tile (<= 32 KB per stage, 6 stages: two tiles being consumed, four in flight) is brought in by TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx), one per column segment, issued by an elected
//     lane of a dedicated producer warp; full/empty mbarriers form the pipeline;
//   * helper warps (3) gather, while the copies fly, the tile's panel rows (Y[j,:], U[j,:]) and reduce
//     the 2k-term corrections that y_j / r_j need;
//   * the two sweeps are done by DIFFERENT warps connected only by mbarriers (no CTA-wide barrier
//     in the loop): 8 sweep-1 warps (c in registers) hand their column sums to a reducer warp, a
//     finisher warp turns the cluster-wide sums into y_j, r_j, 8 sweep-2 warps (t2 accumulators in
//     registers) consume r_j and release the stage.  Every warp runs ahead on its own, so several
//     tiles are in different phases at once, which is what hides the per-tile latency chain;
//   * when a column is taller than one stage can hold (4096 rows per CTA) the rows are split
//     over a thread-block CLUSTER of 2/4/8 CTAs; the per-column partial sums of sweep 1 are
//     exchanged through distributed shared memory (st.async ... mbarrier::complete_tx on the
//     remote barrier, so receivers need no cluster-scope acquire);
//   * clusters walk the column tiles round-robin; every cluster writes its partial t2 and its
//     partial panel dots; the consumer kernel (finish_xf) combines the partials in a fixed order
//     in its prologue (deterministic: no floating-point atomics, no fences, no serial tail).
#pragma once

namespace svdgpu {

constexpr int FZ_STAGE = 4096;            // doubles per shared-memory stage (32 KB)
constexpr int FZ_STAGES = 6;
constexpr int FZ_CBW_MAX = 4;             // columns per tile (rows per CTA <= 1024 -> 4 columns)
constexpr int FZ_XR = 16;                 // ring depth of the cross-CTA exchange (>= 2*FZ_STAGES)
// compile-time switches for A/B builds on the same box (bench/ab_build.sh); the defaults are the measured best
// (profiles/r02_ab_helper_warps.log: 3 helper warps instead of 2 = -2.8 % at 16384^2, -4.3 % at 8192^2, -1.2 % at
// 4096^2 - two helpers, each a trip to L2 per tile, could not keep up with the stream; 4 and 5 are slower again;
// batching the sweeps' shared-memory loads is neutral; requesting a step's first panel rows before the prologue
// costs more in registers than it hides)
#ifndef SVDGPU_FZ_HW
#define SVDGPU_FZ_HW 3
#endif
#ifndef SVDGPU_FZ_BATCH
#define SVDGPU_FZ_BATCH 4
#endif
#ifndef SVDGPU_FZ_WP
#define SVDGPU_FZ_WP 1
#endif
#ifndef SVDGPU_FZ_EARLY
#define SVDGPU_FZ_EARLY 0
#endif
constexpr int FZ_HW = SVDGPU_FZ_HW;       // helper warps (tile n is prepared by helper n % FZ_HW)
constexpr int FZ_GW = 8;                  // warps per sweep group
constexpr int FZ_GT = FZ_GW * 32;         // threads per sweep group
// warp roles: 0 TMA producer | 1..HW helpers | HW+1 reducer | NFIN finishers |
//             GW sweep-1 warps | GW sweep-2 warps
constexpr int FZ_NFIN = 2;                // finisher warps (tile n is finished by finisher n % FZ_NFIN)
constexpr int FZ_W_RED = 1 + FZ_HW;
constexpr int FZ_W_FIN = 2 + FZ_HW;
constexpr int FZ_W_S1 = FZ_W_FIN + FZ_NFIN;
constexpr int FZ_W_S2 = FZ_W_S1 + FZ_GW;
constexpr int FZ_WARPS = FZ_W_S2 + FZ_GW;
constexpr int FZ_THREADS = FZ_WARPS * 32;
constexpr int FZ_MAXCS = 8;               // largest cluster
constexpr int FZ_MAX_CLUSTERS = 148;
constexpr bool FZ_DEFAULT_ON = true;      // measured faster than the split passes at 4096^2 and 16384^2 (profiles/)
constexpr int FZ_MIN_ROWS = 16;           // below this trailing height / width the split passes are used (measured:
constexpr int FZ_MIN_COLS = 8;            //  with dependent launch the fused step wins down to a handful of rows)

constexpr size_t FZ_SMEM_DOUBLES = (size_t)FZ_STAGES * FZ_STAGE                      // tiles
                                   + (size_t)FZ_STAGES * FZ_CBW_MAX * 2 * NBMAX      // panel rows of the tile columns
                                   + 5 * FZ_STAGES * FZ_CBW_MAX                      // corr, g, a_ij, y, r
                                   + FZ_STAGES * FZ_GW * FZ_CBW_MAX                  // per-warp column sums
                                   + FZ_XR * FZ_MAXCS * FZ_CBW_MAX                   // exchanged column sums
                                   + 4 * NBMAX + 8                                   // vTv, xTv, rowV, rowX, scalars
                                   + (FZ_NFIN - 1) * (2 * NBMAX + 2);                // panel-dot partials of the other finishers
constexpr size_t FZ_SMEM_BYTES = FZ_SMEM_DOUBLES * 8 + (4 * FZ_STAGES + FZ_XR) * 8 + 16 * sizeof(int) + 128;

struct FusedArgs {
    double *A; long lda;          // trailing matrix (read) and column i (the reflector is written in place)
    int i, m, n, mpad, k, nb;
    double *P; long ldp;
    double *Q; long ldq;
    const double *c;
    double *rv;
    double *tmpN; long ldt;       // [cluster][row]: partial A r
    const double *dots1;          // [V^T c | X^T c | c.c]: final (nparts1 == 0) or nparts1 partial vectors
    int nparts1;                  //   of stride 2*NBMAX+2 left by finish_xf, combined in the prologue here
    double *dots2p;               // [cluster][2*NBMAX+2] partials of [Y^T r | U^T r | r.r] for finish_xf
    double *alpha;
    int T, NC, Lc;                // column tiles, clusters, rows per CTA (even)
    unsigned long long *trace;    // debugging aid (SVD_GPU_FZ_TRACE=step): per-tile clock64 stamps of CTA 0, else null
    int prefetch;                 // 1: the preceding kernel does not write the trailing matrix, so the first
                                  //    tiles may be fetched before griddepcontrol.wait (programmatic dependent launch)
};

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ unsigned fz_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fz_mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fz_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fz_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fz_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fz_mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fz_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fz_mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FZ_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FZ_DONE_%=;\n"
        "bra FZ_WAIT_%=;\n"
        "FZ_DONE_%=:\n"
        "}\n" ::"r"(fz_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool fz_mbar_test(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(fz_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fz_mbar_wait_cluster(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FZ_WAITC_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FZ_DONEC_%=;\n"
        "bra FZ_WAITC_%=;\n"
        "FZ_DONEC_%=:\n"
        "}\n" ::"r"(fz_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ unsigned fz_mapa(unsigned saddr, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void fz_st_cluster(unsigned caddr, double v)
{
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(caddr), "d"(v) : "memory");
}
// remote shared-memory store that completes a transaction on the REMOTE mbarrier: the data is
// visible to whoever observes the barrier phase with an ordinary CTA-scope wait (no cluster-scope
// acquire, hence no L1 invalidate on the waiting side)
__device__ __forceinline__ void fz_st_async(unsigned caddr, double v, unsigned cbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(caddr),
                 "l"(__double_as_longlong(v)), "r"(cbar)
                 : "memory");
}
__device__ __forceinline__ void fz_mbar_arrive_cluster(unsigned caddr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ void fz_bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     fz_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(fz_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fz_cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}


// ---- the two inner products of a tile column, per sweep thread ---------------------------------
// The shared-memory loads are issued four at a time into distinct registers before the FMAs that use them: with
// one destination register set (what the compiler falls back to under the 80-register cap of a 24-warp CTA) every
// LDS waits for the FMAs of the previous one and a 4096-row column costs ~8 x (LDS + DFMA) latencies per tile.
template <int RPT>
__device__ __forceinline__ double fz_col_dot(const double *__restrict__ col, const double2 (&creg)[RPT], int gt, int len)
{
    constexpr int B = RPT < SVDGPU_FZ_BATCH ? RPT : SVDGPU_FZ_BATCH;
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;      // independent chains (same order as ever: results unchanged)
#pragma unroll
    for (int u0 = 0; u0 < RPT; u0 += B) {
        double2 av[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int lr = 2 * gt + 2 * FZ_GT * (u0 + b);
            av[b] = (lr < len) ? *reinterpret_cast<const double2 *>(col + lr) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int u = u0 + b;
            if (u & 1) { p2 += av[b].x * creg[u].x; p3 += av[b].y * creg[u].y; }
            else       { p0 += av[b].x * creg[u].x; p1 += av[b].y * creg[u].y; }
        }
    }
    return (p0 + p1) + (p2 + p3);
}
template <int RPT>
__device__ __forceinline__ void fz_col_axpy(const double *__restrict__ col, double r, double2 (&acc)[RPT], int gt, int len)
{
    constexpr int B = RPT < SVDGPU_FZ_BATCH ? RPT : SVDGPU_FZ_BATCH;
#pragma unroll
    for (int u0 = 0; u0 < RPT; u0 += B) {
        double2 av[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const int lr = 2 * gt + 2 * FZ_GT * (u0 + b);
            av[b] = (lr < len) ? *reinterpret_cast<const double2 *>(col + lr) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int b = 0; b < B; ++b) {
            acc[u0 + b].x += av[b].x * r;
            acc[u0 + b].y += av[b].y * r;
        }
    }
}

#define FZ_TR(slot, nt) do { if (a.trace && blockIdx.x == 0 && (nt) < 256) a.trace[(nt) * 8 + (slot)] = clock64(); } while (0)
// RPT = row pairs per sweep thread (rows per CTA <= 512*RPT <= 4096), CBW columns per tile, NST stages of
// FZ_STAGES*FZ_STAGE/NST doubles.  Classic geometries: CBW = 8/RPT, 6 stages of 32 KB.  <8, 2, 4>: two columns of up
// to 3072 rows per 48 KB stage - the per-tile role chain (barrier hops, lane reduction, reducer, finisher) is the
// limit for short columns, and a tile of two columns pays it once for twice the data.
template <int RPT, int CBW, int NST>
__global__ void __launch_bounds__(FZ_THREADS, 1) fused_pass_kernel(const FusedArgs a)
{
    constexpr int STG = FZ_STAGES * FZ_STAGE / NST;       // doubles per stage
    static_assert(NST <= FZ_STAGES && CBW <= FZ_CBW_MAX && FZ_XR >= 2 * NST, "fused pass geometry");
    constexpr int WP = SVDGPU_FZ_WP ? FZ_CBW_MAX / CBW : 1;    // lane partials each sweep-1 warp leaves per column (fills the same 32 slots per stage)
    extern __shared__ __align__(128) unsigned char fz_smem[];
    double *tile = reinterpret_cast<double *>(fz_smem);
    double *qrow = tile + (size_t)FZ_STAGES * FZ_STAGE;
    double *hcorr = qrow + FZ_STAGES * FZ_CBW_MAX * 2 * NBMAX;
    double *hg = hcorr + FZ_STAGES * FZ_CBW_MAX;
    double *haij = hg + FZ_STAGES * FZ_CBW_MAX;
    double *yq = haij + FZ_STAGES * FZ_CBW_MAX;
    double *rq = yq + FZ_STAGES * FZ_CBW_MAX;
    double *wsum = rq + FZ_STAGES * FZ_CBW_MAX;                      // [stage][warp in group][q]
    double *xsum = wsum + FZ_STAGES * FZ_GW * FZ_CBW_MAX;            // [ring][rank][q]
    double *s_vTv = xsum + FZ_XR * FZ_MAXCS * FZ_CBW_MAX;
    double *s_xTv = s_vTv + NBMAX;
    double *s_rowV = s_xTv + NBMAX;
    double *s_rowX = s_rowV + NBMAX;
    double *s_sc = s_rowX + NBMAX;
    double *s_fin = s_sc + 8;                                        // [FZ_NFIN-1][2*NBMAX+2]
    uint64_t *full = reinterpret_cast<uint64_t *>(s_fin + (FZ_NFIN - 1) * (2 * NBMAX + 2));
    uint64_t *empty = full + FZ_STAGES;
    uint64_t *wbar = empty + FZ_STAGES;
    uint64_t *rbar = wbar + FZ_STAGES;
    uint64_t *xbar = rbar + FZ_STAGES;
    int *hn = reinterpret_cast<int *>(xbar + FZ_XR);       // [FZ_STAGES] column counts, [FZ_STAGES] = last flag

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned crank, CS;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(CS));
    const int g = blockIdx.x / CS;                         // cluster index (1-D grid)
    const int i = a.i, k = a.k, nb = a.nb, Lc = a.Lc, NC = a.NC;
    const int rs = (i & ~1) + (int)crank * Lc;             // first row of this CTA (even)
    int len = a.mpad - rs;
    if (len > Lc) len = Lc;
    if (len < 0) len = 0;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            fz_mbar_init(full + s, 2);          // TMA transaction arrive + helper
            fz_mbar_init(empty + s, FZ_GW + 1); // the sweep-2 warps + the finisher (it reads the stage's panel rows)
            fz_mbar_init(wbar + s, FZ_GW);      // the sweep-1 warps of the tile's group
            fz_mbar_init(rbar + s, 1);          // finisher
        }
        for (int x = 0; x < FZ_XR; ++x) fz_mbar_init(xbar + x, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- programmatic dependent launch: this grid may start while its predecessor (finish_xf of the
    //      previous step) drains.  The trailing matrix is not written by that kernel, so the producer
    //      fills the pipeline first; everything the predecessor produces is read after the wait.
    const int ntiles = (a.T > g) ? (a.T - g + NC - 1) / NC : 0;
    int npre = 0;
    if (a.prefetch) {
        npre = ntiles < NST ? ntiles : NST;
        if (tid == 0) {
            for (int nt = 0; nt < npre; ++nt) {
                const int j0 = i + 1 + (g + nt * NC) * CBW;
                int ncols = a.n - j0;
                if (ncols > CBW) ncols = CBW;
                const unsigned bytes = (unsigned)ncols * (unsigned)len * 8u;
                FZ_TR(0, nt);
                if (bytes) {
                    fz_mbar_arrive_expect_tx(full + nt, bytes);
                    for (int q = 0; q < ncols; ++q)
                        fz_bulk_g2s(tile + (size_t)nt * STG + (size_t)q * Lc, a.A + rs + (long)(j0 + q) * a.lda,
                                    (unsigned)len * 8u, full + nt);
                } else {
                    fz_mbar_arrive(full + nt);
                }
            }
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // every global load of the prologue that depends on nothing computed here is issued now, so the
    // three dependent phases below do not each pay a trip to L2
    double pro_pv = 0.0, pro_px = 0.0, pro_ci = 0.0;
    if (tid < k) { pro_pv = a.P[i + (long)tid * a.ldp]; pro_px = a.P[i + (long)(nb + tid) * a.ldp]; }
    if (tid == 32) pro_ci = a.c[i];
    double2 creg[RPT];                                     // sweep-1 warps: their slice of the column c
    if (warp >= FZ_W_S1 && warp < FZ_W_S2) {
        const int gt0 = (warp - FZ_W_S1) * 32 + lane;
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            const int lr = 2 * gt0 + 2 * FZ_GT * u;
            creg[u] = (lr < len) ? *reinterpret_cast<const double2 *>(a.c + rs + lr) : make_double2(0.0, 0.0);
        }
    }
    // helper warps: the panel rows of their FIRST tile are requested now, so that the trip to L2 overlaps the
    // prologue below instead of following it (a step's first tile used to become ready ~3 us after the prologue)
    constexpr bool EARLY = SVDGPU_FZ_EARLY && (CBW <= 2);          // four columns per tile: 20 doubles would spill
    double hy0[EARLY ? CBW : 1][2], hu0[EARLY ? CBW : 1][2], ha0[EARLY ? CBW : 1];
    auto helper_load = [&](int nt, double (&yy)[CBW][2], double (&uu)[CBW][2], double (&aa)[CBW]) {
        const int j0 = i + 1 + (g + nt * NC) * CBW;
        int ncols = a.n - j0;
        if (ncols > CBW) ncols = CBW;
#pragma unroll
        for (int q = 0; q < CBW; ++q) {
            const int j = (q < ncols) ? j0 + q : j0;
            aa[q] = (lane == 0) ? a.A[i + (long)j * a.lda] : 0.0;
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int kk = lane + 32 * z;
                yy[q][z] = (kk < k) ? a.Q[j + (long)kk * a.ldq] : 0.0;
                uu[q][z] = (kk < k) ? a.Q[j + (long)(nb + kk) * a.ldq] : 0.0;
            }
        }
    };
    if constexpr (EARLY) {
        if (warp >= 1 && warp <= FZ_HW && warp - 1 < ntiles) helper_load(warp - 1, hy0, hu0, ha0);
    }
    // ---- combine the partial dots of the current column c (left by finish_xf) in a fixed order;
    //      the panel-row area is still unused and serves as scratch (the tiles may already be in flight)
    {
        constexpr int S1 = 2 * NBMAX + 2;
        // TPE threads per entry, each adding a contiguous chunk of the partials with all of its loads
        // in flight at once (a plain accumulate loop serialises one L2 round trip per few partials)
        const int ne = 2 * k + 1;                         // [0..k) V^T c, [nb..nb+k) X^T c, [2nb] c.c
        const int TPE = (ne <= 65) ? 10 : 5;              // 704 threads: 65 x 10 or 129 x 5
        double *s_part = qrow, *s_d1 = qrow + 10 * S1;
        if (a.nparts1 > 0) {
            const int e = tid / TPE, part = tid - TPE * e;
            if (e < ne) {
                const int slot = (e < k) ? e : (e < 2 * k ? nb + (e - k) : 2 * nb);
                const int chunk = (a.nparts1 + TPE - 1) / TPE, p0 = part * chunk, p1 = min(a.nparts1, p0 + chunk);
                constexpr int LB = 13;
                double a2 = 0.0;
                for (int base = p0; base < p1; base += LB) {
                    double v[LB];
#pragma unroll
                    for (int u = 0; u < LB; ++u) v[u] = (base + u < p1) ? a.dots1[(long)(base + u) * S1 + slot] : 0.0;
#pragma unroll
                    for (int u = 0; u < LB; ++u) a2 += v[u];
                }
                s_part[part * S1 + slot] = a2;
            }
            __syncthreads();
            if (tid < S1 && (tid < k || (tid >= nb && tid < nb + k) || tid == 2 * nb)) {
                double a2 = 0.0;
                for (int pz = 0; pz < TPE; ++pz) a2 += s_part[pz * S1 + tid];
                s_d1[tid] = a2;
            }
        } else {
            if (tid < S1 && (tid < k || (tid >= nb && tid < nb + k) || tid == 2 * nb)) s_d1[tid] = a.dots1[tid];
        }
        __syncthreads();
        if (tid == 32) {
            const double ci = pro_ci;
            Refl f = make_refl(ci, s_d1[2 * nb]);
            s_sc[0] = f.snu; s_sc[1] = f.inv; s_sc[2] = (ci + f.snu) * f.inv;
            if (g == 0 && crank == 0) a.alpha[i] = -f.snu;
        }
        __syncthreads();
        if (tid < k) {
            const double snu0 = s_sc[0], inv0 = s_sc[1];
            const double pv = pro_pv, px = pro_px;
            s_rowV[tid] = pv;
            s_rowX[tid] = px;
            s_vTv[tid] = (s_d1[tid] + snu0 * pv) * inv0;
            s_xTv[tid] = (s_d1[nb + tid] + snu0 * px) * inv0;
        }
        __syncthreads();
    }
    const double snu = s_sc[0], inv = s_sc[1], vi = s_sc[2];
    if (CS > 1) fz_cluster_sync();                         // barriers exist everywhere before remote traffic

    constexpr int S2 = 2 * NBMAX + 2;

    if (warp == 0) {
        // ============================ TMA producer warp ============================
        for (int nt = npre; nt < ntiles; ++nt) {
            const int s = nt % NST;
            fz_mbar_wait(empty + s, ((nt / NST) & 1) ^ 1);
            if (lane == 0) {
                const int j0 = i + 1 + (g + nt * NC) * CBW;
                int ncols = a.n - j0;
                if (ncols > CBW) ncols = CBW;
                const unsigned bytes = (unsigned)ncols * (unsigned)len * 8u;
                FZ_TR(0, nt);
                if (bytes) {
                    fz_mbar_arrive_expect_tx(full + s, bytes);
                    for (int q = 0; q < ncols; ++q)
                        fz_bulk_g2s(tile + (size_t)s * STG + (size_t)q * Lc, a.A + rs + (long)(j0 + q) * a.lda,
                                    (unsigned)len * 8u, full + s);
                } else {
                    fz_mbar_arrive(full + s);
                }
            }
        }
    } else if (warp <= FZ_HW) {
        // ============================== helper warps ==============================
        // per tile column j: a_ij, corr_j = Y[j,:].vTv + U[j,:].xTv, g_j = rowV.Y[j,:] + rowX.U[j,:]
        // and a copy of the panel rows for the dot partials; all loads are issued before any use
        for (int nt = warp - 1; nt < ntiles; nt += FZ_HW) {
            const int s = nt % NST;
            const int j0 = i + 1 + (g + nt * NC) * CBW;
            int ncols = a.n - j0;
            if (ncols > CBW) ncols = CBW;
            double yk[CBW][2], uk[CBW][2], aij[CBW];
            if (EARLY && nt == warp - 1) {
                if constexpr (EARLY) {
#pragma unroll
                    for (int q = 0; q < CBW; ++q) {
                        aij[q] = ha0[q];
#pragma unroll
                        for (int z = 0; z < 2; ++z) { yk[q][z] = hy0[q][z]; uk[q][z] = hu0[q][z]; }
                    }
                }
            } else {
                helper_load(nt, yk, uk, aij);
            }
            fz_mbar_wait(empty + s, ((nt / NST) & 1) ^ 1);     // stage (and its helper slots) free
#pragma unroll
            for (int q = 0; q < CBW; ++q) {
                double corr = 0.0, gg = 0.0;
                double *qr = qrow + (size_t)(s * FZ_CBW_MAX + q) * 2 * NBMAX;
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    const int kk = lane + 32 * z;
                    if (kk < k) {
                        corr += yk[q][z] * s_vTv[kk] + uk[q][z] * s_xTv[kk];
                        gg += s_rowV[kk] * yk[q][z] + s_rowX[kk] * uk[q][z];
                        qr[kk] = yk[q][z];
                        qr[NBMAX + kk] = uk[q][z];
                    }
                }
                corr = warp_sum(corr);
                gg = warp_sum(gg);
                if (lane == 0) {
                    hcorr[s * FZ_CBW_MAX + q] = corr; hg[s * FZ_CBW_MAX + q] = gg; haij[s * FZ_CBW_MAX + q] = aij[q];
                }
            }
            if (lane == 0) hn[s] = ncols;
            __syncwarp();
            if (lane == 0) fz_mbar_arrive(full + s);
        }
    } else if (warp == FZ_W_RED) {
        // ============================== reducer warp ==============================
        // combines the sweep-1 warps' column sums of a tile and ships them to every CTA of the cluster
        for (int nt = 0; nt < ntiles; ++nt) {
            const int s = nt % NST, xs = nt % FZ_XR;
            fz_mbar_wait(wbar + s, (nt / NST) & 1);
            if (lane == 0) FZ_TR(3, nt);
            double vals[CBW];
#pragma unroll
            for (int qq = 0; qq < CBW; ++qq) {
                // FZ_WP partials per sweep-1 warp: 8 * FZ_WP values per column
                double v = (lane < FZ_GW * WP) ? wsum[s * (FZ_GW * FZ_CBW_MAX) + qq * (FZ_GW * WP) + lane] : 0.0;
#pragma unroll
                for (int off = FZ_GW * WP / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                vals[qq] = __shfl_sync(0xffffffffu, v, 0);
            }
            if (CS == 1) {
                if (lane == 0) {
#pragma unroll
                    for (int qq = 0; qq < CBW; ++qq) xsum[(xs * FZ_MAXCS) * FZ_CBW_MAX + qq] = vals[qq];
                    fz_mbar_arrive(xbar + xs);
                }
            } else {
                if (lane == 0) fz_mbar_arrive_expect_tx(xbar + xs, CS * CBW * 8u);
                if (lane < (int)CS) {
                    const unsigned base =
                        fz_mapa(fz_smem_u32(xsum + (xs * FZ_MAXCS + (int)crank) * FZ_CBW_MAX), (unsigned)lane);
                    const unsigned rb = fz_mapa(fz_smem_u32(xbar + xs), (unsigned)lane);
#pragma unroll
                    for (int qq = 0; qq < CBW; ++qq) fz_st_async(base + 8u * qq, vals[qq], rb);
                }
            }
        }
    } else if (warp < FZ_W_S1) {
        // ============================== finisher warps ==============================
        // (the per-tile latency chain of this role — barrier wake-up, smem reads, two stores, the
        //  panel dots — is longer than the tile period, so tiles alternate between FZ_NFIN warps)
        // y_j, r_j of every tile column once the cluster-wide column sums are in; rank 0 also stores
        // them (new Y column, row vector r) and accumulates the panel dots Y^T r, U^T r, r.r
        double dY[2] = {0.0, 0.0}, dU[2] = {0.0, 0.0}, rr2 = 0.0, yr = 0.0;
        const int fin = warp - FZ_W_FIN;
        for (int pt = fin; pt < ntiles; pt += FZ_NFIN) {
            const int s = pt % NST, xs = pt % FZ_XR;
            fz_mbar_wait(xbar + xs, (pt / FZ_XR) & 1);
            if (lane == 0) FZ_TR(4, pt);
            const int ncols = hn[s];
            const int j0 = i + 1 + (g + pt * NC) * CBW;
            double y = 0.0, r = 0.0;
            if (lane < ncols) {
                double tsum = 0.0;
                for (unsigned rk = 0; rk < CS; ++rk) tsum += xsum[(xs * FZ_MAXCS + (int)rk) * FZ_CBW_MAX + lane];
                const double aij = haij[s * FZ_CBW_MAX + lane];
                y = 2.0 * ((tsum + snu * aij) * inv - hcorr[s * FZ_CBW_MAX + lane]);
                r = aij - hg[s * FZ_CBW_MAX + lane] - vi * y;
                rq[s * FZ_CBW_MAX + lane] = r;
            }
            __syncwarp();
            if (lane == 0) FZ_TR(5, pt);
            if (lane == 0) fz_mbar_arrive(rbar + s);
            if (crank == 0) {
                if (lane < ncols) {
                    a.Q[(j0 + lane) + (long)k * a.ldq] = y;
                    a.rv[j0 + lane] = r;
                }
#pragma unroll
                for (int q = 0; q < CBW; ++q) {
                    const double rb = __shfl_sync(0xffffffffu, r, q), yb = __shfl_sync(0xffffffffu, y, q);
                    if (q < ncols) {
                        const double *qr = qrow + (size_t)(s * FZ_CBW_MAX + q) * 2 * NBMAX;
#pragma unroll
                        for (int z = 0; z < 2; ++z) {
                            const int kk = lane + 32 * z;
                            if (kk < k) { dY[z] += qr[kk] * rb; dU[z] += qr[NBMAX + kk] * rb; }
                        }
                        rr2 += rb * rb;
                        yr += yb * rb;
                    }
                }
            }
            // the stage's helper slots (qrow) are free for the next tile only now
            __syncwarp();
            if (lane == 0) fz_mbar_arrive(empty + s);
        }
        // combine the finishers' partial panel dots in a fixed order: the others park theirs in
        // shared memory, finisher 0 adds them up and writes the cluster's partial
        if (fin > 0) {
            double *sp = s_fin + (fin - 1) * S2;
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int kk = lane + 32 * z;
                if (kk < k) { sp[kk] = dY[z]; sp[nb + kk] = dU[z]; }
            }
            if (lane == 0) { sp[k] = yr; sp[2 * nb] = rr2; }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(FZ_NFIN * 32) : "memory");
        if (fin == 0 && crank == 0) {
            for (int f2 = 1; f2 < FZ_NFIN; ++f2) {
                const double *sp = s_fin + (f2 - 1) * S2;
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    const int kk = lane + 32 * z;
                    if (kk < k) { dY[z] += sp[kk]; dU[z] += sp[nb + kk]; }
                }
                if (lane == 0) { yr += sp[k]; rr2 += sp[2 * nb]; }
            }
            double *out = a.dots2p + (long)g * S2;
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int kk = lane + 32 * z;
                if (kk < k) { out[kk] = dY[z]; out[nb + kk] = dU[z]; }
            }
            if (lane == 0) { out[k] = yr; out[2 * nb] = rr2; }
        }
    } else if (warp < FZ_W_S2) {
        // ============================== sweep-1 warps ==============================
        // column dots with c: t1_j = sum_r A[r,j] c[r] over this CTA's rows; c lives in registers.
        // No barrier between the warps: each one runs ahead to the next tile as soon as it has landed.
        const int wig = warp - FZ_W_S1;
        const int gt = wig * 32 + lane;
        if (g == 0) {
            // the reflector itself: v = (c + s*nu*e_i) * inv, in place and into the V panel
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
                const int lr = 2 * gt + 2 * FZ_GT * u;
                if (lr < len) {
                    const int r = rs + lr;
                    if (r >= i && r < a.m) {
                        const double v = (creg[u].x + (r == i ? snu : 0.0)) * inv;
                        a.A[r + (long)i * a.lda] = v;
                        a.P[r + (long)k * a.ldp] = v;
                    }
                    if (r + 1 >= i && r + 1 < a.m) {
                        const double v = (creg[u].y + (r + 1 == i ? snu : 0.0)) * inv;
                        a.A[r + 1 + (long)i * a.lda] = v;
                        a.P[r + 1 + (long)k * a.ldp] = v;
                    }
                }
            }
        }
        for (int nt = 0; nt < ntiles; ++nt) {
            const int s = nt % NST;
            fz_mbar_wait(full + s, (nt / NST) & 1);
            if (wig == 0 && lane == 0) FZ_TR(1, nt);
            const int ncols = hn[s];
            const double *tl = tile + (size_t)s * STG;
#pragma unroll
            for (int q = 0; q < CBW; ++q) {
                const double pq = (q < ncols) ? fz_col_dot<RPT>(tl + (size_t)q * Lc, creg, gt, len) : 0.0;
                // the last log2(WP) rounds of the lane reduction are left to the reducer warp (it has slack, this warp is
                // the pipeline's critical role): lanes 0..WP-1 hold the partial sums of the lanes congruent to them mod WP
                double ps = pq;
#pragma unroll
                for (int off = 16; off >= WP; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
                if (lane < WP) wsum[s * (FZ_GW * FZ_CBW_MAX) + q * (FZ_GW * WP) + wig * WP + lane] = ps;
            }
            __syncwarp();
            if (wig == 0 && lane == 0) FZ_TR(2, nt);
            if (lane == 0) fz_mbar_arrive(wbar + s);
        }
    } else {
        // ============================== sweep-2 warps ==============================
        // row dots with r: t2[r] += sum_j A[r,j] r_j; the accumulators live in registers
        const int wig = warp - FZ_W_S2;
        const int gt = wig * 32 + lane;
        double2 acc[RPT];
#pragma unroll
        for (int u = 0; u < RPT; ++u) acc[u] = make_double2(0.0, 0.0);
        for (int nt = 0; nt < ntiles; ++nt) {
            const int s = nt % NST;
            fz_mbar_wait(rbar + s, (nt / NST) & 1);
            if (wig == 0 && lane == 0) FZ_TR(6, nt);
            const int ncols = hn[s];
            const double *tl = tile + (size_t)s * STG;
#pragma unroll
            for (int q = 0; q < CBW; ++q) {
                if (q < ncols) fz_col_axpy<RPT>(tl + (size_t)q * Lc, rq[s * FZ_CBW_MAX + q], acc, gt, len);
            }
            __syncwarp();
            if (wig == 0 && lane == 0) FZ_TR(7, nt);
            if (lane == 0) fz_mbar_arrive(empty + s);
        }
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            const int lr = 2 * gt + 2 * FZ_GT * u;
            if (lr < len) *reinterpret_cast<double2 *>(a.tmpN + (long)g * a.ldt + rs + lr) = acc[u];
        }
    }

    if (CS > 1) fz_cluster_sync();      // nobody exits while a peer may still write into its shared memory
}

} // namespace svdgpu
