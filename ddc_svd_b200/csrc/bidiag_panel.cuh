// bidiag_panel.cuh — one PERSISTENT kernel per panel of the bidiagonalization (included by bidiag.cu).
//
// For trailing blocks of at most 4096 rows (the fused pass runs single-CTA "clusters" there) a step is a
// 20-70 MB stream, i.e. 4-12 us of HBM time, while the two launches of a step (fused pass, finish_xf) cost
// ~19 us of fixed latency: pipeline fill and drain, launch, prologues.  This kernel runs all the steps of one
// panel (nb = 32) in ONE cooperative launch:
//
//     for every step of the panel:
//         pass    - the single-read fused pass of bidiag_fused.cuh (same warp roles, same arithmetic order)
//         grid barrier
//         finish  - x, c', the row reflector u and the next step's partial dots (finish_xf's arithmetic),
//                   32 rows per CTA, the remaining CTAs write u
//         grid barrier
//
// What it buys: (1) the TMA producer warp never joins a barrier - inside a panel the trailing matrix is not
// written (the updates are deferred to the panel GEMM), so it keeps streaming the tiles of step i+1 while the
// other warps are in the barrier / finish of step i: the pipeline never drains; (2) no launches inside a
// panel; (3) the mbarriers, the stage ring and the per-CTA set-up live across steps.
// The mbarrier phases simply keep counting: every role numbers its tiles with a running index (tb + nt).
// Everything another CTA produced is read with ld.global.cg (__ldcg): L1 is not coherent inside one launch.
// Results are bit-identical to the two-kernel path (same partials, same fixed summation order) except for
// the 16- instead of 32-way split of the 2k-term row corrections in the finish.
#pragma once

namespace svdgpu {

struct PanelArgs {
    double *A; long lda;
    int i0, k0, nsteps;           // first step, its column inside the panel, steps in this launch
    int m, n, mpad, nb;
    double *P; long ldp;
    double *Q; long ldq;
    double *c;
    double *rv;
    double *tmpN; long ldt;
    const double *dots1; int nparts1;   // dots of the first step's column (nparts1 == 0: final vector)
    double *dots1p;               // partials written by the finish of every step
    double *dots2p;
    double *alpha, *beta;
    int NC, Lc;
    unsigned *bar;                // grid barrier counter, zero at launch
    unsigned long long *trace;    // measuring aid (SVD_GPU_PPK_TRACE=<first step of a panel>): [step][8] clock64 stamps of CTA 0, else null
};

constexpr bool PPK_DEFAULT_ON = false;  // switched on by SVD_GPU_PPK=1 until measured
constexpr int PK_T = FZ_THREADS - 32;     // threads that take part in the CTA-wide barriers (all but the producer warp)
constexpr int PK_SL = 16;                 // column slices of the finish (warps 1..16)

__device__ __forceinline__ void pk_cta_sync()
{
    asm volatile("bar.sync 2, %0;" ::"n"(PK_T) : "memory");
}
__device__ __forceinline__ void pk_grid_barrier(unsigned *ctr, unsigned target, int t2, unsigned long long *tr)
{
    pk_cta_sync();
    if (t2 == 0) {
        if (tr) tr[0] = clock64();
        __threadfence();
        atomicAdd(ctr, 1u);
        const long long t0 = clock64();
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v < target && clock64() - t0 > 4000000000ll) {
                printf("bidiag panel kernel: grid barrier timed out (block %d, %u of %u)\n", blockIdx.x, v, target);
                asm volatile("trap;");
            }
        } while (v < target);
        __threadfence();
        if (tr) tr[1] = clock64();
    }
    pk_cta_sync();
}

// mbarrier wait with a watchdog: a protocol bug must trap, not hang the device
__device__ __forceinline__ void pk_mbar_wait(uint64_t *bar, unsigned parity)
{
    if (fz_mbar_test(bar, parity)) return;
    const long long t0 = clock64();
    for (;;) {
        unsigned ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(ok) : "r"(fz_smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) {
            printf("bidiag panel kernel: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            asm volatile("trap;");
        }
    }
}

template <int RPT>
__global__ void __launch_bounds__(FZ_THREADS, 1) panel_kernel(const PanelArgs a)
{
    constexpr int CBW = 8 / RPT;
    constexpr int WP = SVDGPU_FZ_WP ? FZ_CBW_MAX / CBW : 1;    // lane partials each sweep-1 warp leaves per column (fills the same 32 slots per stage)
    constexpr int S = 2 * NBMAX + 2;
    extern __shared__ __align__(128) unsigned char fz_smem[];
    double *tile = reinterpret_cast<double *>(fz_smem);
    double *qrow = tile + (size_t)FZ_STAGES * FZ_STAGE;
    double *hcorr = qrow + FZ_STAGES * FZ_CBW_MAX * 2 * NBMAX;
    double *hg = hcorr + FZ_STAGES * FZ_CBW_MAX;
    double *haij = hg + FZ_STAGES * FZ_CBW_MAX;
    double *yq = haij + FZ_STAGES * FZ_CBW_MAX;
    double *rq = yq + FZ_STAGES * FZ_CBW_MAX;
    double *wsum = rq + FZ_STAGES * FZ_CBW_MAX;
    double *xsum = wsum + FZ_STAGES * FZ_GW * FZ_CBW_MAX;
    double *s_vTv = xsum + FZ_XR * FZ_MAXCS * FZ_CBW_MAX;
    double *s_xTv = s_vTv + NBMAX;
    double *s_rowV = s_xTv + NBMAX;
    double *s_rowX = s_rowV + NBMAX;
    double *s_sc = s_rowX + NBMAX;
    double *s_fin = s_sc + 8;
    uint64_t *full = reinterpret_cast<uint64_t *>(s_fin + (FZ_NFIN - 1) * S);
    uint64_t *empty = full + FZ_STAGES;
    uint64_t *wbar = empty + FZ_STAGES;
    uint64_t *rbar = wbar + FZ_STAGES;
    uint64_t *xbar = rbar + FZ_STAGES;
    int *hn = reinterpret_cast<int *>(xbar + FZ_XR);
    // finish scratch: the panel-row area of the pass (idle between the passes), 3072 doubles
    double *f_red = qrow;                          // [3][PK_SL][33]   (1584)
    double *f_part = qrow;                         // [10][S]          (1300), consumed before f_red is written
    double *f_d = qrow + 1600;                     // [S]
    double *f_yTu = f_d + S, *f_uTu = f_yTu + NBMAX, *f_rowY = f_uTu + NBMAX, *f_rowU = f_rowY + NBMAX;
    double *f_c = f_rowU + NBMAX, *f_x = f_c + 32;
    static_assert(1600 + S + 4 * NBMAX + 64 <= FZ_STAGES * FZ_CBW_MAX * 2 * NBMAX, "finish scratch does not fit");
    static_assert(3 * PK_SL * 33 <= 1600 && 10 * S <= 1600, "finish scratch layout");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t2 = tid - 32;                       // index among the non-producer threads
    const int g = blockIdx.x, NC = a.NC, nb = a.nb, Lc = a.Lc;

    if (tid == 0) {
        for (int s = 0; s < FZ_STAGES; ++s) {
            fz_mbar_init(full + s, 2);
            fz_mbar_init(empty + s, FZ_GW + 1);
            fz_mbar_init(wbar + s, FZ_GW);
            fz_mbar_init(rbar + s, 1);
        }
        for (int x = 0; x < FZ_XR; ++x) fz_mbar_init(xbar + x, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int tb = 0;                                    // tiles this CTA has processed in earlier steps
    unsigned nbar = 0;
    const double *d1 = a.dots1;
    int np1 = a.nparts1;

    for (int st = 0; st < a.nsteps; ++st) {
        const int i = a.i0 + st, k = a.k0 + st;
        const int rs = i & ~1;
        int len = a.mpad - rs;
        if (len > Lc) len = Lc;
        const int R = a.n - i - 1, Lb = a.m - i - 1;
        const int T = (R + CBW - 1) / CBW;
        const int ntiles = (T > g) ? (T - g + NC - 1) / NC : 0;

        unsigned long long *tr = (a.trace && g == 0) ? a.trace + st * 8 : nullptr;
        unsigned long long *tt8 = (a.trace && g == 0 && st == 1) ? a.trace + NBMAX * 8 : nullptr;   // per-tile stamps of step 1
#define PK_TR(slot, nt) do { if (tt8 && (nt) < 64) tt8[(nt) * 8 + (slot)] = clock64(); } while (0)
        if (warp == 0) {
            // ============================ TMA producer warp ============================
            for (int nt = 0; nt < ntiles; ++nt) {
                const int gt = tb + nt, s = gt % FZ_STAGES;
                pk_mbar_wait(empty + s, ((gt / FZ_STAGES) & 1) ^ 1);
                if (lane == 0) {
                    PK_TR(0, nt);
                    if (tr && nt == 0) tr[6] = clock64();
                    if (tr && nt == ntiles - 1) tr[7] = clock64();
                    const int j0 = i + 1 + (g + nt * NC) * CBW;
                    int ncols = a.n - j0;
                    if (ncols > CBW) ncols = CBW;
                    const unsigned bytes = (unsigned)ncols * (unsigned)len * 8u;
                    if (bytes) {
                        fz_mbar_arrive_expect_tx(full + s, bytes);
                        for (int q = 0; q < ncols; ++q)
                            fz_bulk_g2s(tile + (size_t)s * FZ_STAGE + (size_t)q * Lc, a.A + rs + (long)(j0 + q) * a.lda,
                                        (unsigned)len * 8u, full + s);
                    } else {
                        fz_mbar_arrive(full + s);
                    }
                }
            }
            tb += ntiles;
            continue;
        }

        // helper warps: the panel rows of their first tile are requested before the prologue (tiles of <= 2 columns:
        // more would not fit the registers), the trip to L2 then overlaps it
        constexpr bool EARLY = SVDGPU_FZ_EARLY && (CBW <= 2);
        double hy[CBW][2], hu[CBW][2], ha[CBW];
        auto helper_load = [&](int nt, double (&yy)[CBW][2], double (&uu)[CBW][2], double (&aa)[CBW]) {
            const int j0 = i + 1 + (g + nt * NC) * CBW;
            int ncols = a.n - j0;
            if (ncols > CBW) ncols = CBW;
#pragma unroll
            for (int q = 0; q < CBW; ++q) {
                const int j = (q < ncols) ? j0 + q : j0;
                aa[q] = (lane == 0) ? __ldcg(a.A + i + (long)j * a.lda) : 0.0;
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    const int kk = lane + 32 * z;
                    yy[q][z] = (kk < k) ? __ldcg(a.Q + j + (long)kk * a.ldq) : 0.0;
                    uu[q][z] = (kk < k) ? __ldcg(a.Q + j + (long)(nb + kk) * a.ldq) : 0.0;
                }
            }
        };
        int hnt0 = ((warp - 1) - tb) % FZ_HW;
        if (hnt0 < 0) hnt0 += FZ_HW;
        if constexpr (EARLY) { if (warp <= FZ_HW && hnt0 < ntiles) helper_load(hnt0, hy, hu, ha); }
        // ------------------------------------------------------------------ pass prologue
        if (tr && t2 == 0) tr[0] = clock64();
        double pro_pv = 0.0, pro_px = 0.0, pro_ci = 0.0;
        if (t2 < k) { pro_pv = __ldcg(a.P + i + (long)t2 * a.ldp); pro_px = __ldcg(a.P + i + (long)(nb + t2) * a.ldp); }
        if (t2 == 32) pro_ci = __ldcg(a.c + i);
        {
            const int ne = 2 * k + 1;
            constexpr int TPE = 10;                        // nb <= 32: at most 65 entries x 10 threads
            double *s_part = qrow, *s_d1 = qrow + 10 * S;
            if (np1 > 0) {
                const int e = t2 / TPE, part = t2 - TPE * e;
                if (e < ne) {
                    const int slot = (e < k) ? e : (e < 2 * k ? nb + (e - k) : 2 * nb);
                    const int chunk = (np1 + TPE - 1) / TPE, p0 = part * chunk, p1 = min(np1, p0 + chunk);
                    constexpr int LB = 13;
                    double a2 = 0.0;
                    for (int base = p0; base < p1; base += LB) {
                        double v[LB];
#pragma unroll
                        for (int u = 0; u < LB; ++u) v[u] = (base + u < p1) ? __ldcg(d1 + (long)(base + u) * S + slot) : 0.0;
#pragma unroll
                        for (int u = 0; u < LB; ++u) a2 += v[u];
                    }
                    s_part[part * S + slot] = a2;
                }
                pk_cta_sync();
                if (t2 < S && (t2 < k || (t2 >= nb && t2 < nb + k) || t2 == 2 * nb)) {
                    double a2 = 0.0;
                    for (int pz = 0; pz < TPE; ++pz) a2 += s_part[pz * S + t2];
                    s_d1[t2] = a2;
                }
            } else {
                if (t2 < S && (t2 < k || (t2 >= nb && t2 < nb + k) || t2 == 2 * nb)) s_d1[t2] = __ldcg(d1 + t2);
            }
            pk_cta_sync();
            if (t2 == 32) {
                const double ci = pro_ci;
                Refl f = make_refl(ci, s_d1[2 * nb]);
                s_sc[0] = f.snu; s_sc[1] = f.inv; s_sc[2] = (ci + f.snu) * f.inv;
                if (g == 0) a.alpha[i] = -f.snu;
            }
            pk_cta_sync();
            if (t2 < k) {
                const double snu0 = s_sc[0], inv0 = s_sc[1];
                s_rowV[t2] = pro_pv;
                s_rowX[t2] = pro_px;
                s_vTv[t2] = (s_d1[t2] + snu0 * pro_pv) * inv0;
                s_xTv[t2] = (s_d1[nb + t2] + snu0 * pro_px) * inv0;
            }
            pk_cta_sync();
        }
        const double snu = s_sc[0], inv = s_sc[1], vi = s_sc[2];
        if (tr && t2 == 0) tr[1] = clock64();

        // ------------------------------------------------------------------ pass roles
        if (warp <= FZ_HW) {
            // helper warps: tile gt is prepared by helper gt % FZ_HW.  The panel rows of a tile's columns come from
            // L2 (~1.5k cycles): the loads of the helper's NEXT tile are issued before the current one is worked on
            // otherwise two helpers cannot keep up with the stream.
            if constexpr (!EARLY) { if (hnt0 < ntiles) helper_load(hnt0, hy, hu, ha); }
            int nt = hnt0;
            while (nt < ntiles) {
                const int gt = tb + nt, s = gt % FZ_STAGES;
                const int j0 = i + 1 + (g + nt * NC) * CBW;
                int ncols = a.n - j0;
                if (ncols > CBW) ncols = CBW;
                double yn[CBW][2], un[CBW][2], an[CBW];
                const int ntn = nt + FZ_HW;
                if (ntn < ntiles) helper_load(ntn, yn, un, an);
                pk_mbar_wait(empty + s, ((gt / FZ_STAGES) & 1) ^ 1);
#pragma unroll
                for (int q = 0; q < CBW; ++q) {
                    double corr = 0.0, gg = 0.0;
                    double *qr = qrow + (size_t)(s * FZ_CBW_MAX + q) * 2 * NBMAX;
#pragma unroll
                    for (int z = 0; z < 2; ++z) {
                        const int kk = lane + 32 * z;
                        if (kk < k) {
                            corr += hy[q][z] * s_vTv[kk] + hu[q][z] * s_xTv[kk];
                            gg += s_rowV[kk] * hy[q][z] + s_rowX[kk] * hu[q][z];
                            qr[kk] = hy[q][z];
                            qr[NBMAX + kk] = hu[q][z];
                        }
                    }
                    corr = warp_sum(corr);
                    gg = warp_sum(gg);
                    if (lane == 0) {
                        hcorr[s * FZ_CBW_MAX + q] = corr; hg[s * FZ_CBW_MAX + q] = gg; haij[s * FZ_CBW_MAX + q] = ha[q];
                    }
                }
                if (lane == 0) hn[s] = ncols;
                __syncwarp();
                if (lane == 0) fz_mbar_arrive(full + s);
#pragma unroll
                for (int q = 0; q < CBW; ++q) {
                    ha[q] = an[q];
#pragma unroll
                    for (int z = 0; z < 2; ++z) { hy[q][z] = yn[q][z]; hu[q][z] = un[q][z]; }
                }
                nt = ntn;
            }
        } else if (warp == FZ_W_RED) {
            for (int nt = 0; nt < ntiles; ++nt) {
                const int gt = tb + nt, s = gt % FZ_STAGES, xs = gt % FZ_XR;
                pk_mbar_wait(wbar + s, (gt / FZ_STAGES) & 1);
                if (lane == 0) PK_TR(3, nt);
                double vals[CBW];
#pragma unroll
                for (int qq = 0; qq < CBW; ++qq) {
                    // FZ_WP partials per sweep-1 warp: 8 * FZ_WP values per column
                    double v = (lane < FZ_GW * WP) ? wsum[s * (FZ_GW * FZ_CBW_MAX) + qq * (FZ_GW * WP) + lane] : 0.0;
#pragma unroll
                    for (int off = FZ_GW * WP / 2; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    vals[qq] = __shfl_sync(0xffffffffu, v, 0);
                }
                if (lane == 0) {
#pragma unroll
                    for (int qq = 0; qq < CBW; ++qq) xsum[(xs * FZ_MAXCS) * FZ_CBW_MAX + qq] = vals[qq];
                    fz_mbar_arrive(xbar + xs);
                }
            }
        } else if (warp < FZ_W_S1) {
            // finisher warps: tile gt is finished by finisher gt % FZ_NFIN
            double dY[2] = {0.0, 0.0}, dU[2] = {0.0, 0.0}, rr2 = 0.0, yr = 0.0;
            const int fin = warp - FZ_W_FIN;
            int pt = (fin - tb) % FZ_NFIN;
            if (pt < 0) pt += FZ_NFIN;
            for (; pt < ntiles; pt += FZ_NFIN) {
                const int gt = tb + pt, s = gt % FZ_STAGES, xs = gt % FZ_XR;
                pk_mbar_wait(xbar + xs, (gt / FZ_XR) & 1);
                if (lane == 0) PK_TR(4, pt);
                const int ncols = hn[s];
                const int j0 = i + 1 + (g + pt * NC) * CBW;
                double y = 0.0, r = 0.0;
                if (lane < ncols) {
                    const double tsum = xsum[(xs * FZ_MAXCS) * FZ_CBW_MAX + lane];
                    const double aij = haij[s * FZ_CBW_MAX + lane];
                    y = 2.0 * ((tsum + snu * aij) * inv - hcorr[s * FZ_CBW_MAX + lane]);
                    r = aij - hg[s * FZ_CBW_MAX + lane] - vi * y;
                    rq[s * FZ_CBW_MAX + lane] = r;
                }
                __syncwarp();
                if (lane == 0) PK_TR(5, pt);
                if (lane == 0) fz_mbar_arrive(rbar + s);
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
                __syncwarp();
                if (lane == 0) fz_mbar_arrive(empty + s);
            }
            // the fixed combination order of the two-kernel path: finisher 0 + finisher 1
            // (which finisher saw which tile depends on tb's parity; keep "the one that took the first tile" first)
            const int first = tb % FZ_NFIN;                // the finisher that took this step's tile 0
            const bool lead = (fin == first);
            if (!lead) {
                double *sp = s_fin;
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    const int kk = lane + 32 * z;
                    if (kk < k) { sp[kk] = dY[z]; sp[nb + kk] = dU[z]; }
                }
                if (lane == 0) { sp[k] = yr; sp[2 * nb] = rr2; }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(FZ_NFIN * 32) : "memory");
            if (lead) {
                const double *sp = s_fin;
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    const int kk = lane + 32 * z;
                    if (kk < k) { dY[z] += sp[kk]; dU[z] += sp[nb + kk]; }
                }
                if (lane == 0) { yr += sp[k]; rr2 += sp[2 * nb]; }
                double *out = a.dots2p + (long)g * S;
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    const int kk = lane + 32 * z;
                    if (kk < k) { out[kk] = dY[z]; out[nb + kk] = dU[z]; }
                }
                if (lane == 0) { out[k] = yr; out[2 * nb] = rr2; }
            }
        } else if (warp < FZ_W_S2) {
            const int wig = warp - FZ_W_S1;
            const int gt0 = wig * 32 + lane;
            // this warp's slice of the column c (loaded here, not before the prologue: the first tile is not ready
            // before the helpers have been to L2 anyway, and the registers stay free during the prologue)
            double2 creg[RPT];
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
                const int lr = 2 * gt0 + 2 * FZ_GT * u;
                creg[u] = (lr < len) ? __ldcg(reinterpret_cast<const double2 *>(a.c + rs + lr)) : make_double2(0.0, 0.0);
            }
            if (g == 0) {
#pragma unroll
                for (int u = 0; u < RPT; ++u) {
                    const int lr = 2 * gt0 + 2 * FZ_GT * u;
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
                const int gt = tb + nt, s = gt % FZ_STAGES;
                pk_mbar_wait(full + s, (gt / FZ_STAGES) & 1);
                if (wig == 0 && lane == 0) PK_TR(1, nt);
                const int ncols = hn[s];
                const double *tl = tile + (size_t)s * FZ_STAGE;
#pragma unroll
                for (int q = 0; q < CBW; ++q) {
                    const double pq = (q < ncols) ? fz_col_dot<RPT>(tl + (size_t)q * Lc, creg, gt0, len) : 0.0;
                    // the last log2(WP) rounds of the lane reduction are left to the reducer warp (it has slack, this warp is
                    // the pipeline's critical role): lanes 0..WP-1 hold the partial sums of the lanes congruent to them mod WP
                    double ps = pq;
#pragma unroll
                    for (int off = 16; off >= WP; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
                    if (lane < WP) wsum[s * (FZ_GW * FZ_CBW_MAX) + q * (FZ_GW * WP) + wig * WP + lane] = ps;
                }
                __syncwarp();
                if (wig == 0 && lane == 0) PK_TR(2, nt);
                if (lane == 0) fz_mbar_arrive(wbar + s);
            }
        } else {
            const int wig = warp - FZ_W_S2;
            const int gt0 = wig * 32 + lane;
            double2 acc[RPT];
#pragma unroll
            for (int u = 0; u < RPT; ++u) acc[u] = make_double2(0.0, 0.0);
            for (int nt = 0; nt < ntiles; ++nt) {
                const int gt = tb + nt, s = gt % FZ_STAGES;
                pk_mbar_wait(rbar + s, (gt / FZ_STAGES) & 1);
                if (wig == 0 && lane == 0) PK_TR(6, nt);
                const int ncols = hn[s];
                const double *tl = tile + (size_t)s * FZ_STAGE;
#pragma unroll
                for (int q = 0; q < CBW; ++q) {
                    if (q < ncols) fz_col_axpy<RPT>(tl + (size_t)q * Lc, rq[s * FZ_CBW_MAX + q], acc, gt0, len);
                }
                __syncwarp();
                if (wig == 0 && lane == 0) PK_TR(7, nt);
                if (lane == 0) fz_mbar_arrive(empty + s);
            }
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
                const int lr = 2 * gt0 + 2 * FZ_GT * u;
                if (lr < len) *reinterpret_cast<double2 *>(a.tmpN + (long)g * a.ldt + rs + lr) = acc[u];
            }
        }
        tb += ntiles;

        // ------------------------------------------------------------------ everybody's partials are out
        nbar += 1;
        pk_grid_barrier(a.bar, nbar * (unsigned)NC, t2, tr ? tr + 2 : nullptr);

        // ------------------------------------------------------------------ finish: x, c', u, dots of c'
        const int nRowBlk = (Lb + 31) / 32;
        {
            const int w2 = t2 >> 5;
            const bool rowblk = g < nRowBlk;
            const int idx = g * 32 + lane;
            const bool live = rowblk && w2 < PK_SL && idx < Lb;
            double vk[2], xk[2], tt = 0.0, ar = 0.0;
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int q = w2 + PK_SL * z;
                vk[z] = (live && q <= k) ? __ldcg(a.P + (i + 1 + idx) + (long)q * a.ldp) : 0.0;
                xk[z] = (live && q < k) ? __ldcg(a.P + (i + 1 + idx) + (long)(nb + q) * a.ldp) : 0.0;
            }
            if (live) {
                double tv[10];
#pragma unroll
                for (int u = 0; u < 10; ++u) {
                    const int sp = w2 + u * PK_SL;
                    tv[u] = (sp < NC) ? __ldcg(a.tmpN + (long)sp * a.ldt + i + 1 + idx) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 10; ++u) tt += tv[u];
                if (w2 == 0) ar = __ldcg(a.A + (i + 1 + idx) + (long)(i + 1) * a.lda);
            }
            const double rf = __ldcg(a.rv + i + 1);
            double qy = 0.0, qu = 0.0;
            if (t2 <= k) qy = __ldcg(a.Q + (i + 1) + (long)t2 * a.ldq);
            if (t2 < k) qu = __ldcg(a.Q + (i + 1) + (long)(nb + t2) * a.ldq);
            {
                const int ne = 2 * k + 2;
                constexpr int TPE = 10;
                const int e = t2 / TPE, part = t2 - TPE * e;
                if (e < ne) {
                    const int slot = (e <= k) ? e : (e <= 2 * k ? nb + (e - k - 1) : 2 * nb);
                    const int chunk = (NC + TPE - 1) / TPE, p0 = part * chunk, p1 = min(NC, p0 + chunk);
                    constexpr int LB = 8;
                    double a2 = 0.0;
                    for (int base = p0; base < p1; base += LB) {
                        double v[LB];
#pragma unroll
                        for (int u = 0; u < LB; ++u) v[u] = (base + u < p1) ? __ldcg(a.dots2p + (long)(base + u) * S + slot) : 0.0;
#pragma unroll
                        for (int u = 0; u < LB; ++u) a2 += v[u];
                    }
                    f_part[part * S + slot] = a2;
                }
            }
            pk_cta_sync();
            if (t2 < S && (t2 <= k || (t2 >= nb && t2 < nb + k) || t2 == 2 * nb)) {
                double a2 = 0.0;
                for (int pz = 0; pz < 10; ++pz) a2 += f_part[pz * S + t2];
                f_d[t2] = a2;
            }
            pk_cta_sync();
            const Refl f = make_refl(rf, f_d[2 * nb]);
            const double ufirst = (rf + f.snu) * f.inv;
            if (!rowblk) {
                // the row reflector itself, by the CTAs that own no rows
                const int ncb = NC - nRowBlk;
                for (int cidx = (g - nRowBlk) * PK_T + t2; cidx < R; cidx += ncb * PK_T) {
                    const int j = i + 1 + cidx;
                    const double u = (__ldcg(a.rv + j) + (cidx == 0 ? f.snu : 0.0)) * f.inv;
                    a.A[i + (long)j * a.lda] = u;
                    a.Q[j + (long)(nb + k) * a.ldq] = u;
                }
                if (g == nRowBlk && t2 == 0) a.beta[i] = -f.snu;
            } else {
                if (t2 <= k) { f_rowY[t2] = qy; f_yTu[t2] = (f_d[t2] + f.snu * qy) * f.inv; }
                if (t2 < k) { f_rowU[t2] = qu; f_uTu[t2] = (f_d[nb + t2] + f.snu * qu) * f.inv; }
                pk_cta_sync();
                double corr = 0.0, sub = 0.0;
                if (w2 < PK_SL) {
#pragma unroll
                    for (int z = 0; z < 2; ++z) {
                        const int q = w2 + PK_SL * z;
                        if (q <= k) { corr += vk[z] * f_yTu[q]; sub += vk[z] * f_rowY[q]; }
                        if (q < k) { corr += xk[z] * f_uTu[q]; sub += xk[z] * f_rowU[q]; }
                    }
                    f_red[(0 * PK_SL + w2) * 33 + lane] = corr;
                    f_red[(1 * PK_SL + w2) * 33 + lane] = sub;
                    f_red[(2 * PK_SL + w2) * 33 + lane] = tt;
                }
                pk_cta_sync();
                double cc2 = 0.0;
                if (w2 == 0) {
                    double cc = 0.0, x = 0.0;
                    if (live) {
                        corr = 0.0; sub = 0.0; tt = 0.0;
#pragma unroll
                        for (int z = 0; z < PK_SL; ++z) {
                            corr += f_red[(0 * PK_SL + z) * 33 + lane];
                            sub += f_red[(1 * PK_SL + z) * 33 + lane];
                            tt += f_red[(2 * PK_SL + z) * 33 + lane];
                        }
                        const int r = i + 1 + idx;
                        x = 2.0 * ((tt + f.snu * ar) * f.inv - corr);
                        a.P[r + (long)(nb + k) * a.ldp] = x;
                        cc = ar - sub - x * ufirst;
                        a.c[r] = cc;
                    }
                    f_c[lane] = cc;
                    f_x[lane] = x;
                    cc2 = cc * cc;
                    if (g == 0 && lane == 0) a.c[i] = 0.0;
                }
                pk_cta_sync();
                double *out = a.dots1p + (long)g * S;
                if (w2 < PK_SL) {
                    const double cl = f_c[lane];
#pragma unroll
                    for (int z = 0; z < 2; ++z) {
                        const int q = w2 + PK_SL * z;
                        if (q <= k) {                           // warp-uniform
                            const double xv = (q < k) ? xk[z] : f_x[lane];
                            const double a2 = warp_sum(vk[z] * cl), b2 = warp_sum(xv * cl);
                            if (lane == 0) { out[q] = a2; out[nb + q] = b2; }
                        }
                    }
                    if (w2 == 0) {
                        const double a2 = warp_sum(cc2);
                        if (lane == 0) out[2 * nb] = a2;
                    }
                }
            }
        }
        d1 = a.dots1p;
        np1 = nRowBlk;
        if (st + 1 < a.nsteps) {
            nbar += 1;
            pk_grid_barrier(a.bar, nbar * (unsigned)NC, t2, tr ? tr + 4 : nullptr);
        }
    }
}

} // namespace svdgpu
