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
//   * four helper warps (one tile in four each) gather, while the copies fly, the tile's panel
//     rows (Y[j,:], U[j,:]) and reduce the 2k-term corrections the consumers will need;
//   * 16 consumer warps own fixed rows (c and the t2 accumulators live in registers for the
//     whole pass) and read the tile from shared memory twice with conflict-free 128-bit loads;
//   * when a column is taller than one stage can hold (4096 rows per CTA) the rows are split
//     over a thread-block CLUSTER of 2/4/8 CTAs; the per-column partial sums of sweep 1 are
//     exchanged through distributed shared memory (st.async ... mbarrier::complete_tx on the
//     remote barrier, so receivers need no cluster-scope acquire), software-pipelined one tile ahead so the exchange latency hides behind sweep 2
//     of the previous tile;
//   * clusters walk the column tiles round-robin; every cluster writes its partial t2 and its
//     partial panel dots; the last cluster to finish combines the dot partials in a fixed
//     order (deterministic, no floating-point atomics).
#pragma once

namespace svdgpu {

constexpr int FZ_STAGE = 4096;            // doubles per shared-memory stage (32 KB)
constexpr int FZ_STAGES = 6;              // FZ_D+1 tiles in consumption (pipelined exchange), the rest in flight
constexpr int FZ_CBW_MAX = 4;             // columns per tile (rows per CTA <= 1024 -> 4 columns)
constexpr int FZ_D = 2;                   // sweep 2 runs FZ_D tiles behind sweep 1 (hides the exchange + CTA skew)
constexpr int FZ_XR = 8;                  // ring depth of the cross-CTA exchange (>= 2*FZ_D + 2)
constexpr int FZ_CW = 16;                 // consumer warps
constexpr int FZ_CT = FZ_CW * 32;         // consumer threads
constexpr int FZ_HW = 4;                  // helper warps (tile n is prepared by helper n % FZ_HW)
constexpr int FZ_PRE = (1 + FZ_HW) * 32;  // threads before the consumers: TMA producer warp + helpers
constexpr int FZ_THREADS = FZ_PRE + FZ_CT;
constexpr int FZ_MAXCS = 8;               // largest cluster
constexpr int FZ_MAX_CLUSTERS = 148;
constexpr bool FZ_DEFAULT_ON = false;      // see profiles/: flipped when the fused pass beats the split passes
constexpr int FZ_MIN_ROWS = 1024;         // below this trailing height the split passes are used
constexpr int FZ_MIN_COLS = 64;

constexpr size_t FZ_SMEM_DOUBLES = (size_t)FZ_STAGES * FZ_STAGE       // tiles
                                   + (size_t)FZ_STAGES * FZ_CBW_MAX * 2 * NBMAX // panel rows of the tile columns
                                   + 3 * FZ_STAGES * 8                 // corr, g, a_ij
                                   + 2 * FZ_CW * 8                     // per-warp column sums
                                   + FZ_XR * FZ_MAXCS * 8              // exchanged column sums
                                   + 4 * NBMAX + 8;                    // vTv, xTv, rowV, rowX, scalars
constexpr size_t FZ_SMEM_BYTES = FZ_SMEM_DOUBLES * 8 + (2 * FZ_STAGES + FZ_XR) * 8 + 8 * sizeof(int) + 128;

struct FusedArgs {
    double *A; long lda;          // trailing matrix (read) and column i (the reflector is written in place)
    int i, m, n, mpad, k, nb;
    double *P; long ldp;
    double *Q; long ldq;
    const double *c;
    double *rv;
    double *tmpN; long ldt;       // [cluster][row]: partial A r
    const double *dots1;          // final [V^T c | X^T c | c.c]
    double *dots2p;               // [cluster][2nb+2] partials
    double *dots2;                // final [Y^T r | U^T r | r.r]
    unsigned *counter;
    double *alpha;
    int T, NC, Lc;                // column tiles, clusters, rows per CTA (even)
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
__device__ __forceinline__ void fz_consumer_bar()
{
    asm volatile("bar.sync 1, %0;" ::"n"(FZ_CT) : "memory");
}

// RPT = row pairs per consumer thread (rows per CTA <= 1024*RPT <= 4096), CBW = 4/RPT columns per tile
template <int RPT>
__global__ void __launch_bounds__(FZ_THREADS, 1) fused_pass_kernel(const FusedArgs a)
{
    constexpr int CBW = FZ_CBW_MAX / RPT;
    extern __shared__ __align__(128) unsigned char fz_smem[];
    double *tile = reinterpret_cast<double *>(fz_smem);
    double *qrow = tile + (size_t)FZ_STAGES * FZ_STAGE;
    double *hcorr = qrow + FZ_STAGES * FZ_CBW_MAX * 2 * NBMAX;
    double *hg = hcorr + FZ_STAGES * 8;
    double *haij = hg + FZ_STAGES * 8;
    double *wsum = haij + FZ_STAGES * 8;
    double *xsum = wsum + 2 * FZ_CW * 8;
    double *s_vTv = xsum + FZ_XR * FZ_MAXCS * 8;
    double *s_xTv = s_vTv + NBMAX;
    double *s_rowV = s_xTv + NBMAX;
    double *s_rowX = s_rowV + NBMAX;
    double *s_sc = s_rowX + NBMAX;
    uint64_t *full = reinterpret_cast<uint64_t *>(s_sc + 8);
    uint64_t *empty = full + FZ_STAGES;
    uint64_t *xbar = empty + FZ_STAGES;
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
        for (int s = 0; s < FZ_STAGES; ++s) { fz_mbar_init(full + s, 2); fz_mbar_init(empty + s, FZ_CW); }
        for (int x = 0; x < FZ_XR; ++x) fz_mbar_init(xbar + x, 1);
        hn[FZ_STAGES] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid == FZ_PRE) {
        const double ci = a.c[i];
        Refl f = make_refl(ci, a.dots1[2 * nb]);
        s_sc[0] = f.snu; s_sc[1] = f.inv; s_sc[2] = (ci + f.snu) * f.inv;
        if (g == 0 && crank == 0) a.alpha[i] = -f.snu;
    }
    __syncthreads();
    const double snu = s_sc[0], inv = s_sc[1], vi = s_sc[2];
    if (tid < k) {
        const double pv = a.P[i + (long)tid * a.ldp], px = a.P[i + (long)(nb + tid) * a.ldp];
        s_rowV[tid] = pv;
        s_rowX[tid] = px;
        s_vTv[tid] = (a.dots1[tid] + snu * pv) * inv;
        s_xTv[tid] = (a.dots1[nb + tid] + snu * px) * inv;
    }
    __syncthreads();
    fz_cluster_sync();                                     // barriers exist everywhere before remote arrives

    const int ntiles = (a.T > g) ? (a.T - g + NC - 1) / NC : 0;

    if (warp == 0) {
        // ============================ TMA producer warp ============================
        for (int nt = 0; nt < ntiles; ++nt) {
            const int s = nt % FZ_STAGES;
            fz_mbar_wait(empty + s, ((nt / FZ_STAGES) & 1) ^ 1);
            if (lane == 0) {
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
    } else if (warp <= FZ_HW) {
        // ============================== helper warps ==============================
        // per tile column j: a_ij, corr_j = Y[j,:].vTv + U[j,:].xTv, g_j = rowV.Y[j,:] + rowX.U[j,:]
        // and a copy of the panel rows for the dot partials; all loads are issued before any use
        for (int nt = warp - 1; nt < ntiles; nt += FZ_HW) {
            const int s = nt % FZ_STAGES;
            const int j0 = i + 1 + (g + nt * NC) * CBW;
            int ncols = a.n - j0;
            if (ncols > CBW) ncols = CBW;
            double yk[CBW][2], uk[CBW][2], aij[CBW];
#pragma unroll
            for (int q = 0; q < CBW; ++q) {
                const int j = (q < ncols) ? j0 + q : j0;
                aij[q] = (lane == 0) ? a.A[i + (long)j * a.lda] : 0.0;
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    const int kk = lane + 32 * z;
                    yk[q][z] = (kk < k) ? a.Q[j + (long)kk * a.ldq] : 0.0;
                    uk[q][z] = (kk < k) ? a.Q[j + (long)(nb + kk) * a.ldq] : 0.0;
                }
            }
            fz_mbar_wait(empty + s, ((nt / FZ_STAGES) & 1) ^ 1);     // stage (and its helper slots) free
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
                if (lane == 0) { hcorr[s * 8 + q] = corr; hg[s * 8 + q] = gg; haij[s * 8 + q] = aij[q]; }
            }
            if (lane == 0) hn[s] = ncols;
            __syncwarp();
            if (lane == 0) fz_mbar_arrive(full + s);
        }
    } else {
        // ============================ consumer warps ============================
        const int cw = warp - 1 - FZ_HW, ct = tid - FZ_PRE;
        double2 creg[RPT], acc[RPT];
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            const int lr = 2 * ct + 1024 * u;
            creg[u] = (lr < len) ? *reinterpret_cast<const double2 *>(a.c + rs + lr) : make_double2(0.0, 0.0);
            acc[u] = make_double2(0.0, 0.0);
        }
        if (g == 0) {
            // the reflector itself: v = (c + s*nu*e_i) * inv, in place and into the V panel
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
                const int lr = 2 * ct + 1024 * u;
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
        double dY[2] = {0.0, 0.0}, dU[2] = {0.0, 0.0}, rr2 = 0.0, yr = 0.0;
        const bool colwarp = (cw == 0);

        for (int nt = 0; nt < ntiles + FZ_D; ++nt) {
            if (nt < ntiles) {
                // ---- part A of tile nt: sweep 1 (column dots with c) and the cluster exchange
                const int s = nt % FZ_STAGES;
                fz_mbar_wait(full + s, (nt / FZ_STAGES) & 1);
                const int ncols = hn[s];
                const double *tl = tile + (size_t)s * FZ_STAGE;
                double psum[CBW];
#pragma unroll
                for (int q = 0; q < CBW; ++q) {
                    psum[q] = 0.0;
                    if (q < ncols) {
#pragma unroll
                        for (int u = 0; u < RPT; ++u) {
                            const int lr = 2 * ct + 1024 * u;
                            if (lr < len) {
                                const double2 av = *reinterpret_cast<const double2 *>(tl + (size_t)q * Lc + lr);
                                psum[q] += av.x * creg[u].x + av.y * creg[u].y;
                            }
                        }
                    }
                    psum[q] = warp_sum(psum[q]);
                }
                if (lane == 0) {
#pragma unroll
                    for (int q = 0; q < CBW; ++q) wsum[((nt & 1) * FZ_CW + cw) * 8 + q] = psum[q];
                }
                fz_consumer_bar();
                if (colwarp) {
                    const int xs = nt % FZ_XR;
                    // fixed-shape shuffle tree over the 16 per-warp partials of every column
                    double vals[CBW];
#pragma unroll
                    for (int qq = 0; qq < CBW; ++qq) {
                        double v = (lane < FZ_CW) ? wsum[((nt & 1) * FZ_CW + lane) * 8 + qq] : 0.0;
                        vals[qq] = warp_sum(v);
                    }
                    if (CS == 1) {
                        if (lane == 0) {
#pragma unroll
                            for (int qq = 0; qq < CBW; ++qq) xsum[(xs * FZ_MAXCS) * 8 + qq] = vals[qq];
                        }
                        __syncwarp();
                        if (lane == 0) fz_mbar_arrive(xbar + xs);
                    } else {
                        // every CTA of the cluster receives CBW sums from each of the CS CTAs
                        if (lane == 0) fz_mbar_arrive_expect_tx(xbar + xs, CS * CBW * 8u);
                        if (lane < (int)CS) {
                            const unsigned base = fz_mapa(fz_smem_u32(xsum + (xs * FZ_MAXCS + (int)crank) * 8), (unsigned)lane);
                            const unsigned rbar = fz_mapa(fz_smem_u32(xbar + xs), (unsigned)lane);
#pragma unroll
                            for (int qq = 0; qq < CBW; ++qq) fz_st_async(base + 8u * qq, vals[qq], rbar);
                        }
                    }
                }
            }
            if (nt >= FZ_D) {
                // ---- part B of tile nt-FZ_D: y, r per column, sweep 2 (row dots with r)
                const int pt = nt - FZ_D;
                const int s = pt % FZ_STAGES, xs = pt % FZ_XR;
                fz_mbar_wait(xbar + xs, (pt / FZ_XR) & 1);
                const int ncols = hn[s];
                const double *tl = tile + (size_t)s * FZ_STAGE;
                const int j0 = i + 1 + (g + pt * NC) * CBW;
#pragma unroll
                for (int q = 0; q < CBW; ++q) {
                    if (q < ncols) {
                        double tsum = 0.0;
                        for (unsigned rk = 0; rk < CS; ++rk) tsum += xsum[(xs * FZ_MAXCS + (int)rk) * 8 + q];
                        const double aij = haij[s * 8 + q];
                        const double y = 2.0 * ((tsum + snu * aij) * inv - hcorr[s * 8 + q]);
                        const double r = aij - hg[s * 8 + q] - vi * y;
#pragma unroll
                        for (int u = 0; u < RPT; ++u) {
                            const int lr = 2 * ct + 1024 * u;
                            if (lr < len) {
                                const double2 av = *reinterpret_cast<const double2 *>(tl + (size_t)q * Lc + lr);
                                acc[u].x += av.x * r;
                                acc[u].y += av.y * r;
                            }
                        }
                        if (colwarp && crank == 0) {
                            if (lane == 0) {
                                a.Q[(j0 + q) + (long)k * a.ldq] = y;
                                a.rv[j0 + q] = r;
                            }
                            const double *qr = qrow + (size_t)(s * FZ_CBW_MAX + q) * 2 * NBMAX;
#pragma unroll
                            for (int z = 0; z < 2; ++z) {
                                const int kk = lane + 32 * z;
                                if (kk < k) { dY[z] += qr[kk] * r; dU[z] += qr[NBMAX + kk] * r; }
                            }
                            rr2 += r * r;
                            yr += y * r;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) fz_mbar_arrive(empty + s);
            }
        }

        // ---- epilogue: partial t2 of this cluster, partial panel dots, last-cluster combine
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            const int lr = 2 * ct + 1024 * u;
            if (lr < len) *reinterpret_cast<double2 *>(a.tmpN + (long)g * a.ldt + rs + lr) = acc[u];
        }
        const int S2 = 2 * nb + 2;
        if (colwarp && crank == 0) {
            double *out = a.dots2p + (long)g * S2;
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int kk = lane + 32 * z;
                if (kk < k) { out[kk] = dY[z]; out[nb + kk] = dU[z]; }
            }
            if (lane == 0) { out[k] = yr; out[2 * nb] = rr2; }
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                const unsigned old = atomicAdd(a.counter, 1u);
                hn[FZ_STAGES] = (old == (unsigned)(NC - 1)) ? 1 : 0;
            }
        }
        fz_consumer_bar();
        if (hn[FZ_STAGES]) {
            __threadfence();
            // entries: [0..k] Y^T r, [nb..nb+k) U^T r, [2nb] r.r  -> 2k+2 values, 2 threads each
            const int e = ct >> 1, half = ct & 1, ne = 2 * k + 2;
            const int slot = (e <= k) ? e : (e <= 2 * k ? nb + (e - k - 1) : 2 * nb);
            double sacc = 0.0;
            if (e < ne) {
                const int h0 = half ? (NC + 1) / 2 : 0, h1 = half ? NC : (NC + 1) / 2;
                for (int p2 = h0; p2 < h1; ++p2) sacc += __ldcg(a.dots2p + (long)p2 * S2 + slot);
            }
            const double other = __shfl_xor_sync(0xffffffffu, sacc, 1);
            if (e < ne && half == 0) a.dots2[slot] = sacc + other;
            if (ct == 0) *a.counter = 0u;
        }
    }
    fz_cluster_sync();      // nobody exits while a peer may still write into its shared memory
}

} // namespace svdgpu
