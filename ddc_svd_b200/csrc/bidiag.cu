// bidiag.cu — Golub-Kahan Householder bidiagonalization on sm_100a.
//
// Replaces the device half of the reference's bidiag_par() (bidiag_par.c:34-450) and its ten
// OpenCL kernels (normsq_matcol/matrow, sum, update_scale_matcol/matrow, sanders,
// left/right_dotprods, left/right_update_mat).  Same mathematics and the same stored
// result — unit-2-norm reflectors v (H = I - 2 v v^T) left in place in A, alpha/beta with
// the reference's sign rule (update_scale_matcol.cl:58-82) — but a different organisation:
//
//   * panel-deferred updates (the dlabrd idea): inside a panel of nb steps the trailing
//     matrix is NOT rewritten; A_cur = A - V Y^T - X U^T is carried by four thin panels and
//     the trailing block is updated once per panel by a rank-2nb FP64 DMMA GEMM.  Per step
//     the trailing matrix is therefore only READ twice (the reference reads/writes it 6x).
//   * per step exactly two streaming passes over the trailing matrix:
//       gemvT:  t = A^T c   (column dots; warp per 4 columns, 128-bit loads down the column)
//       gemvN:  t = A  r    (row dots; thread per 2 rows, 128-bit loads, columns unrolled x8)
//     with the *unnormalised* current column c / row r, so that the norm, the sign and the
//     scale of the reflector (normsq_* + sum + update_scale_* in the reference) are applied
//     algebraically afterwards — no extra pass and no extra grid-wide dependency.
//   * the panel dot products (V^T c, X^T c, Y^T r, U^T r) and the norms c.c / r.r ride along
//     in extra CTAs of the same two launches.
//   * two small elementwise kernels (finish_y, finish_x) turn the pass results into the new
//     panel columns, write the reflectors in place and form the next column / row.
// 4 launches per step (reference: 10-12), no host synchronisation inside the loop.
#include "common.cuh"
#include "bidiag.cuh"
#include <vector>

namespace svdgpu {

constexpr int NBMAX = 64;
constexpr size_t TAIL_WS_SLOTS = (size_t)148 * 2048 + 3 * 2048 + 2 * 148 + 8;   // == TL_WS_SLOTS (bidiag_tail.cuh)
constexpr int TAIL_DEFAULT_ON = 1;          // on-chip tail kernel (bidiag_tail.cuh)
constexpr int NB_BIG_DEFAULT = 0;          // > 0: panel width while the trailing block is large (see bidiag_device)
constexpr int NB_BIG_MIN_DEFAULT = 6144;   // trailing rows/columns from which NB_BIG is used
constexpr int GT_CW = 4;          // columns per warp in gemvT
constexpr int GT_WARPS = 8;
constexpr int GN_THREADS = 256;   // each thread owns 2 rows in gemvN
constexpr int GN_UNROLL = 8;
constexpr int GN_CHUNK = 4096;    // columns of r staged in shared memory at a time (32 KB)

// ---------------------------------------------------------------------------------------
// block-wide dot of two strided-1 vectors (used by the "extra" CTAs)
__device__ double block_dot(const double *__restrict__ x, const double *__restrict__ y, int len)
{
    double acc = 0.0;
    int t = threadIdx.x;
    int l4 = len & ~3;
    for (int p = t * 4; p < l4; p += blockDim.x * 4) {
        acc += x[p] * y[p] + x[p + 1] * y[p + 1] + x[p + 2] * y[p + 2] + x[p + 3] * y[p + 3];
    }
    for (int p = l4 + t; p < len; p += blockDim.x) acc += x[p] * y[p];
    acc = warp_sum(acc);
    __shared__ double red[32];
    __syncthreads();
    if ((t & 31) == 0) red[t >> 5] = acc;
    __syncthreads();
    double tot = 0.0;
    if (t < 32) {
        tot = (t < (int)(blockDim.x >> 5)) ? red[t] : 0.0;
        tot = warp_sum(tot);
    }
    return tot;   // valid in warp 0
}

// ---------------------------------------------------------------------------------------
// gemvT: tmp[s][j] = sum_{r in split s} A[r,j] * c[r]   for j in (i, n)
//   grid.x = colGroups + nextra, grid.y = nsplit.  Extra CTAs (y == 0 only): panel dots
//   dots[w] = P[:,w] . c  (w < k), dots[nb+w] = P[:,nb+w] . c (w < k), dots[2nb] = c . c,
//   all over rows [i, m).
__global__ void __launch_bounds__(GT_WARPS * 32)
gemvT_kernel(const double *__restrict__ A, long lda, int i, int m, int n, int mpad,
             const double *__restrict__ c, double *__restrict__ tmp, long ldt,
             int colGroups, int rowsPerSplit,
             const double *__restrict__ P, long ldp, int nb, int k, double *__restrict__ dots)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if ((int)blockIdx.x >= colGroups) {
        if (blockIdx.y != 0) return;
        int w = blockIdx.x - colGroups;                       // 0 .. 2k
        const double *x = (w < k) ? P + (long)w * ldp : (w < 2 * k ? P + (long)(nb + w - k) * ldp : c);
        int slot = (w < k) ? w : (w < 2 * k ? nb + (w - k) : 2 * nb);
        double d = block_dot(x + i, c + i, m - i);
        if (threadIdx.x == 0) dots[slot] = d;
        return;
    }
    const int j0 = i + 1 + (blockIdx.x * GT_WARPS + warp) * GT_CW;
    if (j0 >= n) return;
    const int rbeg = (i & ~1) + blockIdx.y * rowsPerSplit;
    const int rend = min(mpad, rbeg + rowsPerSplit);
    const double *col[GT_CW];
#pragma unroll
    for (int q = 0; q < GT_CW; ++q) col[q] = A + (long)min(j0 + q, n - 1) * lda;
    double acc[GT_CW];
#pragma unroll
    for (int q = 0; q < GT_CW; ++q) acc[q] = 0.0;

    int r = rbeg + 2 * lane;
    for (; r + 192 < rend; r += 256) {
        double2 cv[4], av[4][GT_CW];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            cv[u] = *reinterpret_cast<const double2 *>(c + r + 64 * u);
#pragma unroll
            for (int q = 0; q < GT_CW; ++q) av[u][q] = ldg_stream2(col[q] + r + 64 * u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int q = 0; q < GT_CW; ++q) acc[q] += av[u][q].x * cv[u].x + av[u][q].y * cv[u].y;
    }
    for (; r < rend; r += 64) {
        double2 cv = *reinterpret_cast<const double2 *>(c + r);
#pragma unroll
        for (int q = 0; q < GT_CW; ++q) {
            double2 av = ldg_stream2(col[q] + r);
            acc[q] += av.x * cv.x + av.y * cv.y;
        }
    }
#pragma unroll
    for (int q = 0; q < GT_CW; ++q) {
        double s = warp_sum(acc[q]);
        if (lane == 0 && j0 + q < n) tmp[(long)blockIdx.y * ldt + j0 + q] = s;
    }
}

// ---------------------------------------------------------------------------------------
// gemvN: tmp[s][r] = sum_{j in split s} A[r,j] * rv[j]   for rows r >= rbase=(i+1)&~1
//   grid.x = rowBlocks + nextra, grid.y = nsplit.  Extra CTAs: dots[w] = Q[:,w] . rv (w <= k),
//   dots[nb+w] = Q[:,nb+w] . rv (w < k), dots[2nb] = rv . rv, over j in (i, n).
__global__ void __launch_bounds__(GN_THREADS)
gemvN_kernel(const double *__restrict__ A, long lda, int i, int m, int n, int mpad,
             const double *__restrict__ rv, double *__restrict__ tmp, long ldt,
             int rowBlocks, int colsPerSplit,
             const double *__restrict__ Q, long ldq, int nb, int k, double *__restrict__ dots)
{
    if ((int)blockIdx.x >= rowBlocks) {
        if (blockIdx.y != 0) return;
        int w = blockIdx.x - rowBlocks;                       // 0 .. 2k+1
        const double *x = (w <= k) ? Q + (long)w * ldq
                                   : (w <= 2 * k ? Q + (long)(nb + w - k - 1) * ldq : rv);
        int slot = (w <= k) ? w : (w <= 2 * k ? nb + (w - k - 1) : 2 * nb);
        double d = block_dot(x + i + 1, rv + i + 1, n - i - 1);
        if (threadIdx.x == 0) dots[slot] = d;
        return;
    }
    __shared__ double s_rv[GN_CHUNK];                         // r is staged GN_CHUNK columns at a time
    const int jbeg = i + 1 + blockIdx.y * colsPerSplit;
    const int jend = min(n, jbeg + colsPerSplit);
    if (jbeg >= jend) return;
    const int r = ((i + 1) & ~1) + (blockIdx.x * GN_THREADS + threadIdx.x) * 2;
    const bool live = r < mpad;
    double2 acc = make_double2(0.0, 0.0);
    for (int jc = jbeg; jc < jend; jc += GN_CHUNK) {
        const int nc = min(GN_CHUNK, jend - jc);
        if (jc > jbeg) __syncthreads();
        for (int t = threadIdx.x; t < nc; t += GN_THREADS) s_rv[t] = rv[jc + t];
        __syncthreads();
        if (!live) continue;
        const double *a = A + r + (long)jc * lda;
        int j = 0;
        for (; j + GN_UNROLL <= nc; j += GN_UNROLL) {
            double2 av[GN_UNROLL];
#pragma unroll
            for (int u = 0; u < GN_UNROLL; ++u) av[u] = ldg_stream2(a + (long)(j + u) * lda);
#pragma unroll
            for (int u = 0; u < GN_UNROLL; ++u) {
                double w = s_rv[j + u];
                acc.x += av[u].x * w;
                acc.y += av[u].y * w;
            }
        }
        for (; j < nc; ++j) {
            double2 av = ldg_stream2(a + (long)j * lda);
            double w = s_rv[j];
            acc.x += av.x * w;
            acc.y += av.y * w;
        }
    }
    if (!live) return;
    *reinterpret_cast<double2 *>(tmp + (long)blockIdx.y * ldt + r) = acc;
}

// ---------------------------------------------------------------------------------------
// The reflector recipe of the reference (update_scale_matcol.cl:45-82, bidiag.c:76-94)
// expressed on the unnormalised vector: x0 first entry, nrm2 = x.x
//   out = -s*nu ; v = (x + s*nu*e0) * inv ; inv = 1/(sqrt2*sqrt(nu^2+|nu*x0|))
struct Refl { double snu, inv; };
__device__ __forceinline__ Refl make_refl(double x0, double nrm2)
{
    Refl f;
    double nu = sqrt(nrm2);
    double s = (x0 < 0.0) ? -1.0 : 1.0;
    f.snu = s * nu;
    double sc = sqrt(2.0) * sqrt(nu * nu + fabs(nu * x0));
    f.inv = (sc > 0.0) ? 1.0 / sc : 0.0;      // zero vector: H = I (the reference would NaN here)
    return f;
}

// The two "finish" kernels are elementwise in the row / column index but carry a 2k-term panel
// correction per element.  To keep them off the critical path they are parallelised over the
// panel index as well: a CTA is 32 elements (lanes) x FK_SLICES panel slices (warps); slice w
// handles panel columns q = w, w+8, ... and split partials s = w, w+8, ...; the slices are
// combined through shared memory in a fixed order (deterministic).
constexpr int FK_SLICES = 8;
constexpr int DOT_SLOTS_C = 2 * NBMAX + 2;

// finish_y: after gemvT of step i (panel column k).
//   blocks [0, nColBlk): 32 trailing columns each -> y_j (new Y column), r_j (row i after H)
//   blocks [nColBlk, ..): 256 rows each -> v written in place (A[:,i]) and into the V panel
__global__ void __launch_bounds__(256)
finish_y_kernel(double *__restrict__ A, long lda, int i, int m, int n, int k, int nb, int do_col,
                double *__restrict__ P, long ldp, double *__restrict__ Q, long ldq,
                const double *__restrict__ c, double *__restrict__ rv,
                const double *__restrict__ tmp, long ldt, int nsplit,
                const double *__restrict__ dots, double *__restrict__ alpha, int nColBlk)
{
    __shared__ double s_vTv[NBMAX], s_xTv[NBMAX], s_rowV[NBMAX], s_rowX[NBMAX];
    __shared__ double s_red[3][FK_SLICES][32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const double ci = c[i];
    Refl f;
    if (do_col) f = make_refl(ci, dots[2 * nb]); else { f.snu = 0.0; f.inv = 0.0; }
    const int R = n - i - 1, L = m - i;
    if ((int)blockIdx.x >= nColBlk) {
        // ---- row part: the reflector itself
        const int idx = (blockIdx.x - nColBlk) * 256 + t;
        if (idx == 0) alpha[i] = do_col ? -f.snu : ci;
        if (idx < L) {
            const int r = i + idx;
            if (do_col) {
                double v = (c[r] + (idx == 0 ? f.snu : 0.0)) * f.inv;
                A[r + (long)i * lda] = v;
                P[r + (long)k * ldp] = v;
            } else {
                if (idx == 0) A[r + (long)i * lda] = 0.0;     // bidiag.c:160-162 "no reflection on left"
                P[r + (long)k * ldp] = 0.0;
            }
        }
        return;
    }
    if (t < k) {
        double pv = P[i + (long)t * ldp], px = P[i + (long)(nb + t) * ldp];
        s_rowV[t] = pv;
        s_rowX[t] = px;
        s_vTv[t] = (dots[t] + f.snu * pv) * f.inv;
        s_xTv[t] = (dots[nb + t] + f.snu * px) * f.inv;
    }
    __syncthreads();
    const int idx = blockIdx.x * 32 + lane;
    const int j = i + 1 + idx;
    double corr = 0.0, sub = 0.0, tt = 0.0;
    if (idx < R) {
        if (do_col) for (int sp = w; sp < nsplit; sp += FK_SLICES) tt += tmp[(long)sp * ldt + j];
#pragma unroll 4
        for (int q = w; q < k; q += FK_SLICES) {
            double yk = Q[j + (long)q * ldq], uk = Q[j + (long)(nb + q) * ldq];
            corr += yk * s_vTv[q] + uk * s_xTv[q];
            sub += s_rowV[q] * yk + s_rowX[q] * uk;
        }
    }
    s_red[0][w][lane] = corr; s_red[1][w][lane] = sub; s_red[2][w][lane] = tt;
    __syncthreads();
    if (w == 0 && idx < R) {
        corr = 0.0; sub = 0.0; tt = 0.0;
#pragma unroll
        for (int z = 0; z < FK_SLICES; ++z) { corr += s_red[0][z][lane]; sub += s_red[1][z][lane]; tt += s_red[2][z][lane]; }
        const double aij = A[i + (long)j * lda];
        const double vi = (ci + f.snu) * f.inv;
        double y = do_col ? 2.0 * ((tt + f.snu * aij) * f.inv - corr) : 0.0;
        Q[j + (long)k * ldq] = y;
        rv[j] = aij - sub - vi * y;
    }
}

// finish_x: after the row-dot pass of step i.  Also forms the next current column c' (column i+1).
//   blocks [0, nRowBlk): 32*RG rows each -> x_r (new X column), c'_r
//   blocks [nRowBlk, ..): 256*RG columns each -> u written in place (A[i,:]) and into the U panel
// With dots1p != NULL (fused path) every row block also leaves its partial [V^T c' | X^T c' | c'.c']
// and the last row block to finish combines all partials, in a fixed order, into dots1.
template <int RG>
__global__ void __launch_bounds__(256 * RG)
finish_x_kernel(double *__restrict__ A, long lda, int i, int m, int n, int k, int nb, int do_row,
                double *__restrict__ P, long ldp, double *__restrict__ Q, long ldq,
                double *__restrict__ c, const double *__restrict__ rv,
                const double *__restrict__ tmp, long ldt, int nsplit,
                const double *__restrict__ dots, double *__restrict__ beta, int nRowBlk,
                double *__restrict__ dots1p, double *__restrict__ dots1, unsigned *__restrict__ counter)
{
    __shared__ double s_yTu[NBMAX], s_uTu[NBMAX], s_rowY[NBMAX], s_rowU[NBMAX];
    __shared__ double s_red[RG][3][FK_SLICES][32];
    __shared__ double s_c[RG][32], s_x[RG][32];
    __shared__ double s_dp[RG][2 * NBMAX + 2];
    __shared__ int s_last;
    const int t = threadIdx.x, lane = t & 31, grp = t >> 8, w = (t >> 5) & (FK_SLICES - 1);
    const int R = n - i - 1, Lb = m - i - 1;
    // ---- issue the long-latency loads of the row part first (they do not depend on the scalars)
    constexpr int ZV = NBMAX / FK_SLICES + 1, ZX = NBMAX / FK_SLICES;
    const int idx = (blockIdx.x * RG + grp) * 32 + lane;
    const int r = i + 1 + idx;
    const bool live = ((int)blockIdx.x < nRowBlk && R > 0 && idx < Lb);
    double vk[ZV], xk[ZX], tt = 0.0, ar = 0.0;
#pragma unroll
    for (int z = 0; z < ZV; ++z) { const int q = w + FK_SLICES * z; vk[z] = (live && q <= k) ? P[r + (long)q * ldp] : 0.0; }
#pragma unroll
    for (int z = 0; z < ZX; ++z) { const int q = w + FK_SLICES * z; xk[z] = (live && q < k) ? P[r + (long)(nb + q) * ldp] : 0.0; }
    if (live) {
        if (do_row) {
#pragma unroll 4
            for (int sp = w; sp < nsplit; sp += FK_SLICES) tt += tmp[(long)sp * ldt + r];
        }
        if (w == 0) ar = A[r + (long)(i + 1) * lda];
    }
    const double rf = (R > 0) ? rv[i + 1] : 0.0;
    Refl f;
    if (do_row) f = make_refl(rf, dots[2 * nb]); else { f.snu = 0.0; f.inv = 0.0; }
    const double ufirst = do_row ? (rf + f.snu) * f.inv : 0.0;
    if ((int)blockIdx.x >= nRowBlk) {
        // ---- column part: the row reflector itself
        const int idx = (blockIdx.x - nRowBlk) * (256 * RG) + t;
        if (idx == 0 && R > 0) {
            beta[i] = do_row ? -f.snu : rf;
            if (!do_row) A[i + (long)(i + 1) * lda] = 0.0;    // bidiag.c:124-128
        }
        if (idx < R) {
            const int j = i + 1 + idx;
            double u = do_row ? (rv[j] + (idx == 0 ? f.snu : 0.0)) * f.inv : 0.0;
            if (do_row) A[i + (long)j * lda] = u;
            Q[j + (long)(nb + k) * ldq] = u;
        }
        return;
    }
    if (R > 0) {
        if (t <= k) {
            double qy = Q[(i + 1) + (long)t * ldq];
            s_rowY[t] = qy;
            s_yTu[t] = (dots[t] + f.snu * qy) * f.inv;
        }
        if (t < k) {
            double qu = Q[(i + 1) + (long)(nb + t) * ldq];
            s_rowU[t] = qu;
            s_uTu[t] = (dots[nb + t] + f.snu * qu) * f.inv;
        }
    }
    __syncthreads();
    double corr = 0.0, sub = 0.0;
#pragma unroll
    for (int z = 0; z < ZV; ++z) { const int q = w + FK_SLICES * z; if (q <= k) { corr += vk[z] * s_yTu[q]; sub += vk[z] * s_rowY[q]; } }
#pragma unroll
    for (int z = 0; z < ZX; ++z) { const int q = w + FK_SLICES * z; if (q < k) { corr += xk[z] * s_uTu[q]; sub += xk[z] * s_rowU[q]; } }
    s_red[grp][0][w][lane] = corr; s_red[grp][1][w][lane] = sub; s_red[grp][2][w][lane] = tt;
    __syncthreads();
    if (w == 0) {
        double cc = 0.0, x = 0.0;
        if (live) {
            corr = 0.0; sub = 0.0; tt = 0.0;
#pragma unroll
            for (int z = 0; z < FK_SLICES; ++z) {
                corr += s_red[grp][0][z][lane]; sub += s_red[grp][1][z][lane]; tt += s_red[grp][2][z][lane];
            }
            x = do_row ? 2.0 * ((tt + f.snu * ar) * f.inv - corr) : 0.0;
            P[r + (long)(nb + k) * ldp] = x;
            cc = ar - sub - x * ufirst;
            c[r] = cc;
        }
        s_c[grp][lane] = cc;
        s_x[grp][lane] = x;
        if (blockIdx.x == 0 && grp == 0 && lane == 0) c[i] = 0.0;   // keeps the 128-bit loads of the next pass harmless
    }
    if (dots1p == nullptr) return;
    // ---- partial panel dots of the NEW column c' for the fused pass of step i+1:
    //      [0..k]: V^T c' (k+1 columns incl. the v just stored), [nb..nb+k]: X^T c', [2nb]: c'.c'
    __syncthreads();
    const int S1 = 2 * nb + 2;
    const double cl = s_c[grp][lane];
#pragma unroll
    for (int z = 0; z < ZV; ++z) {
        const int q = w + FK_SLICES * z;
        if (q <= k) {                                       // warp-uniform
            const double xv = (q < k) ? xk[z < ZX ? z : 0] : s_x[grp][lane];
            const double dv = warp_sum(vk[z] * cl), dx = warp_sum(xv * cl);
            if (lane == 0) { s_dp[grp][q] = dv; s_dp[grp][nb + q] = dx; }
        }
    }
    if (w == 0) {
        double cc2 = warp_sum(cl * cl);
        if (lane == 0) s_dp[grp][2 * nb] = cc2;
    }
    __syncthreads();
    const int ne = 2 * k + 3;                       // entries: [0..k], [nb..nb+k], [2nb]
    {
        double *out = dots1p + (long)blockIdx.x * S1;
        for (int e = t; e < ne; e += 256 * RG) {
            const int slot = (e <= k) ? e : (e <= 2 * k + 1 ? nb + (e - k - 1) : 2 * nb);
            double a2 = 0.0;
#pragma unroll
            for (int gq = 0; gq < RG; ++gq) a2 += s_dp[gq][slot];
            out[slot] = a2;
        }
    }
    __threadfence();
    __syncthreads();
    if (t == 0) s_last = (atomicAdd(counter, 1u) == (unsigned)(nRowBlk - 1)) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        // TPE threads per entry, partials split in TPE contiguous ranges, combined in a fixed order
        constexpr int TPE = (RG >= 4) ? 4 : 1;
        const int part = t % TPE;
        for (int e0 = 0; e0 < ne; e0 += (256 * RG) / TPE) {
            const int e = e0 + t / TPE;
            const int slot = (e <= k) ? e : (e <= 2 * k + 1 ? nb + (e - k - 1) : 2 * nb);
            double sacc = 0.0;
            if (e < ne) {
                const int chunk = (nRowBlk + TPE - 1) / TPE;
                const int p0 = part * chunk, p1 = min(nRowBlk, p0 + chunk);
#pragma unroll 8
                for (int pz = p0; pz < p1; ++pz) sacc += __ldcg(dots1p + (long)pz * S1 + slot);
            }
            if (TPE == 4) {
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            }
            if (e < ne && part == 0) dots1[slot] = sacc;
        }
        if (t == 0) *counter = 0u;
    }
}

// finish_xf: finish_x for the fused path.  1024 threads = 32 rows (lanes) x 32 panel slices (warps);
// a CTA walks RB row groups.  Both partial-dot reductions live in the consumers' prologues: this
// kernel first combines the fused pass's per-cluster partials [Y^T r | U^T r | r.r] (fixed order),
// and leaves per-CTA partials [V^T c' | X^T c' | c'.c'] for the next fused pass to combine — no
// atomics, no fences, no serial "last block".
constexpr int XF_SL = 32;
template <int RB>
__global__ void __launch_bounds__(1024)
finish_xf_kernel(double *__restrict__ A, long lda, int i, int m, int n, int k, int nb,
                 double *__restrict__ P, long ldp, double *__restrict__ Q, long ldq,
                 double *__restrict__ c, const double *__restrict__ rv,
                 const double *__restrict__ tmp, long ldt, int nsplit,
                 const double *__restrict__ dots2p, int nparts2, double *__restrict__ beta, int nRowBlk,
                 double *__restrict__ dots1p)
{
    constexpr int S = DOT_SLOTS_C;
    __shared__ double s_d[S];
    __shared__ double s_yTu[NBMAX], s_uTu[NBMAX], s_rowY[NBMAX], s_rowU[NBMAX];
    __shared__ double s_red[3][XF_SL][33];
    __shared__ double s_red2[3][8][32];
    __shared__ double s_c[32], s_x[32];
    // the prologue's partial sums live in s_red's storage (used only after they are consumed)
    static_assert(15 * DOT_SLOTS_C <= 3 * XF_SL * 33, "s_part does not fit into s_red");
    double(*s_part)[S] = reinterpret_cast<double(*)[S]>(&s_red[0][0][0]);
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int R = n - i - 1, Lb = m - i - 1;
    const bool rowblk = (int)blockIdx.x < nRowBlk;
    constexpr int ZV = NBMAX / XF_SL + 1, ZX = NBMAX / XF_SL;       // 3, 2
    // programmatic dependent launch: started while the fused pass drains; its results are read below
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // ---- loads that depend on nothing: first row group, the row of Q, r_first
    int idx = blockIdx.x * (32 * RB) + lane;
    bool live = rowblk && idx < Lb;
    double vk[ZV], xk[ZX], tt = 0.0, ar = 0.0;
#pragma unroll
    for (int z = 0; z < ZV; ++z) { const int q = w + XF_SL * z; vk[z] = (live && q <= k) ? P[(i + 1 + idx) + (long)q * ldp] : 0.0; }
#pragma unroll
    for (int z = 0; z < ZX; ++z) { const int q = w + XF_SL * z; xk[z] = (live && q < k) ? P[(i + 1 + idx) + (long)(nb + q) * ldp] : 0.0; }
    if (live) {
        // all of this slice's partials in flight at once (<= 5 for up to 160 clusters)
        for (int sp0 = w; sp0 < nsplit; sp0 += 5 * XF_SL) {
            double tv[5];
#pragma unroll
            for (int u = 0; u < 5; ++u) {
                const int sp = sp0 + u * XF_SL;
                tv[u] = (sp < nsplit) ? tmp[(long)sp * ldt + i + 1 + idx] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 5; ++u) tt += tv[u];
        }
        if (w == 0) ar = A[(i + 1 + idx) + (long)(i + 1) * lda];
    }
    const double rf = rv[i + 1];
    double qy = 0.0, qu = 0.0;
    if (t <= k) qy = Q[(i + 1) + (long)t * ldq];
    if (t < k) qu = Q[(i + 1) + (long)(nb + t) * ldq];
    // ---- combine the pass's partial dots: entries [0..k], [nb..nb+k), [2nb]; TPE threads per entry,
    //      each with all the loads of its chunk in flight at once
    const int ne = 2 * k + 2;
    const int TPE = (ne <= 68) ? 15 : 7;                    // 1024 threads: 68 x 15 or 130 x 7
    {
        const int e = t / TPE, part = t - TPE * e;
        if (e < ne) {
            const int slot = (e <= k) ? e : (e <= 2 * k ? nb + (e - k - 1) : 2 * nb);
            const int chunk = (nparts2 + TPE - 1) / TPE, p0 = part * chunk, p1 = min(nparts2, p0 + chunk);
            constexpr int LB = 10;
            double a2 = 0.0;
            for (int base = p0; base < p1; base += LB) {
                double v[LB];
#pragma unroll
                for (int u = 0; u < LB; ++u) v[u] = (base + u < p1) ? dots2p[(long)(base + u) * S + slot] : 0.0;
#pragma unroll
                for (int u = 0; u < LB; ++u) a2 += v[u];
            }
            s_part[part][slot] = a2;
        }
    }
    __syncthreads();
    if (t < S && (t <= k || (t >= nb && t < nb + k) || t == 2 * nb)) {
        double a2 = 0.0;
        for (int pz = 0; pz < TPE; ++pz) a2 += s_part[pz][t];
        s_d[t] = a2;
    }
    __syncthreads();
    const Refl f = make_refl(rf, s_d[2 * nb]);
    const double ufirst = (rf + f.snu) * f.inv;
    if (!rowblk) {
        // ---- column part: the row reflector itself
        const int cidx = (blockIdx.x - nRowBlk) * 1024 + t;
        if (cidx == 0) beta[i] = -f.snu;
        if (cidx < R) {
            const int j = i + 1 + cidx;
            const double u = (rv[j] + (cidx == 0 ? f.snu : 0.0)) * f.inv;
            A[i + (long)j * lda] = u;
            Q[j + (long)(nb + k) * ldq] = u;
        }
        return;
    }
    if (t <= k) { s_rowY[t] = qy; s_yTu[t] = (s_d[t] + f.snu * qy) * f.inv; }
    if (t < k) { s_rowU[t] = qu; s_uTu[t] = (s_d[nb + t] + f.snu * qu) * f.inv; }
    __syncthreads();

    double dv[ZV], dx[ZV], cc2 = 0.0;
#pragma unroll
    for (int z = 0; z < ZV; ++z) { dv[z] = 0.0; dx[z] = 0.0; }
#pragma unroll 1
    for (int rb = 0; rb < RB; ++rb) {
        if (rb > 0) {
            idx = blockIdx.x * (32 * RB) + rb * 32 + lane;
            live = idx < Lb;
            tt = 0.0; ar = 0.0;
#pragma unroll
            for (int z = 0; z < ZV; ++z) { const int q = w + XF_SL * z; vk[z] = (live && q <= k) ? P[(i + 1 + idx) + (long)q * ldp] : 0.0; }
#pragma unroll
            for (int z = 0; z < ZX; ++z) { const int q = w + XF_SL * z; xk[z] = (live && q < k) ? P[(i + 1 + idx) + (long)(nb + q) * ldp] : 0.0; }
            if (live) {
                for (int sp = w; sp < nsplit; sp += XF_SL) tt += tmp[(long)sp * ldt + i + 1 + idx];
                if (w == 0) ar = A[(i + 1 + idx) + (long)(i + 1) * lda];
            }
        }
        const int r = i + 1 + idx;
        double corr = 0.0, sub = 0.0;
#pragma unroll
        for (int z = 0; z < ZV; ++z) { const int q = w + XF_SL * z; if (q <= k) { corr += vk[z] * s_yTu[q]; sub += vk[z] * s_rowY[q]; } }
#pragma unroll
        for (int z = 0; z < ZX; ++z) { const int q = w + XF_SL * z; if (q < k) { corr += xk[z] * s_uTu[q]; sub += xk[z] * s_rowU[q]; } }
        s_red[0][w][lane] = corr; s_red[1][w][lane] = sub; s_red[2][w][lane] = tt;
        __syncthreads();
        if (w < 8) {
#pragma unroll
            for (int qn = 0; qn < 3; ++qn)
                s_red2[qn][w][lane] = (s_red[qn][4 * w][lane] + s_red[qn][4 * w + 1][lane]) +
                                      (s_red[qn][4 * w + 2][lane] + s_red[qn][4 * w + 3][lane]);
        }
        __syncthreads();
        if (w == 0) {
            double cc = 0.0, x = 0.0;
            if (live) {
                corr = 0.0; sub = 0.0; tt = 0.0;
#pragma unroll
                for (int z = 0; z < 8; ++z) { corr += s_red2[0][z][lane]; sub += s_red2[1][z][lane]; tt += s_red2[2][z][lane]; }
                x = 2.0 * ((tt + f.snu * ar) * f.inv - corr);
                P[r + (long)(nb + k) * ldp] = x;
                cc = ar - sub - x * ufirst;
                c[r] = cc;
            }
            s_c[lane] = cc;
            s_x[lane] = x;
            cc2 += cc * cc;
            if (blockIdx.x == 0 && rb == 0 && lane == 0) c[i] = 0.0;   // keeps the 128-bit loads of the next pass harmless
        }
        __syncthreads();
        const double cl = s_c[lane];
#pragma unroll
        for (int z = 0; z < ZV; ++z) {
            const int q = w + XF_SL * z;
            if (q <= k) {
                const double xv = (q < k) ? xk[z < ZX ? z : 0] : s_x[lane];
                dv[z] += vk[z] * cl;
                dx[z] += xv * cl;
            }
        }
    }
    // ---- this CTA's partial [V^T c' | X^T c' | c'.c'] (each warp owns its panel columns)
    double *out = dots1p + (long)blockIdx.x * S;
#pragma unroll
    for (int z = 0; z < ZV; ++z) {
        const int q = w + XF_SL * z;
        if (q <= k) {                                       // warp-uniform
            const double a2 = warp_sum(dv[z]), b2 = warp_sum(dx[z]);
            if (lane == 0) { out[q] = a2; out[nb + q] = b2; }
        }
    }
    if (w == 0) {
        const double a2 = warp_sum(cc2);
        if (lane == 0) out[2 * nb] = a2;
    }
}

// The same step for long columns (more rows than one wave of 32-row blocks holds): 128 rows per CTA, all of
// them in flight at once - warp w owns rows 32*(w&3).. of the block and the panel columns (w>>2) + 8z, z < 4
// (panel width <= 32) - instead of RB 32-row groups one after the other: one pass through the
// load -> reduce -> x,c' -> partial dots chain per CTA (finish_xf<4>: four).
constexpr int XW_ROWS = 128, XW_SL = 8, XW_Z = 4;
__global__ void __launch_bounds__(1024)
finish_xw_kernel(double *__restrict__ A, long lda, int i, int m, int n, int k, int nb,
                 double *__restrict__ P, long ldp, double *__restrict__ Q, long ldq,
                 double *__restrict__ c, const double *__restrict__ rv,
                 const double *__restrict__ tmp, long ldt, int nsplit,
                 const double *__restrict__ dots2p, int nparts2, double *__restrict__ beta, int nRowBlk,
                 int nGroups, double *__restrict__ dots1p)
{
    constexpr int S = DOT_SLOTS_C;
    __shared__ double s_d[S];
    __shared__ double s_yTu[NBMAX], s_uTu[NBMAX], s_rowY[NBMAX], s_rowU[NBMAX];
    __shared__ double s_red[3][XW_SL][XW_ROWS + 1];
    __shared__ double s_c[XW_ROWS], s_x[XW_ROWS];
    __shared__ double s_o[4][2 * 32 + 1];
    static_assert(15 * DOT_SLOTS_C <= 3 * XW_SL * (XW_ROWS + 1), "s_part does not fit into s_red");
    double(*s_part)[S] = reinterpret_cast<double(*)[S]>(&s_red[0][0][0]);
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int rq = w & 3, sl = w >> 2, rr = rq * 32 + lane;
    const int R = n - i - 1, Lb = m - i - 1;
    const bool rowblk = (int)blockIdx.x < nRowBlk;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    int idx = blockIdx.x * XW_ROWS + rr;
    bool live = rowblk && idx < Lb;
    double vk[XW_Z], xk[XW_Z], tt = 0.0, ar = 0.0;
    auto load_group = [&]() {
        tt = 0.0; ar = 0.0;
#pragma unroll
        for (int z = 0; z < XW_Z; ++z) {
            const int q = sl + XW_SL * z;
            vk[z] = (live && q <= k) ? P[(i + 1 + idx) + (long)q * ldp] : 0.0;
            xk[z] = (live && q < k) ? P[(i + 1 + idx) + (long)(nb + q) * ldp] : 0.0;
        }
        if (live) {
            // this slice's partials of the pass, ten in flight at a time (<= 80 clusters per round)
            for (int sp0 = sl; sp0 < nsplit; sp0 += 10 * XW_SL) {
                double tv[10];
#pragma unroll
                for (int u = 0; u < 10; ++u) {
                    const int sp = sp0 + u * XW_SL;
                    tv[u] = (sp < nsplit) ? tmp[(long)sp * ldt + i + 1 + idx] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 10; ++u) tt += tv[u];
            }
            if (sl == 0) ar = A[(i + 1 + idx) + (long)(i + 1) * lda];
        }
    };
    load_group();
    const double rf = rv[i + 1];
    double qy = 0.0, qu = 0.0;
    if (t <= k) qy = Q[(i + 1) + (long)t * ldq];
    if (t < k) qu = Q[(i + 1) + (long)(nb + t) * ldq];
    const int ne = 2 * k + 2;
    const int TPE = (ne <= 68) ? 15 : 7;
    {
        const int e = t / TPE, part = t - TPE * e;
        if (e < ne) {
            const int slot = (e <= k) ? e : (e <= 2 * k ? nb + (e - k - 1) : 2 * nb);
            const int chunk = (nparts2 + TPE - 1) / TPE, p0 = part * chunk, p1 = min(nparts2, p0 + chunk);
            constexpr int LB = 10;
            double a2 = 0.0;
            for (int base = p0; base < p1; base += LB) {
                double v[LB];
#pragma unroll
                for (int u = 0; u < LB; ++u) v[u] = (base + u < p1) ? dots2p[(long)(base + u) * S + slot] : 0.0;
#pragma unroll
                for (int u = 0; u < LB; ++u) a2 += v[u];
            }
            s_part[part][slot] = a2;
        }
    }
    __syncthreads();
    if (t < S && (t <= k || (t >= nb && t < nb + k) || t == 2 * nb)) {
        double a2 = 0.0;
        for (int pz = 0; pz < TPE; ++pz) a2 += s_part[pz][t];
        s_d[t] = a2;
    }
    __syncthreads();
    const Refl f = make_refl(rf, s_d[2 * nb]);
    const double ufirst = (rf + f.snu) * f.inv;
    if (!rowblk) {
        const int cidx = (blockIdx.x - nRowBlk) * 1024 + t;
        if (cidx == 0) beta[i] = -f.snu;
        if (cidx < R) {
            const int j = i + 1 + cidx;
            const double u = (rv[j] + (cidx == 0 ? f.snu : 0.0)) * f.inv;
            A[i + (long)j * lda] = u;
            Q[j + (long)(nb + k) * ldq] = u;
        }
        return;
    }
    if (t <= k) { s_rowY[t] = qy; s_yTu[t] = (s_d[t] + f.snu * qy) * f.inv; }
    if (t < k) { s_rowU[t] = qu; s_uTu[t] = (s_d[nb + t] + f.snu * qu) * f.inv; }
    __syncthreads();

    double dv[XW_Z], dx[XW_Z], cc2 = 0.0;
#pragma unroll
    for (int z = 0; z < XW_Z; ++z) { dv[z] = 0.0; dx[z] = 0.0; }
#pragma unroll 1
    for (int g = blockIdx.x; g < nGroups; g += nRowBlk) {
        if (g != (int)blockIdx.x) {
            idx = g * XW_ROWS + rr;
            live = idx < Lb;
            load_group();
        }
        const int r = i + 1 + idx;
        double corr = 0.0, sub = 0.0;
#pragma unroll
        for (int z = 0; z < XW_Z; ++z) {
            const int q = sl + XW_SL * z;
            if (q <= k) { corr += vk[z] * s_yTu[q]; sub += vk[z] * s_rowY[q]; }
            if (q < k) { corr += xk[z] * s_uTu[q]; sub += xk[z] * s_rowU[q]; }
        }
        s_red[0][sl][rr] = corr; s_red[1][sl][rr] = sub; s_red[2][sl][rr] = tt;
        __syncthreads();
        if (sl == 0) {                                      // warps 0..3: one row each
            double cc = 0.0, x = 0.0;
            if (live) {
                corr = 0.0; sub = 0.0; tt = 0.0;
#pragma unroll
                for (int z = 0; z < XW_SL; ++z) { corr += s_red[0][z][rr]; sub += s_red[1][z][rr]; tt += s_red[2][z][rr]; }
                x = 2.0 * ((tt + f.snu * ar) * f.inv - corr);
                P[r + (long)(nb + k) * ldp] = x;
                cc = ar - sub - x * ufirst;
                c[r] = cc;
            }
            s_c[rr] = cc;
            s_x[rr] = x;
            cc2 += cc * cc;
            if (g == 0 && t == 0) c[i] = 0.0;               // keeps the 128-bit loads of the next pass harmless
        }
        __syncthreads();
        const double cl = s_c[rr];
#pragma unroll
        for (int z = 0; z < XW_Z; ++z) {
            const int q = sl + XW_SL * z;
            if (q <= k) {
                const double xv = (q < k) ? xk[z] : s_x[rr];
                dv[z] += vk[z] * cl;
                dx[z] += xv * cl;
            }
        }
    }
    // ---- this CTA's partial [V^T c' | X^T c' | c'.c']: per warp, then over the four row quarters
#pragma unroll
    for (int z = 0; z < XW_Z; ++z) {
        const int q = sl + XW_SL * z;
        if (q <= k) {                                       // warp-uniform
            const double a2 = warp_sum(dv[z]), b2 = warp_sum(dx[z]);
            if (lane == 0) { s_o[rq][q] = a2; s_o[rq][32 + q] = b2; }
        }
    }
    if (sl == 0) {
        const double a2 = warp_sum(cc2);
        if (lane == 0) s_o[rq][64] = a2;
    }
    __syncthreads();
    double *out = dots1p + (long)blockIdx.x * S;
    if (t < 65) {
        const int q = t & 31;
        if (t == 64 || q <= k) {
            const double a2 = (s_o[0][t] + s_o[1][t]) + (s_o[2][t] + s_o[3][t]);
            out[t == 64 ? 2 * nb : (t < 32 ? q : nb + q)] = a2;
        }
    }
}

__global__ void col_init_kernel(const double *__restrict__ A, int m, long lda, double *__restrict__ c)
{
    long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < lda) c[r] = (r < m) ? A[r] : 0.0;
}

} // namespace svdgpu
#include "bidiag_fused.cuh"
#include "bidiag_tail.cuh"
#include "bidiag_panel.cuh"
namespace svdgpu {
static_assert(TAIL_WS_SLOTS == TL_WS_SLOTS, "workspace sizing of the on-chip tail");

// ---------------------------------------------------------------------------------------
constexpr int TMPN_ROWS = (BIDIAG_MAX_SPLIT > FZ_MAX_CLUSTERS) ? BIDIAG_MAX_SPLIT : FZ_MAX_CLUSTERS;
constexpr int DOT_SLOTS = 2 * NBMAX + 2;

size_t bidiag_workspace_bytes(int m, int n, long lda)
{
    long ldq = round_up(n, 2);
    size_t d = 0;
    d += (size_t)lda * 2 * NBMAX;          // P
    d += (size_t)ldq * 2 * NBMAX;          // Q
    d += (size_t)lda;                      // c
    d += (size_t)ldq + 2;                  // rv
    d += (size_t)BIDIAG_MAX_SPLIT * ldq;   // tmpT
    d += (size_t)TMPN_ROWS * lda;          // tmpN
    d += 2 * DOT_SLOTS;                    // dots1, dots2 (final)
    d += (size_t)(ceil_div(m, 128) + 130) * DOT_SLOTS; // dots1 partials (finish_xf row blocks)
    d += (size_t)FZ_MAX_CLUSTERS * DOT_SLOTS;          // dots2 partials (fused pass clusters)
    d += 8;                                // counters
    d += 2 * TAIL_WS_SLOTS + 2;            // tagged exchange slots of the on-chip tail (16 bytes each, 16-byte aligned)
    return d * sizeof(double);
}

struct BidiagBufs {
    double *P, *Q, *c, *rv, *tmpT, *tmpN, *dots1, *dots2, *dots1p, *dots2p;
    unsigned *counters;
    void *tail;
    long ldp, ldq;
};
static BidiagBufs carve(void *workspace, int m, int n, long lda)
{
    BidiagBufs b;
    b.ldp = lda; b.ldq = round_up(n, 2);
    double *w = (double *)workspace;
    b.P = w;      w += (size_t)b.ldp * 2 * NBMAX;
    b.Q = w;      w += (size_t)b.ldq * 2 * NBMAX;
    b.c = w;      w += (size_t)lda;
    b.rv = w;     w += (size_t)b.ldq + 2;
    b.tmpT = w;   w += (size_t)BIDIAG_MAX_SPLIT * b.ldq;
    b.tmpN = w;   w += (size_t)TMPN_ROWS * lda;
    b.dots1 = w;  w += DOT_SLOTS;
    b.dots2 = w;  w += DOT_SLOTS;
    b.dots1p = w; w += (size_t)(ceil_div(m, 128) + 130) * DOT_SLOTS;
    b.dots2p = w; w += (size_t)FZ_MAX_CLUSTERS * DOT_SLOTS;
    b.counters = (unsigned *)w;             w += 8;
    b.tail = (void *)(((uintptr_t)w + 15) & ~(uintptr_t)15);
    return b;
}

// gemvT launch for step i with k panel columns; returns the number of row splits written
static int launch_gemvT(const double *A, long lda, int i, int m, int n, int mpad, const BidiagBufs &b,
                        int nb, int k, int target, cudaStream_t st)
{
    const int R = n - i - 1;
    int colGroups = (R > 0) ? ceil_div(R, GT_WARPS * GT_CW) : 0;
    int rows = mpad - (i & ~1);
    int rowsPerSplit = rows, nsplit = 1;
    if (colGroups > 0) {
        nsplit = target / colGroups;
        if (nsplit < 1) nsplit = 1;
        if (nsplit > BIDIAG_MAX_SPLIT) nsplit = BIDIAG_MAX_SPLIT;
        rowsPerSplit = (int)round_up(ceil_div(rows, nsplit), 256);
        nsplit = ceil_div(rows, rowsPerSplit);
    }
    dim3 grid(colGroups + 2 * k + 1, nsplit);
    gemvT_kernel<<<grid, GT_WARPS * 32, 0, st>>>(A, lda, i, m, n, mpad, b.c, b.tmpT, b.ldq, colGroups,
                                                 rowsPerSplit, b.P, b.ldp, nb, k, b.dots1);
    SVD_KERNEL_CHECK();
    return colGroups == 0 ? 0 : nsplit;
}

static int launch_gemvN(const double *A, long lda, int i, int m, int n, int mpad, const BidiagBufs &b,
                        int nb, int k, int target, cudaStream_t st)
{
    const int R = n - i - 1, Lb = m - i - 1;
    int rows = (Lb > 0) ? mpad - ((i + 1) & ~1) : 0;
    int rowBlocks = (rows > 0) ? ceil_div(rows, 2 * GN_THREADS) : 0;
    int colsPerSplit = R > 0 ? R : 1, nsplit = 1;
    if (rowBlocks > 0) {
        nsplit = target / rowBlocks;
        if (nsplit < 1) nsplit = 1;
        if (nsplit > BIDIAG_MAX_SPLIT) nsplit = BIDIAG_MAX_SPLIT;
        colsPerSplit = (int)round_up(ceil_div(R, nsplit), 32);
        nsplit = ceil_div(R, colsPerSplit);
    }
    dim3 grid(rowBlocks + 2 * k + 2, nsplit);
    gemvN_kernel<<<grid, GN_THREADS, 0, st>>>(A, lda, i, m, n, mpad, b.rv, b.tmpN, lda, rowBlocks,
                                                 colsPerSplit, b.Q, b.ldq, nb, k, b.dots2);
    SVD_KERNEL_CHECK();
    return rowBlocks == 0 ? 0 : nsplit;
}

static int sm_targets(int &targetT, int &targetN)
{
    int dev = 0, nsm = 148;
    SVD_CUDA_CHECK(cudaGetDevice(&dev));
    SVD_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    targetT = nsm * 4; targetN = nsm * 8;
    return nsm;
}


// ---- fused pass launch (cluster launch, TMA/mbarrier kernel of bidiag_fused.cuh) ---------------
struct FusedPlan { bool ok; int CS, RPT, Lc, T, NC, var; };
// kernel variants: 0 <2,4,6>  1 <4,2,6>  2 <8,1,6> (classic: 8/RPT columns per 32 KB stage)  3 <8,2,4> (two columns of up to
// 3072 rows per 48 KB stage)
constexpr int FZ_NVAR = 4;
static const int fz_var_rpt[FZ_NVAR] = {2, 4, 8, 8}, fz_var_cbw[FZ_NVAR] = {4, 2, 1, 2};
static int g_max_clusters[FZ_NVAR][FZ_MAXCS + 1];   // [variant][cluster size]: co-resident clusters (occupancy API)
static bool g_fz_wide = false;          // SVD_GPU_FZ_WIDE=1: variant 3 for 2048 < rows per CTA <= 3072 (measured neutral to
                                        // -0.6 %: the cost of a tile scales with its columns, profiles/r02_ab_helper_warps.log)
static int force_cs = 0;
static FusedPlan plan_fused(int i, int m, int n, int mpad, int nsm, int min_rows, int min_cols, bool classic = false)
{
    FusedPlan p = {false, 1, 4, 0, 0, 0, 1};
    const int L = m - i, R = n - i - 1;
    if (L < min_rows || R < min_cols) return p;
    const int Ltot = mpad - (i & ~1);
    // Cluster size: the smallest one whose row slice fits a stage is not always the best.  Clusters
    // are placed inside one GPC, so only g_max_clusters[..][CS] of them are co-resident (measured by
    // the occupancy API, e.g. fewer than 148/4 for CS = 4); a grid with more clusters than that runs
    // in two waves.  Among the sizes whose row slice fits a stage pick the one that keeps the most SMs busy.
    double best = -1.0;
    for (int CS = 1; CS <= FZ_MAXCS; ++CS) {
        const int Lc = (int)round_up(ceil_div(Ltot, CS), 2);
        if (Lc > FZ_STAGE) continue;
        if (force_cs > 0 && CS != force_cs) continue;
        int ri = Lc <= 1024 ? 0 : Lc <= 2048 ? 1 : 2;
        if (ri == 2 && Lc <= 3072 && g_fz_wide && !classic) ri = 3;
        const int RPT = fz_var_rpt[ri], cbw = fz_var_cbw[ri];
        const int T = ceil_div(R, cbw);
        int maxc = g_max_clusters[ri][CS];
        if (maxc <= 0) continue;
        if (maxc > FZ_MAX_CLUSTERS) maxc = FZ_MAX_CLUSTERS;
        const int NC = T < maxc ? T : maxc;
        // SMs kept streaming; ties go to the smaller cluster (shorter exchange, larger tiles)
        const double score = (double)NC * CS - 1e-3 * CS;
        if (score > best) {
            best = score;
            p.CS = CS; p.Lc = Lc; p.RPT = RPT; p.T = T; p.NC = NC; p.var = ri;
        }
    }
    p.ok = (best > 0.0 && p.NC >= 1);
    return p;
}

// Programmatic dependent launch between the fused pass and finish_xf.  g_pdl_mode (SVD_GPU_PDL): 0 never,
// 1 (default) only for steps whose pass runs single-CTA "clusters" (trailing rows <= 4096), 2 always.  Measured
// on a B200 (profiles/r02_bidiag_pdl.log): the early launch wins ~4 us per step where it applies (4096^2: 89.6 ->
// 81.7 ms), but with real clusters (CS >= 2, placed inside one GPC) the early finish_xf CTAs fragment the SMs
// the next pass's clusters need and part of them runs as a second wave: 16384^2 2.78 s with PDL everywhere,
// 2.55 s without it.
static int g_pdl_mode = 1;
static int g_pdl_cs = 2;        // largest cluster size whose pass is launched as a programmatic dependent (SVD_GPU_PDL_CS)
static int g_pdl_fin = 1;       // finish_xf as a programmatic dependent of a clustered pass (SVD_GPU_PDL_FIN)
static bool g_pdl = true;       // the decision for the pass launch of the current step
static bool g_pdl_f = true;     // ... and for its finish_xf
template <int RPT, int CBW, int NST> static void launch_fused_t(const FusedArgs &fa, const FusedPlan &pl, cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.NC * pl.CS);
    cfg.blockDim = dim3(FZ_THREADS);
    cfg.dynamicSmemBytes = FZ_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = pl.CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_pdl ? 2 : 1;
    SVD_CUDA_CHECK(cudaLaunchKernelEx(&cfg, fused_pass_kernel<RPT, CBW, NST>, fa));
    SVD_KERNEL_CHECK();
}
// finish_xf as the programmatic dependent of the fused pass
template <int RB, typename... Args> static void launch_finish_xf(int grid, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_pdl_f ? 1 : 0;
    SVD_CUDA_CHECK(cudaLaunchKernelEx(&cfg, finish_xf_kernel<RB>, args...));
}
template <typename... Args> static void launch_finish_xw(int grid, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_pdl_f ? 1 : 0;
    SVD_CUDA_CHECK(cudaLaunchKernelEx(&cfg, finish_xw_kernel, args...));
}
static void launch_fused(const FusedArgs &fa, const FusedPlan &pl, cudaStream_t st)
{
    switch (pl.var) {
    case 0: launch_fused_t<2, 4, 6>(fa, pl, st); break;
    case 1: launch_fused_t<4, 2, 6>(fa, pl, st); break;
    case 3: launch_fused_t<8, 2, 4>(fa, pl, st); break;
    default: launch_fused_t<8, 1, 6>(fa, pl, st); break;
    }
}
template <int RPT, int CBW, int NST> static void query_clusters_t(int ri)
{
    for (int CS = 1; CS <= FZ_MAXCS; ++CS) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CS * 148);
        cfg.blockDim = dim3(FZ_THREADS);
        cfg.dynamicSmemBytes = FZ_SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, fused_pass_kernel<RPT, CBW, NST>, &cfg) != cudaSuccess) { nc = 0; (void)cudaGetLastError(); }
        g_max_clusters[ri][CS] = nc;
    }
}
static void fused_set_attributes()
{
    static DeviceOnce once;
    if (!first_on_device(once)) return;
    SVD_CUDA_CHECK(cudaFuncSetAttribute(fused_pass_kernel<2, 4, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FZ_SMEM_BYTES));
    SVD_CUDA_CHECK(cudaFuncSetAttribute(fused_pass_kernel<4, 2, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FZ_SMEM_BYTES));
    SVD_CUDA_CHECK(cudaFuncSetAttribute(fused_pass_kernel<8, 1, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FZ_SMEM_BYTES));
    SVD_CUDA_CHECK(cudaFuncSetAttribute(fused_pass_kernel<8, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FZ_SMEM_BYTES));
    query_clusters_t<2, 4, 6>(0);
    query_clusters_t<4, 2, 6>(1);
    query_clusters_t<8, 1, 6>(2);
    query_clusters_t<8, 2, 4>(3);
    if (getenv("SVD_GPU_VERBOSE")) {
        fprintf(stderr, "fused pass: co-resident clusters by size 1..%d (RPT 8):", FZ_MAXCS);
        for (int CS = 1; CS <= FZ_MAXCS; ++CS) fprintf(stderr, " %d", g_max_clusters[2][CS]);
        fprintf(stderr, "\n");
    }
    const char *fc = getenv("SVD_GPU_FUSED_CS");       // experiments: smallest cluster size considered
    if (fc) force_cs = atoi(fc);
}

// ---- persistent per-panel kernel (bidiag_panel.cuh)
static int g_panel_ctas = -1;           // co-resident CTAs of the panel kernel (0: not available)
static void panel_init(int nsm)
{
    static DeviceOnce once;
    if (!first_on_device(once)) return;
    int per = 0;
    g_panel_ctas = 1 << 30;
    const void *fns[3] = {(const void *)panel_kernel<2>, (const void *)panel_kernel<4>, (const void *)panel_kernel<8>};
    for (int z = 0; z < 3; ++z) {
        if (cudaFuncSetAttribute(fns[z], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FZ_SMEM_BYTES) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, fns[z], FZ_THREADS, FZ_SMEM_BYTES) != cudaSuccess) {
            (void)cudaGetLastError();
            per = 0;
        }
        g_panel_ctas = std::min(g_panel_ctas, per * nsm);
    }
    int coop = 0, dev = 0;
    SVD_CUDA_CHECK(cudaGetDevice(&dev));
    SVD_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) g_panel_ctas = 0;
}
static void launch_panel(const PanelArgs &pa, int RPT, cudaStream_t st)
{
    SVD_CUDA_CHECK(cudaMemsetAsync(pa.bar, 0, sizeof(unsigned), st));
    void *args[] = {(void *)&pa};
    const void *fn = RPT == 2 ? (const void *)panel_kernel<2> : RPT == 4 ? (const void *)panel_kernel<4> : (const void *)panel_kernel<8>;
    SVD_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(pa.NC), dim3(FZ_THREADS), args, FZ_SMEM_BYTES, st));
    SVD_KERNEL_CHECK();
}

// ---- on-chip tail (bidiag_tail.cuh): from step i on, if the trailing block fits the SMs' shared memory
static int g_tail_ctas = -1;            // co-resident CTAs of the tail kernel (0: not available)
static bool g_tail2_ok = false;
static int g_tail_mode = 2;             // SVD_GPU_TAIL: 2 = kernel version 2 (default), 1 = version 1 (kept for cross-checks)
// does the trailing block of step i fit the shared memory of `ctas` CTAs (columns dealt round-robin)?
static bool tail_fits_ctas(int m, int n, int i, int ctas)
{
    if (m < n || ctas <= 0) return false;
    const int L0 = m - i, R0 = n - i;
    if (L0 > TL_MAXROWS || R0 < 1) return false;
    const long Lp = round_up(L0, 2);
    const int cpc = ceil_div(R0, ctas);
    // one warp per owned row below the pivot, and the last warp never owns one (it polls RR / R1): with fewer
    // co-resident CTAs than a full B200 (MIG slice, cut-down part) tall blocks must stay on the streaming path
    if (L0 > 1 && ceil_div(L0 - 1, ctas) > TL_WARPS - 1) return false;
    return cpc <= TL_CPC && (long)cpc * Lp <= TL_CAP;
}
static bool tail_fits(int m, int n, int i) { return tail_fits_ctas(m, n, i, g_tail_ctas); }
// First step handed to the on-chip tail kernel (a panel boundary: multiples of nb), min(m,n) if none.
// Pure planning, no device needed (used by bench.py's byte accounting and by the CPU tests).
int bidiag_tail_start(int m, int n, int nb, int ctas)
{
    const int mn = m < n ? m : n;
    if (nb <= 0 || nb > NBMAX) nb = 32;
    if (ctas > TL_MAXG) ctas = TL_MAXG;
    for (int i = 0; i < mn; i += nb)
        if (tail_fits_ctas(m, n, i, ctas)) return i;
    return mn;
}
static void tail_init()
{
    // (function attributes are per device; the co-resident CTA count is the same on every device of a box)
    static DeviceOnce once;
    if (!first_on_device(once)) return;
    g_tail_ctas = 0;
    int dev = 0, nsm = 0, coop = 0, per_sm = 0;
    SVD_CUDA_CHECK(cudaGetDevice(&dev));
    SVD_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    SVD_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    const int smem = (TL_CAP + TL_AUX) * (int)sizeof(double);
    if (!coop) return;
    // the experimental version 2 must not be able to disable the default one
    int per_sm2 = 0;
    g_tail2_ok = cudaFuncSetAttribute(bidiag_tail_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess &&
                 cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, bidiag_tail_kernel<2>, TL_THREADS, smem) == cudaSuccess &&
                 per_sm2 >= 1;
    (void)cudaGetLastError();
    if (cudaFuncSetAttribute(bidiag_tail_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
        (void)cudaGetLastError();
        return;
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bidiag_tail_kernel<1>, TL_THREADS, smem) != cudaSuccess ||
        per_sm < 1) {
        (void)cudaGetLastError();
        return;
    }
    g_tail_ctas = nsm < TL_MAXG ? nsm : TL_MAXG;
}
static void launch_tail(int m, int n, int i, double *A, long lda, double *alpha, double *beta, const BidiagBufs &b,
                        cudaStream_t st)
{
    TailArgs ta;
    ta.A = A; ta.lda = lda; ta.m = m; ta.n = n; ta.i0 = i; ta.alpha = alpha; ta.beta = beta;
    TlSlot *sl = reinterpret_cast<TlSlot *>(b.tail);
    ta.W = sl;   sl += (size_t)TL_MAXG * TL_MAXROWS;
    ta.X = sl;   sl += TL_MAXROWS;
    ta.C = sl;   sl += TL_MAXROWS;
    ta.A1 = sl;  sl += TL_MAXROWS;
    ta.RR = sl;  sl += 2 * TL_MAXG;
    ta.R1 = sl;
    ta.counter = b.counters + 8;
    // tags start at 1: a zeroed slot is never valid
    SVD_CUDA_CHECK(cudaMemsetAsync(b.tail, 0, TL_WS_SLOTS * sizeof(TlSlot), st));
    ta.Lp = (int)round_up(m - i, 2);
    const int cpc = ceil_div(n - i, g_tail_ctas);
    const size_t smem = ((size_t)cpc * ta.Lp + TL_AUX) * sizeof(double);
    SVD_CUDA_CHECK(cudaMemsetAsync(ta.counter, 0, sizeof(unsigned), st));
    void *args[] = {&ta};
    const void *fn = (g_tail_mode == 2 && g_tail2_ok) ? (const void *)bidiag_tail_kernel<2> : (const void *)bidiag_tail_kernel<1>;
    SVD_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(g_tail_ctas), dim3(TL_THREADS), args, smem, st));
    SVD_KERNEL_CHECK();
}

// SVD_GPU_BIDIAG_PROFILE=1: CUDA events around every launch of the three streaming kernel classes (fused pass,
// finish_xf, panel GEMM), summed per class and per quarter of the factorization at the end (synchronises; run it
// with SVD_GPU_PDL=0 so that the classes do not overlap).  A measuring aid, off by default.
struct BdProf {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> cls, step;
    void begin(int c, int i, cudaStream_t st) {
        if (!on) return;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st); ev.push_back(a); ev.push_back(b); cls.push_back(c); step.push_back(i);
    }
    void end(cudaStream_t st) { if (on) cudaEventRecord(ev.back(), st); }
    void report(int mn, cudaStream_t st) {
        if (!on) return;
        cudaStreamSynchronize(st);
        double tot[3][4] = {}; int cnt[3][4] = {};
        for (size_t k = 0; k < cls.size(); ++k) {
            float ms = 0.f; cudaEventElapsedTime(&ms, ev[2 * k], ev[2 * k + 1]);
            const int q = (int)((long)step[k] * 4 / (mn > 0 ? mn : 1)); tot[cls[k]][q < 4 ? q : 3] += ms; cnt[cls[k]][q < 4 ? q : 3]++;
            cudaEventDestroy(ev[2 * k]); cudaEventDestroy(ev[2 * k + 1]);
        }
        const char *nm[3] = {"fused_pass", "finish_xf", "panel_gemm"};
        for (int c = 0; c < 3; ++c)
            fprintf(stderr, "BDPROF %-10s by quarter of the steps: %8.2f %8.2f %8.2f %8.2f ms  (launches %d %d %d %d)\n", nm[c],
                    tot[c][0], tot[c][1], tot[c][2], tot[c][3], cnt[c][0], cnt[c][1], cnt[c][2], cnt[c][3]);
    }
};

void bidiag_device(int m, int n, double *A, long lda, double *alpha, double *beta, void *workspace,
                   int nb, cudaStream_t st, const ProgressHook *hook)
{
    BdProf prof;
    prof.on = getenv("SVD_GPU_BIDIAG_PROFILE") != nullptr;
    if (nb <= 0 || nb > NBMAX) nb = 32;
    const int mn = (m < n) ? m : n;
    const int mpad = (int)round_up(m, 2);
    if (lda < mpad || (lda & 1)) {
        fprintf(stderr, "bidiag_device: lda (%ld) must be even and >= round_up(m,2)\n", lda);
        abort();
    }
    const BidiagBufs b = carve(workspace, m, n, lda);
    int targetT, targetN;
    const int nsm = sm_targets(targetT, targetN);
    // SVD_GPU_FUSED=1 selects the single-read fused pass (bidiag_fused.cuh); the default is decided by
    // measurement (profiles/): the split passes unless the fused pass is faster at this size
    const char *fenv = getenv("SVD_GPU_FUSED");
    const bool use_fused = (fenv != nullptr) ? (fenv[0] != '0') : FZ_DEFAULT_ON;
    { const char *pe = getenv("SVD_GPU_PDL"); g_pdl_mode = pe ? atoi(pe) : 1; g_pdl = g_pdl_f = g_pdl_mode != 0; }
    { const char *pe = getenv("SVD_GPU_PDL_CS"); if (pe) g_pdl_cs = atoi(pe); }
    { const char *pe = getenv("SVD_GPU_PDL_FIN"); if (pe) g_pdl_fin = atoi(pe); }
    if (use_fused) fused_set_attributes();
    // thresholds below which the split passes are used (overridable for tests)
    const char *e1 = getenv("SVD_GPU_FUSED_MIN_ROWS"), *e2 = getenv("SVD_GPU_FUSED_MIN_COLS");
    const int fz_min_rows = e1 ? atoi(e1) : FZ_MIN_ROWS, fz_min_cols = e2 ? atoi(e2) : FZ_MIN_COLS;
    // SVD_GPU_TAIL=0/1: finish on chip once the trailing block fits into shared memory (bidiag_tail.cuh)
    g_fz_wide = getenv("SVD_GPU_FZ_WIDE") ? atoi(getenv("SVD_GPU_FZ_WIDE")) != 0 : false;
    const int xw_mode = getenv("SVD_GPU_XW") ? atoi(getenv("SVD_GPU_XW")) : 1;     // finish_xw: 0 never, 1 long columns, 2 always
    // SVD_GPU_PPK=1: whole panels of single-CTA passes in one persistent cooperative launch (bidiag_panel.cuh)
    const bool use_ppk = use_fused && (getenv("SVD_GPU_PPK") ? atoi(getenv("SVD_GPU_PPK")) != 0 : PPK_DEFAULT_ON);
    if (use_ppk) panel_init(nsm);
    const char *tenv = getenv("SVD_GPU_TAIL");
    const bool use_tail = tenv ? (tenv[0] != '0') : (TAIL_DEFAULT_ON != 0);
    g_tail_mode = (tenv && tenv[0] == '1') ? 1 : 2;
    if (use_tail) tail_init();
    bool dots1_ready = false;       // dots1p holds dots1_parts partial dot vectors of the current column c
    int dots1_parts = 0;
    // Panel width schedule.  The deferred trailing update is a K = 2nb GEMM that reads and writes the
    // trailing block once per panel: with nb = 32 it is bound by that C traffic and by per-tile
    // latency, with nb = 64 by the DMMA pipe.  The per-step panel corrections grow with nb, which only
    // pays while the trailing block is large: panels that start with >= nb_big_min trailing rows and
    // columns use nb_big, later ones the caller's nb.  (The panel buffers are sized for NBMAX.)
    const int nb_small = nb;
    int nb_big = nb, nb_big_min = 1 << 30;
    {
        const char *eb = getenv("SVD_GPU_NB_BIG"), *em = getenv("SVD_GPU_NB_BIG_MIN");
        if (eb) { nb_big = atoi(eb); nb_big_min = em ? atoi(em) : NB_BIG_MIN_DEFAULT; }
        else if (NB_BIG_DEFAULT > 0 && nb == 32) { nb_big = NB_BIG_DEFAULT; nb_big_min = em ? atoi(em) : NB_BIG_MIN_DEFAULT; }
        if (nb_big <= 0 || nb_big > NBMAX) nb_big = nb;
    }

    SVD_CUDA_CHECK(cudaMemsetAsync(b.rv, 0, sizeof(double) * ((size_t)b.ldq + 2), st));
    SVD_CUDA_CHECK(cudaMemsetAsync(b.dots1, 0, sizeof(double) * 2 * DOT_SLOTS, st));
    SVD_CUDA_CHECK(cudaMemsetAsync(b.counters, 0, 8 * sizeof(double), st));
    col_init_kernel<<<ceil_div(lda, 256), 256, 0, st>>>(A, m, lda, b.c);
    SVD_KERNEL_CHECK();

    int k = 0;
    for (int i = 0; i < mn; ++i) {              // steps 0..mn-2 regular, step mn-1 is the tail
        const bool tail = (i == mn - 1);
        if (use_tail && k == 0 && tail_fits(m, n, i)) {
            // the rest of the factorization in one cooperative launch with the matrix on chip
            launch_tail(m, n, i, A, lda, alpha, beta, b, st);
            break;
        }
        // reflectors [0, i) are final (column j and row j are last written by step j)
        if (hook && hook->fn && i > 0 && hook->every > 0 && i % hook->every == 0) hook->fn(hook->user, i, st);
        if (k == 0) {
            const int nb_next = (m - i >= nb_big_min && n - i >= nb_big_min) ? nb_big : nb_small;
            // the dot slots finish_xf left for this step are laid out for the previous panel's width
            if (nb_next != nb) dots1_ready = false;
            nb = nb_next;
        }
        const int do_col = tail ? (n >= m + 1 ? 0 : 1) : 1;
        const int do_row = tail ? (n >= m + 1 ? 1 : 0) : (i < n - 2 ? 1 : 0);
        const int R = n - i - 1, Lb = m - i - 1;
        int nsplitT = 0, nsplitN = 0;

        FusedPlan pl = {false, 1, 4, 0, 0, 0};
        if (use_fused && !tail && do_col && do_row) pl = plan_fused(i, m, n, mpad, nsm, fz_min_rows < 2 ? 2 : fz_min_rows, fz_min_cols < 1 ? 1 : fz_min_cols);
        // a whole panel at once: every step of it a regular single-CTA fused step with at least one tile per CTA
        if (pl.ok && use_ppk && k == 0 && pl.CS == 1) pl = plan_fused(i, m, n, mpad, nsm, 2, 1, true);   // the panel kernel has the classic tiles
        if (pl.ok && use_ppk && k == 0 && pl.CS == 1 && nb <= 32 && g_panel_ctas > 0 && pl.NC <= g_panel_ctas &&
            i + nb < mn - 1 && i + nb < n - 2 && n - (i + nb) >= pl.NC * FZ_CBW_MAX && m - (i + nb) >= 64 &&
            ceil_div(Lb, 32) < pl.NC && (!hook || !hook->fn || hook->every <= 0 || hook->every % nb == 0)) {
            if (!dots1_ready) {
                gemvT_kernel<<<dim3(1, 1), GT_WARPS * 32, 0, st>>>(A, lda, i, m, n, mpad, b.c, b.tmpT, b.ldq,
                                                                   0, 0, b.P, b.ldp, nb, 0, b.dots1);
                SVD_KERNEL_CHECK();
            }
            PanelArgs pa;
            pa.A = A; pa.lda = lda; pa.i0 = i; pa.k0 = 0; pa.nsteps = nb; pa.m = m; pa.n = n; pa.mpad = mpad; pa.nb = nb;
            pa.P = b.P; pa.ldp = b.ldp; pa.Q = b.Q; pa.ldq = b.ldq; pa.c = b.c; pa.rv = b.rv;
            pa.tmpN = b.tmpN; pa.ldt = lda;
            pa.dots1 = dots1_ready ? b.dots1p : b.dots1; pa.nparts1 = dots1_ready ? dots1_parts : 0;
            pa.dots1p = b.dots1p; pa.dots2p = b.dots2p; pa.alpha = alpha; pa.beta = beta;
            pa.NC = pl.NC; pa.Lc = pl.Lc; pa.bar = b.counters + 12;
            pa.trace = nullptr;
            static unsigned long long *d_ptrace = nullptr;
            const char *pte = getenv("SVD_GPU_PPK_TRACE");
            const bool ptracing = pte && atoi(pte) == i;
            if (ptracing) {
                if (!d_ptrace) SVD_CUDA_CHECK(cudaMalloc(&d_ptrace, 8 * (NBMAX + 64) * sizeof(unsigned long long)));
                SVD_CUDA_CHECK(cudaMemsetAsync(d_ptrace, 0, 8 * (NBMAX + 64) * sizeof(unsigned long long), st));
                pa.trace = d_ptrace;
            }
            prof.begin(0, i, st);
            launch_panel(pa, pl.RPT, st);
            prof.end(st);
            if (ptracing) {
                static unsigned long long h[8 * (NBMAX + 64)];
                SVD_CUDA_CHECK(cudaMemcpyAsync(h, d_ptrace, sizeof h, cudaMemcpyDeviceToHost, st));
                SVD_CUDA_CHECK(cudaStreamSynchronize(st));
                fprintf(stderr, "PPKTRACE panel at step %d RPT %d Lc %d NC %d: per step, cycles since the step's start: prologue end | "
                                "pass done in CTA 0 | barrier 1 passed | finish done | barrier 2 passed | first TMA | last TMA | step length\n",
                        i, pl.RPT, pl.Lc, pl.NC);
                for (int z = 0; z < nb; ++z) {
                    const unsigned long long t0 = h[z * 8];
                    fprintf(stderr, "PPKTRACE %2d", z);
                    for (int q = 1; q < 8; ++q) fprintf(stderr, " %7lld", h[z * 8 + q] ? (long long)(h[z * 8 + q] - t0) : -1ll);
                    fprintf(stderr, " %7lld\n", z + 1 < nb ? (long long)(h[(z + 1) * 8] - t0) : -1ll);
                }
                fprintf(stderr, "PPKTILE step %d, per tile of CTA 0, cycles since that step's start: TMA issue | sweep-1 start | sweep-1 done | "
                                "reducer | finisher start | finisher done | sweep-2 start | sweep-2 done\n", i + 1);
                for (int nt = 0; nt < 64 && h[(NBMAX + nt) * 8 + 1]; ++nt) {
                    fprintf(stderr, "PPKTILE %2d", nt);
                    for (int q = 0; q < 8; ++q)
                        fprintf(stderr, " %7lld", h[(NBMAX + nt) * 8 + q] ? (long long)(h[(NBMAX + nt) * 8 + q] - h[8]) : -1ll);
                    fprintf(stderr, "\n");
                }
            }
            i += nb - 1;
            k = nb - 1;
            dots1_ready = true;
            dots1_parts = ceil_div(m - i - 1, 32);
        } else
        if (pl.ok) {
            g_pdl = (g_pdl_mode == 2) || (g_pdl_mode == 1 && pl.CS <= g_pdl_cs);
            g_pdl_f = (g_pdl_mode == 2) || (g_pdl_mode == 1 && (pl.CS <= g_pdl_cs || g_pdl_fin));
            // ---- fused step: ONE read of the trailing matrix gives both A^T c and A r
            if (!dots1_ready) {
                // only the panel dots / norm of c (the "extra" CTAs of gemvT, no column groups)
                gemvT_kernel<<<dim3(2 * k + 1, 1), GT_WARPS * 32, 0, st>>>(A, lda, i, m, n, mpad, b.c, b.tmpT, b.ldq,
                                                                           0, 0, b.P, b.ldp, nb, k, b.dots1);
                SVD_KERNEL_CHECK();
            }
            FusedArgs fa;
            fa.A = A; fa.lda = lda; fa.i = i; fa.m = m; fa.n = n; fa.mpad = mpad; fa.k = k; fa.nb = nb;
            fa.P = b.P; fa.ldp = b.ldp; fa.Q = b.Q; fa.ldq = b.ldq; fa.c = b.c; fa.rv = b.rv;
            fa.tmpN = b.tmpN; fa.ldt = lda;
            fa.dots1 = dots1_ready ? b.dots1p : b.dots1; fa.nparts1 = dots1_ready ? dots1_parts : 0;
            fa.dots2p = b.dots2p; fa.alpha = alpha; fa.T = pl.T; fa.NC = pl.NC; fa.Lc = pl.Lc;
            fa.prefetch = (g_pdl && dots1_ready && k > 0) ? 1 : 0;   // predecessor is finish_xf (not the panel GEMM)
            fa.trace = nullptr;
            static unsigned long long *d_trace = nullptr;
            const char *te = getenv("SVD_GPU_FZ_TRACE");
            const bool tracing = te && atoi(te) == i;
            if (tracing) {
                if (!d_trace) SVD_CUDA_CHECK(cudaMalloc(&d_trace, 8 * 256 * sizeof(unsigned long long)));
                SVD_CUDA_CHECK(cudaMemsetAsync(d_trace, 0, 8 * 256 * sizeof(unsigned long long), st));
                fa.trace = d_trace;
            }
            prof.begin(0, i, st);
            launch_fused(fa, pl, st);
            prof.end(st);
            if (tracing) {
                static unsigned long long h[8 * 256];
                SVD_CUDA_CHECK(cudaMemcpyAsync(h, d_trace, sizeof h, cudaMemcpyDeviceToHost, st));
                SVD_CUDA_CHECK(cudaStreamSynchronize(st));
                unsigned long long t0 = ~0ull;
                for (int z = 0; z < 8 * 256; ++z) if (h[z] && h[z] < t0) t0 = h[z];
                fprintf(stderr, "FZTRACE step %d CS %d Lc %d RPT %d T %d NC %d prefetch %d\n", i, pl.CS, pl.Lc, pl.RPT, pl.T, pl.NC, fa.prefetch);
                for (int nt = 0; nt < 256 && h[nt * 8]; ++nt) {
                    fprintf(stderr, "FZTRACE %3d", nt);
                    for (int z = 0; z < 8; ++z) fprintf(stderr, " %8lld", h[nt * 8 + z] ? (long long)(h[nt * 8 + z] - t0) : -1ll);
                    fprintf(stderr, "\n");
                }
            }
            // rows per CTA = 32 * RB: the smallest RB that keeps the row blocks within one wave
            // (one 1024-thread CTA per SM); the RB row groups of a CTA run one after the other
            {
                const int nColBlk = ceil_div(R, 1024);
                static const int old_rule = getenv("SVD_GPU_XF_OLD") ? 1 : 0;      // experiments
                const int rb = old_rule ? (Lb > 4096 ? 4 : 1)
                                        : (Lb <= 32 * (nsm - nColBlk)) ? 1 : (Lb <= 64 * (nsm - nColBlk)) ? 2 : 4;
                // SVD_GPU_XW: 0 = never the 128-rows-at-once kernel, 1 = where rb would exceed 1, 2 = always
                const bool wide = nb <= 32 && (xw_mode == 2 || (xw_mode == 1 && rb > 1));
                const int nGroups = ceil_div(Lb, XW_ROWS);
                const int nRowBlk = wide ? std::max(1, std::min(nGroups, nsm - nColBlk)) : ceil_div(Lb, 32 * rb);
                prof.begin(1, i, st);
                if (wide)
                    launch_finish_xw(nRowBlk + nColBlk, st, A, lda, i, m, n, k, nb, b.P, b.ldp, b.Q, b.ldq, b.c,
                                     (const double *)b.rv, (const double *)b.tmpN, lda, pl.NC,
                                     (const double *)b.dots2p, pl.NC, beta, nRowBlk, nGroups, b.dots1p);
                else if (rb == 1)
                    launch_finish_xf<1>(nRowBlk + nColBlk, st, A, lda, i, m, n, k, nb, b.P, b.ldp, b.Q, b.ldq, b.c,
                                        (const double *)b.rv, (const double *)b.tmpN, lda, pl.NC,
                                        (const double *)b.dots2p, pl.NC, beta, nRowBlk, b.dots1p);
                else if (rb == 2)
                    launch_finish_xf<2>(nRowBlk + nColBlk, st, A, lda, i, m, n, k, nb, b.P, b.ldp, b.Q, b.ldq, b.c,
                                        (const double *)b.rv, (const double *)b.tmpN, lda, pl.NC,
                                        (const double *)b.dots2p, pl.NC, beta, nRowBlk, b.dots1p);
                else
                    launch_finish_xf<4>(nRowBlk + nColBlk, st, A, lda, i, m, n, k, nb, b.P, b.ldp, b.Q, b.ldq, b.c,
                                        (const double *)b.rv, (const double *)b.tmpN, lda, pl.NC,
                                        (const double *)b.dots2p, pl.NC, beta, nRowBlk, b.dots1p);
                dots1_parts = nRowBlk;
                prof.end(st);
            }
            SVD_KERNEL_CHECK();
            dots1_ready = true;
        } else {
        dots1_ready = false;
        if (do_col) nsplitT = launch_gemvT(A, lda, i, m, n, mpad, b, nb, k, targetT, st);
        {
            int nColBlk = ceil_div(R, 32), nRowBlk = ceil_div(m - i, 256);
            finish_y_kernel<<<nColBlk + nRowBlk, 256, 0, st>>>(A, lda, i, m, n, k, nb, do_col, b.P, b.ldp,
                                                               b.Q, b.ldq, b.c, b.rv, b.tmpT, b.ldq,
                                                               nsplitT, b.dots1, alpha, nColBlk);
            SVD_KERNEL_CHECK();
        }
        if (tail && !do_row) break;             // tall/square tail: last column reflector only
        if (do_row) nsplitN = launch_gemvN(A, lda, i, m, n, mpad, b, nb, k, targetN, st);
        {
            int nRowBlk = Lb > 0 ? ceil_div(Lb, 32) : 1, nColBlk = ceil_div(R, 256);
            finish_x_kernel<1><<<nRowBlk + nColBlk, 256, 0, st>>>(A, lda, i, m, n, k, nb, do_row, b.P, b.ldp,
                                                                  b.Q, b.ldq, b.c, b.rv, b.tmpN, lda, nsplitN,
                                                                  b.dots2, beta, nRowBlk, nullptr, nullptr, nullptr);
            SVD_KERNEL_CHECK();
        }
        }   // split path
        ++k;
        if (k == nb && i + 1 < mn) {
            // trailing update: A[i+1:, i+1:] -= [V|X][i+1:, :] * [Y|U][i+1:, :]^T
            // (the panel layout keeps V at columns [0,nb) and X at [nb,2nb) => one K = 2nb GEMM)
            GemmArgs g = {};
            g.M = m - i - 1; g.N = n - i - 1; g.K = 2 * nb;
            g.A = b.P + (i + 1); g.lda = b.ldp; g.transA = 0;
            g.B = b.Q + (i + 1); g.ldb = b.ldq; g.transB = 1;
            g.C = A + (i + 1) + (long)(i + 1) * lda; g.ldc = lda;
            g.alpha = -1.0; g.beta = 1.0; g.batch = 1; g.splitk = 1;
            prof.begin(2, i, st);
            dgemm_dmma(g, st);
            prof.end(st);
            k = 0;
        }
    }
    if (hook && hook->fn) hook->fn(hook->user, mn, st);
    prof.report(mn, st);
}

// One streaming pass over the full matrix (step 0, empty panel), for roofline measurements.
void bidiag_pass_probe(int m, int n, const double *A, long lda, void *workspace, int which, cudaStream_t st)
{
    const int mpad = (int)round_up(m, 2);
    const BidiagBufs b = carve(workspace, m, n, lda);
    int targetT, targetN;
    sm_targets(targetT, targetN);
    if (which == 0) launch_gemvT(A, lda, 0, m, n, mpad, b, 32, 0, targetT, st);
    else if (which == 1) launch_gemvN(A, lda, 0, m, n, mpad, b, 32, 0, targetN, st);
    else {
        // the fused single-read pass of step 0 (writes the reflector into column 0 of A and panel
        // scratch in the workspace: hand it a scratch copy of the matrix)
        int dev = 0, nsm = 148;
        SVD_CUDA_CHECK(cudaGetDevice(&dev));
        SVD_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
        fused_set_attributes();
        const FusedPlan pl = plan_fused(0, m, n, mpad, nsm, 2, 1);
        if (!pl.ok) return;
        FusedArgs fa;
        fa.A = const_cast<double *>(A); fa.lda = lda; fa.i = 0; fa.m = m; fa.n = n; fa.mpad = mpad; fa.k = 0; fa.nb = 32;
        fa.P = b.P; fa.ldp = b.ldp; fa.Q = b.Q; fa.ldq = b.ldq; fa.c = b.c; fa.rv = b.rv;
        fa.tmpN = b.tmpN; fa.ldt = lda; fa.dots1 = b.dots1; fa.nparts1 = 0; fa.dots2p = b.dots2p;
        fa.alpha = b.dots2;                      // scratch: the probe has no alpha array
        fa.T = pl.T; fa.NC = pl.NC; fa.Lc = pl.Lc; fa.trace = nullptr; fa.prefetch = 0;
        launch_fused(fa, pl, st);
    }
}

} // namespace svdgpu
