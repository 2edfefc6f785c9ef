// twisted.cu — singular vectors of the bidiagonal by twisted factorizations on sm_100a.
//
// Replaces the host/OpenMP phase CalcRightSingularVectors / RighttoLeftSingularVectors
// (parallel-twisted.c:554-637, :530-551): SquareB (:290-316), CholFactorization (:320-362),
// TwistedFactorization (:431-528, with which_min_gamma :240-288), backsolve (:365-427),
// NormalizeVectors (:88-123), BidiagMatVec (:58-86).
//
// Same method — for every sigma factor B^T B - sigma^2 I from the top (forward) and from the
// bottom (backward), twist the two factorizations where |gamma| is smallest and solve
// N_k x = e_k outward from the twist — with these changes:
//   * the factorizations are the differential qd transforms dstqds / dpqds of
//     B^T B = L diag(a^2) L^T (no tridiagonal is formed; parallel-twisted.c:304-314 squares B
//     and runs plain LDL^T, which is what costs the reference its orthogonality);
//   * instead of the reference's extra solve with the twist pinned at m/2 (:392-424) one
//     Rayleigh-quotient correction  sigma^2 += gamma_k / ||z||^2  is applied and the vector is
//     recomputed (gamma_k and ||z||^2 are by-products);
//   * left vectors come from the same machinery on B B^T (the index-reversed bidiagonal);
//   * near pairs are orthogonalized to first order at the end (tw_pairfix).
//
// Mapping (round 2).  The two qd recurrences are serial in the position j with a division on the
// dependent chain, and independent across sigma: ONE LANE PER SINGULAR VALUE, forward and backward
// sweep in different CTAs, the right problem and the (independent) left problem of the last round in the
// same launch; the reciprocal is MUFU + two Newton steps so that the chain is ~100 cycles per position
// instead of the ~290 of an IEEE division.  What follows the twist index is NOT a serial solve any more:
//     z_j = prod_{i=j}^{k-1} r_i   (j < k),      z_j = prod_{i=k}^{j-1} u_i   (j > k)
// are running products of independent ratios, so tw_solve_scan cuts the positions into one segment per
// warp: a first sweep leaves per-segment products and sums of squares (=> ||z||^2 and the value entering
// every segment, all the Rayleigh-quotient round needs), a second sweep recomputes the segment from its
// entering value and writes the NORMALISED vector straight into V(:,i) / U(:,i) through a transposing
// shared-memory tile.  No z scratch, no separate normalise / transpose kernels; the phase moves
// ~16 n^2 doubles instead of ~30 n^2 and is bound by the two qd chains + HBM, not by a serial solve.
// The forward/backward pivots still pass through two [position][sigma] panels (the twist index needs
// both sweeps complete at every position); sigma is chunked so that they stay within 4 GiB.
// north_star suggests a thread block per sigma: with serial recurrences that leaves 31 of 32 lanes idle in
// the chain-bound sweeps, and the per-sigma arrays (2 x 128 KB at n = 16384) do not fit one SM's shared
// memory; the block-level parallelism north_star asks for is in tw_solve_scan (warps split the positions).
#include "common.cuh"
#include "twisted.cuh"
#include <cfloat>
#include <cstdlib>

namespace svdgpu {

struct TwProb { const double *q, *e, *ab; };

__global__ void tw_prep_kernel(int n, int mb, const double *__restrict__ a, const double *__restrict__ b,
                               double *__restrict__ q, double *__restrict__ e, double *__restrict__ ab,
                               double *__restrict__ pivmin)
{
    __shared__ double red[32];
    double mx = 0.0;
    for (int j = threadIdx.x; j < mb; j += blockDim.x) {
        double aj = (j < n) ? a[j] : 0.0;
        double bj = (j < mb - 1) ? b[j] : 0.0;
        q[j] = aj * aj; e[j] = bj * bj; ab[j] = aj * bj;
        mx = fmax(mx, fmax(aj * aj, bj * bj));
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, red[w]);
        *pivmin = fmax(mx, 1.0) * 1e-290;
    }
}

__global__ void tw_tau_init_kernel(int ns, const double *__restrict__ sigma, double *__restrict__ tau)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ns) tau[t] = sigma[t] * sigma[t];
}

__device__ __forceinline__ double guard_pivot(double d, double pivmin)
{
    return (fabs(d) < pivmin) ? -pivmin : d;
}
// 1/x for guarded pivots (|x| >= pivmin, finite): MUFU seed (20 bits) + two Newton steps, <= 2 ulp.
// The qd transforms are mixed-stable under such relative errors (they only perturb q, e by ulps).
__device__ __forceinline__ double tw_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// blockIdx.y == 0: dstqds  s_0 = -tau ; d+_j = q_j + s_j ; s_{j+1} = s_j e_j / d+_j - tau
// blockIdx.y == 1: dpqds   p_{m-1} = q_{m-1} - tau ; d-_{j+1} = e_j + p_{j+1} ; p_j = p_{j+1} q_j / d-_{j+1} - tau
// blockIdx.z: problem (0 right, 1 left = reversed bidiagonal); S, P of problem z start z*pstride further
__global__ void __launch_bounds__(32)
tw_qd_kernel(int mb, int ns, TwProb pr0, TwProb pr1, const double *__restrict__ tau, const double *__restrict__ pivmin_p,
             double *__restrict__ S, double *__restrict__ P, size_t pstride)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ns) return;
    const TwProb pr = blockIdx.z ? pr1 : pr0;
    const double *__restrict__ q = pr.q, *__restrict__ e = pr.e;
    S += blockIdx.z * pstride; P += blockIdx.z * pstride;
    const double tv = tau[t], pivmin = *pivmin_p;
    if (blockIdx.y == 0) {
        double s = -tv;
#pragma unroll 4
        for (int j = 0; j < mb; ++j) {
            S[(size_t)j * ns + t] = s;
            const double se = s * __ldg(e + j);
            const double dp = guard_pivot(__ldg(q + j) + s, pivmin);
            s = fma(se, tw_rcp(dp), -tv);
        }
    } else {
        double p = __ldg(q + mb - 1) - tv;
        P[(size_t)(mb - 1) * ns + t] = p;
#pragma unroll 4
        for (int j = mb - 2; j >= 0; --j) {
            const double pq = p * __ldg(q + j);
            const double dm = guard_pivot(__ldg(e + j) + p, pivmin);
            p = fma(pq, tw_rcp(dm), -tv);
            P[(size_t)j * ns + t] = p;
        }
    }
}

// gamma_j = s_j + p_j + tau; twist index = argmin |gamma_j|, later index on ties
// (which_min_gamma, parallel-twisted.c:277-284).  32 sigmas (lanes) x TWS_SL position slices (warps);
// every slice scans its positions with coalesced loads, the slices are merged in a fixed order.
constexpr int TWS_SL = 32;
__global__ void __launch_bounds__(32 * TWS_SL)
tw_select_kernel(int mb, int ns, const double *__restrict__ tau, const double *__restrict__ S,
                 const double *__restrict__ P, size_t pstride, int *__restrict__ kidx, double *__restrict__ gk)
{
    __shared__ double s_best[TWS_SL][32], s_g[TWS_SL][32];
    __shared__ int s_k[TWS_SL][32];
    const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;
    S += blockIdx.y * pstride; P += blockIdx.y * pstride;
    kidx += (size_t)blockIdx.y * ns; gk += (size_t)blockIdx.y * ns;
    double best = DBL_MAX, bestg = 0.0;
    int bk = 0;
    if (t < ns) {
        const double tv = tau[t];
        const int chunk = (mb + TWS_SL - 1) / TWS_SL;
        const int j0 = sl * chunk, j1 = min(mb, j0 + chunk);
#pragma unroll 8
        for (int j = j0; j < j1; ++j) {
            double g = S[(size_t)j * ns + t] + P[(size_t)j * ns + t] + tv;
            double ag = fabs(g);
            if (ag <= best) { best = ag; bestg = g; bk = j; }
        }
    }
    s_best[sl][lane] = best; s_g[sl][lane] = bestg; s_k[sl][lane] = bk;
    __syncthreads();
    if (sl == 0 && t < ns) {
        // slices cover increasing position ranges: "<=" keeps the later index on ties
        for (int z = 1; z < TWS_SL; ++z)
            if (s_best[z][lane] <= best) { best = s_best[z][lane]; bestg = s_g[z][lane]; bk = s_k[z][lane]; }
        kidx[t] = bk;
        gk[t] = bestg;
    }
}

// ---- the outward solve as a segmented product scan ----------------------------------------------
// z_k = 1;  p < k: z_p = r_p z_{p+1},  r_p = -ab_p / (q_p + s_p);   p > k: z_p = u_{p-1} z_{p-1},
// u_{p-1} = -ab_{p-1} / (e_{p-1} + p_p)   (TwistedFactorization, parallel-twisted.c:495-521).
// CTA = 32 sigmas (lanes) x TSC_W warps; warp w owns output positions [w*Lseg, (w+1)*Lseg), Lseg a
// multiple of 32.  Sweep 1: per segment the product of its ratios on either side of the twist (what maps
// the value entering the segment to the value leaving it) and the sum of squares relative to the entering
// value.  Then every warp knows its entering values, and ||z||^2 is assembled in a fixed order.
// WRITE: sweep 2 recomputes the segment from the entering value, scaled by sgn / ||z||, and writes
// X[t*ldx + p] (or, reversed, X[t*ldx + mb-1-p] for the left vectors) through a transposing tile.
constexpr int TSC_W = 8;
constexpr int TSC_UB = 16;               // rows whose loads are in flight per batch
constexpr int TSC_TILE_BYTES = TSC_W * 32 * 33 * (int)sizeof(double);

struct TwSolveArgs {
    int mb, ns;
    TwProb pr;
    const double *pivmin;
    const int *kidx;
    const double *S, *P;                 // [position][sigma]
    double *nrm2;                        // [ns]   ||z||^2
    const double *tau;
    double *sigma_out;                   // optional: sqrt(tau)
    const double *sgn;                   // optional: +-1 per sigma
    double *X; long ldx; int rev;        // output (WRITE only)
};

template <bool WRITE>
__global__ void __launch_bounds__(32 * TSC_W)
tw_solve_scan_kernel(const TwSolveArgs a)
{
    __shared__ double st_dn[TSC_W][32], st_up[TSC_W][32], st_sqd[TSC_W][32], st_squ[TSC_W][32];
    __shared__ double s_scale[32];
    __shared__ int s_k[32];
    extern __shared__ __align__(16) double tw_tile[];       // WRITE: [TSC_W][32][33]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int t0 = blockIdx.x * 32, tt = t0 + lane;
    const bool valid = tt < a.ns;
    const int t = valid ? tt : a.ns - 1;                    // idle lanes shadow a valid column, never store
    const int mb = a.mb, ns = a.ns;
    const int k = a.kidx[t];
    const double pivmin = *a.pivmin;
    const double *__restrict__ q = a.pr.q, *__restrict__ e = a.pr.e, *__restrict__ ab = a.pr.ab;
    const int Lseg = ((mb + TSC_W - 1) / TSC_W + 31) / 32 * 32;
    const int sa = min(mb, w * Lseg), sb = min(mb, sa + Lseg);   // this warp's output positions [sa, sb)
    int kmax = k, kmin = k;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    }
    if (w == 0) s_k[lane] = k;

    // ---- sweep 1 --------------------------------------------------------------------------------
    double dn = 1.0, sqd = 0.0, up = 1.0, squ = 0.0;
    {   // down side: positions p in [sa, sb) with p < k, from the top of the segment downwards
        const int hi = min(sb, kmax);                        // warp-uniform: nobody is active at p >= kmax
        for (int pb = hi - 1; pb >= sa; pb -= TSC_UB) {
            double sv[TSC_UB];
#pragma unroll
            for (int u = 0; u < TSC_UB; ++u) { const int p = pb - u; sv[u] = (p >= sa) ? a.S[(size_t)p * ns + t] : 0.0; }
#pragma unroll
            for (int u = 0; u < TSC_UB; ++u) {
                const int p = pb - u;
                if (p >= sa && p < k) {
                    const double dp = guard_pivot(__ldg(q + p) + sv[u], pivmin);
                    dn *= -(__ldg(ab + p) * tw_rcp(dp));
                    sqd = fma(dn, dn, sqd);
                }
            }
        }
    }
    {   // up side: positions p in [sa, sb) with p > k, from the bottom of the segment upwards
        const int lo = max(sa, kmin + 1);
        for (int pb = lo; pb < sb; pb += TSC_UB) {
            double pv[TSC_UB];
#pragma unroll
            for (int u = 0; u < TSC_UB; ++u) { const int p = pb + u; pv[u] = (p < sb) ? a.P[(size_t)p * ns + t] : 0.0; }
#pragma unroll
            for (int u = 0; u < TSC_UB; ++u) {
                const int p = pb + u;
                if (p < sb && p > k) {
                    const double dm = guard_pivot(__ldg(e + p - 1) + pv[u], pivmin);
                    up *= -(__ldg(ab + p - 1) * tw_rcp(dm));
                    squ = fma(up, up, squ);
                }
            }
        }
    }
    st_dn[w][lane] = dn; st_sqd[w][lane] = sqd; st_up[w][lane] = up; st_squ[w][lane] = squ;
    __syncthreads();
    // value entering this segment from above (down side) / from below (up side), and ||z||^2
    double cdn = 1.0, cup = 1.0;
    for (int v = TSC_W - 1; v > w; --v) cdn *= st_dn[v][lane];
    for (int v = 0; v < w; ++v) cup *= st_up[v][lane];
    if (w == 0) {
        double nn = 1.0, c = 1.0;
        for (int v = TSC_W - 1; v >= 0; --v) { nn = fma(c * c, st_sqd[v][lane], nn); c *= st_dn[v][lane]; }
        c = 1.0;
        for (int v = 0; v < TSC_W; ++v) { nn = fma(c * c, st_squ[v][lane], nn); c *= st_up[v][lane]; }
        if (valid) {
            a.nrm2[t] = nn;
            if (WRITE && a.sigma_out) a.sigma_out[t] = sqrt(a.tau[t]);
        }
        s_scale[lane] = rsqrt(nn) * ((WRITE && a.sgn) ? a.sgn[t] : 1.0);
    }
    if (!WRITE) return;
    __syncthreads();
    // ---- sweep 2: the segment again, from its entering value, into the output ---------------------
    const double scale = s_scale[lane];
    double *tl = tw_tile + (size_t)w * 32 * 33;              // tl[i*33 + r]: sigma t0+r at position jb+i
    auto flush_tile = [&](int jb, bool down) {
        __syncwarp();
        const int pos = jb + lane;
        for (int r = 0; r < 32; ++r) {
            const int kr = s_k[r];
            const bool mine = down ? (pos < kr) : (pos >= kr);
            if (t0 + r < ns && pos >= sa && pos < sb && mine) {
                const long o = a.rev ? (long)(mb - 1 - pos) : (long)pos;
                a.X[(size_t)(t0 + r) * a.ldx + o] = tl[lane * 33 + r];
            }
        }
        __syncwarp();
    };
    {   // down side, blocks of 32 positions from the top
        const int hi = min(sb, kmax);
        double z = cdn * scale;
        for (int jb = (hi - 1) & ~31; hi > sa && jb >= sa; jb -= 32) {
            const int ptop = min(jb + 32, hi) - 1;
            for (int pb = ptop; pb >= jb; pb -= TSC_UB) {
                double sv[TSC_UB];
#pragma unroll
                for (int u = 0; u < TSC_UB; ++u) { const int p = pb - u; sv[u] = (p >= jb) ? a.S[(size_t)p * ns + t] : 0.0; }
#pragma unroll
                for (int u = 0; u < TSC_UB; ++u) {
                    const int p = pb - u;
                    if (p >= jb) {
                        if (p < k) {
                            const double dp = guard_pivot(__ldg(q + p) + sv[u], pivmin);
                            z *= -(__ldg(ab + p) * tw_rcp(dp));
                        }
                        tl[(p - jb) * 33 + lane] = z;
                    }
                }
            }
            flush_tile(jb, true);
        }
    }
    {   // up side (and the twist position itself, z_k = 1), blocks of 32 positions from the bottom
        // (an empty trailing segment, sa == sb == mb, must not touch the block that holds position mb-1: it belongs
        //  to another warp — found by compute-sanitizer timing, see tests/test_gpu_round2.py::test_twisted_every_entry_written)
        const int lo = max(sa, kmin);
        double z = cup * scale;
        for (int jb = lo & ~31; sa < sb && jb < sb; jb += 32) {
            const int pbot = max(jb, lo), pend = min(jb + 32, sb);
            for (int pb = pbot; pb < pend; pb += TSC_UB) {
                double pv[TSC_UB];
#pragma unroll
                for (int u = 0; u < TSC_UB; ++u) { const int p = pb + u; pv[u] = (p < pend) ? a.P[(size_t)p * ns + t] : 0.0; }
#pragma unroll
                for (int u = 0; u < TSC_UB; ++u) {
                    const int p = pb + u;
                    if (p < pend) {
                        if (p > k) {
                            const double dm = guard_pivot(__ldg(e + p - 1) + pv[u], pivmin);
                            z *= -(__ldg(ab + p - 1) * tw_rcp(dm));
                        }
                        tl[(p - jb) * 33 + lane] = z;
                    }
                }
            }
            flush_tile(jb, false);
        }
    }
}

// Rayleigh-quotient correction tau += gamma_k/||z||^2, accepted only while sigma stays
// between the midpoints to its neighbours.
__global__ void tw_rqi_kernel(int ns, int i0, int ntot, const double *__restrict__ sigma_all,
                              const double *__restrict__ gk, const double *__restrict__ nrm2,
                              double *__restrict__ tau)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ns) return;
    double tn = tau[t] + gk[t] / nrm2[t];
    int gi = i0 + t;
    double s = sigma_all[gi];
    double lo = (gi > 0) ? 0.5 * (sigma_all[gi - 1] + s) : 0.0;
    double hi = (gi + 1 < ntot) ? 0.5 * (sigma_all[gi + 1] + s) : DBL_MAX;
    if (tn > 0.0) {
        double sn = sqrt(tn);
        if (sn >= lo && sn <= hi) tau[t] = tn;
    }
}

// y = B x / sigma for the N x (N+1) bidiagonal of a wide input (BidiagMatVec :75-84,
// RighttoLeftSingularVectors :545-549): Y[t*ldy + j] = (a_j x_j + b_j x_{j+1}) / sigma_t
__global__ void __launch_bounds__(256)
tw_y_from_x_kernel(int n, int mb, const double *__restrict__ a, const double *__restrict__ b,
                   const double *__restrict__ tau, const double *__restrict__ X, long ldx, double *__restrict__ Y, long ldy)
{
    const int t = blockIdx.y;
    const double sg = sqrt(tau[t]);
    const double inv = (sg > 0.0) ? 1.0 / sg : 0.0;
    const double *x = X + (size_t)t * ldx;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < n; j += gridDim.x * 256) {
        const double bj = (j < mb - 1) ? b[j] : 0.0;
        const double x1 = (j + 1 < mb) ? x[j + 1] : 0.0;
        Y[(size_t)t * ldy + j] = (a[j] * x[j] + bj * x1) * inv;
    }
}

// ---- left vectors by their own twisted factorization ------------------------------------------
// For a square upper bidiagonal B, P B^T P (P = index reversal) is again upper bidiagonal, with
// diagonal a'_j = a_{n-1-j} and super-diagonal b'_j = b_{n-2-j}, and its right singular vectors are
// the reversed left singular vectors of B.  Running the same qd kernels on (a', b') gives y_i with
// the accuracy of x_i, whereas y = B x / sigma (parallel-twisted.c:545-549) loses orthogonality
// like eps * sigma_max / sigma_i.  The pair (x_i, y_i) is oriented by the sign of (B x_i) at the
// twist position of y_i (its dominant component).
__global__ void tw_prep_rev_kernel(int n, const double *__restrict__ a, const double *__restrict__ b,
                                   double *__restrict__ q, double *__restrict__ e, double *__restrict__ ab)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const double aj = a[n - 1 - j];
        const double bj = (j < n - 1) ? b[n - 2 - j] : 0.0;
        q[j] = aj * aj; e[j] = bj * bj; ab[j] = aj * bj;
    }
}

__global__ void tw_left_sign_kernel(int n, int ns, const double *__restrict__ a, const double *__restrict__ b,
                                    const int *__restrict__ kidx, const double *__restrict__ X, long ldx,
                                    double *__restrict__ sgn)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ns) return;
    const int r = n - 1 - kidx[t];                         // position of the dominant component of y_t
    const double *x = X + (size_t)t * ldx;
    double bx = a[r] * x[r];
    if (r < n - 1) bx += b[r] * x[r + 1];
    sgn[t] = (bx < 0.0) ? -1.0 : 1.0;
}

// ---- near-pair orthogonalization -------------------------------------------------------------------
// Vectors of a twisted factorization are computed independently, so two neighbours with a small relative
// gap come out orthogonal only to ~eps / relgap: at n = 16384 (relative gaps down to 3e-6) the adjacent
// pairs alone are 99 % of ||X^T X - I||_F (bench/orth_probe.py, profiles/r02_orth_probe_16384_pairfix.log).
// The reference has no counterpart (its vectors are orthogonal to 1e-2 .. 1e-5).  For the
// pair (t, t+k), g = x_t . x_{t+k} is tiny, and the symmetric first-order (Loewdin) correction
//     x_t <- x_t - (g/2) x_{t+k},   x_{t+k} <- x_{t+k} - (g/2) x_t
// removes it at O(g^2), changes the norms at O(g^2) and the residual by ~sigma * g * relgap.  One CTA per
// pair; the pairs of one launch are disjoint (phase = which residue of t mod 2k starts a pair), g is taken
// from the current vectors.  1024-thread CTAs keep at most two pairs per SM in flight (~76 MB of vectors at
// n = 16384), so the update's second read of the pair comes from L2, not from HBM.  Pairs that straddle two
// ranks' blocks are left alone.
constexpr int PF_T = 1024;
__global__ void __launch_bounds__(PF_T)
tw_pairfix_kernel(double *__restrict__ Z, long ld, int len, int ns, int k, int phase)
{
    __shared__ double red[PF_T / 32];
    __shared__ double s_g;
    // pairs start at t with (t / k) % 2 == phase: t in [0,k) pairs with [k,2k) ... disjoint within a launch
    const int p = blockIdx.x;
    const int t = (p / k) * 2 * k + phase * k + (p % k);
    if (t + k >= ns) return;
    double *x = Z + (size_t)t * ld, *y = Z + (size_t)(t + k) * ld;
    double acc0 = 0.0, acc1 = 0.0;
    int j = threadIdx.x;
    for (; j + PF_T < len; j += 2 * PF_T) {
        acc0 = fma(x[j], y[j], acc0);
        acc1 = fma(x[j + PF_T], y[j + PF_T], acc1);
    }
    if (j < len) acc0 = fma(x[j], y[j], acc0);
    double acc = warp_sum(acc0 + acc1);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double g = 0.0;
        for (int w = 0; w < PF_T / 32; ++w) g += red[w];
        s_g = g;
    }
    __syncthreads();
    const double h = 0.5 * s_g;
    // |g| <= 2e-14 is already at the level the independent vectors reach (its share of ||X^T X - I||_F stays below
    // 2e-14 sqrt(4 n)); |g| > 2e-6 is not a "nearly orthogonal" pair: leave both alone (and skip the write pass)
    if (!(fabs(h) > 1e-14) || fabs(h) > 1e-6) return;
    for (j = threadIdx.x; j < len; j += PF_T) {
        const double a = x[j], b = y[j];
        x[j] = fma(-h, b, a);
        y[j] = fma(-h, a, b);
    }
}

static void tw_pairfix(double *Z, long ld, int len, int ns, cudaStream_t st)
{
    static const int kmax = getenv("SVD_GPU_PAIRFIX") ? atoi(getenv("SVD_GPU_PAIRFIX")) : 2;
    for (int k = 1; k <= kmax && k < ns; ++k)
        for (int phase = 0; phase < 2; ++phase) {
            // number of pair slots: groups of 2k vectors, k pairs each (the kernel drops pairs beyond ns)
            const int groups = (ns + 2 * k - 1) / (2 * k);
            tw_pairfix_kernel<<<groups * k, PF_T, 0, st>>>(Z, ld, len, ns, k, phase);
            SVD_KERNEL_CHECK();
        }
}

static int tw_chunk(int mb, int ns)
{
    const size_t budget = (size_t)4 << 30;                 // bytes for the four scratch panels (S, P of two problems)
    long c = (long)(budget / (4 * sizeof(double) * (size_t)mb));
    c = c / 32 * 32;
    if (c < 32) c = 32;
    if (c > ns) c = ns;
    return (int)c;
}

size_t twisted_workspace_bytes(int n, int mb, int ns)
{
    (void)n;
    int c = tw_chunk(mb, ns);
    size_t d = 0;
    d += 6 * (size_t)mb + 8;            // q, e, ab (+ the reversed set for the left vectors), pivmin
    d += 4 * (size_t)mb * c;            // S, P of the right and of the left problem
    d += 6 * (size_t)c + 16;            // tau, gk[2], nrm2[2], sgn
    return d * sizeof(double) + 2 * (size_t)c * sizeof(int) + 4096;
}

void twisted_vectors_device(int n, int mb, const double *a, const double *b, const double *sigma_all,
                            int ntot, int i0, int ns, double *X, long ldx, double *Y, long ldy,
                            double *sigma_out, int rqi_steps, void *workspace, cudaStream_t st)
{
    if (ns <= 0) return;
    const int cmax = tw_chunk(mb, ns);
    const size_t pstride = (size_t)mb * cmax;
    double *w = (double *)workspace;
    double *q = w;       w += mb;
    double *e = w;       w += mb;
    double *ab = w;      w += mb;
    double *q2 = w;      w += mb;
    double *e2 = w;      w += mb;
    double *ab2 = w;     w += mb;
    double *pivmin = w;  w += 8;
    double *S = w;       w += 2 * pstride;
    double *P = w;       w += 2 * pstride;
    double *tau = w;     w += cmax;
    double *gk = w;      w += 2 * (size_t)cmax;
    double *nrm2 = w;    w += 2 * (size_t)cmax;
    double *sgn = w;     w += cmax + 8;
    int *kidx = (int *)w;

    static DeviceOnce once;
    if (first_on_device(once))
        SVD_CUDA_CHECK(cudaFuncSetAttribute(tw_solve_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TSC_TILE_BYTES));

    tw_prep_kernel<<<1, 1024, 0, st>>>(n, mb, a, b, q, e, ab, pivmin);
    SVD_KERNEL_CHECK();
    // left vectors from their own twisted factorization when B is square (SVD_GPU_LEFT=bx restores y = Bx/sigma)
    const char *lenv = getenv("SVD_GPU_LEFT");
    const bool left_by_twist = (Y != nullptr) && (mb == n) && (n > 1) && !(lenv && lenv[0] == 'b');
    if (left_by_twist) {
        tw_prep_rev_kernel<<<ceil_div(n, 256), 256, 0, st>>>(n, a, b, q2, e2, ab2);
        SVD_KERNEL_CHECK();
    }
    const TwProb prR = {q, e, ab}, prL = {q2, e2, ab2};
    for (int c0 = 0; c0 < ns; c0 += cmax) {
        const int c = (ns - c0 < cmax) ? ns - c0 : cmax;
        const int gi0 = i0 + c0;
        const int g32 = ceil_div(c, 32);
        // within a chunk the panels are laid out for c columns: problem z starts mb*c further
        const size_t ps = (size_t)mb * c;
        tw_tau_init_kernel<<<ceil_div(c, 256), 256, 0, st>>>(c, sigma_all + gi0, tau);
        SVD_KERNEL_CHECK();
        TwSolveArgs sa = {};
        sa.mb = mb; sa.ns = c; sa.pr = prR; sa.pivmin = pivmin; sa.kidx = kidx; sa.S = S; sa.P = P; sa.nrm2 = nrm2;
        sa.tau = tau;
        // Rayleigh-quotient rounds (right problem): qd sweeps, twist index, ||z||^2 only
        for (int sweep = 0; sweep < rqi_steps; ++sweep) {
            tw_qd_kernel<<<dim3(g32, 2, 1), 32, 0, st>>>(mb, c, prR, prL, tau, pivmin, S, P, ps);
            SVD_KERNEL_CHECK();
            tw_select_kernel<<<dim3(g32, 1), 32 * TWS_SL, 0, st>>>(mb, c, tau, S, P, ps, kidx, gk);
            SVD_KERNEL_CHECK();
            tw_solve_scan_kernel<false><<<g32, 32 * TSC_W, 0, st>>>(sa);
            SVD_KERNEL_CHECK();
            tw_rqi_kernel<<<ceil_div(c, 256), 256, 0, st>>>(c, gi0, ntot, sigma_all, gk, nrm2, tau);
            SVD_KERNEL_CHECK();
        }
        // last round with the polished sigma^2: the right and the left problem are independent and their
        // latency-bound qd sweeps share one launch
        const int nprob = left_by_twist ? 2 : 1;
        tw_qd_kernel<<<dim3(g32, 2, nprob), 32, 0, st>>>(mb, c, prR, prL, tau, pivmin, S, P, ps);
        SVD_KERNEL_CHECK();
        tw_select_kernel<<<dim3(g32, nprob), 32 * TWS_SL, 0, st>>>(mb, c, tau, S, P, ps, kidx, gk);
        SVD_KERNEL_CHECK();
        double *Xc = X + (size_t)c0 * ldx;
        double *Yc = Y ? Y + (size_t)c0 * ldy : nullptr;
        sa.X = Xc; sa.ldx = ldx; sa.rev = 0; sa.sgn = nullptr; sa.sigma_out = sigma_out ? sigma_out + c0 : nullptr;
        tw_solve_scan_kernel<true><<<g32, 32 * TSC_W, TSC_TILE_BYTES, st>>>(sa);
        SVD_KERNEL_CHECK();
        if (left_by_twist) {
            tw_left_sign_kernel<<<ceil_div(c, 256), 256, 0, st>>>(n, c, a, b, kidx + c, Xc, ldx, sgn);
            SVD_KERNEL_CHECK();
            TwSolveArgs sl = sa;
            sl.pr = prL; sl.kidx = kidx + c; sl.S = S + ps; sl.P = P + ps; sl.nrm2 = nrm2 + c;
            sl.X = Yc; sl.ldx = ldy; sl.rev = 1; sl.sgn = sgn; sl.sigma_out = nullptr;
            tw_solve_scan_kernel<true><<<g32, 32 * TSC_W, TSC_TILE_BYTES, st>>>(sl);
            SVD_KERNEL_CHECK();
        } else if (Yc != nullptr) {
            tw_y_from_x_kernel<<<dim3(ceil_div(n, 1024), c), 256, 0, st>>>(n, mb, a, b, tau, Xc, ldx, Yc, ldy);
            SVD_KERNEL_CHECK();
        }
    }
    tw_pairfix(X, ldx, mb, ns, st);
    if (Y != nullptr) tw_pairfix(Y, ldy, n, ns, st);
}

} // namespace svdgpu
