// twisted.cu — singular vectors of the bidiagonal by twisted factorizations on sm_100a.
//
// Replaces the host/OpenMP phase CalcRightSingularVectors / RighttoLeftSingularVectors
// (parallel-twisted.c:554-637, :530-551): SquareB (:290-316), CholFactorization (:320-362),
// TwistedFactorization (:431-528, with which_min_gamma :240-288), backsolve (:365-427),
// NormalizeVectors (:88-123), BidiagMatVec (:58-86).
//
// Same method — for every sigma factor B^T B - sigma^2 I from the top (forward) and from the
// bottom (backward), twist the two factorizations where |gamma| is smallest and solve
// N_k x = e_k outward from the twist — with three changes:
//   * the factorizations are the differential qd transforms dstqds / dpqds of
//     B^T B = L diag(a^2) L^T (no tridiagonal is formed; parallel-twisted.c:304-314 squares B
//     and runs plain LDL^T, which is what costs the reference its orthogonality);
//   * instead of the reference's extra solve with the twist pinned at m/2 (:392-424) one
//     Rayleigh-quotient correction  sigma^2 += gamma_k / ||z||^2  is applied and the vector is
//     recomputed (gamma_k and ||z||^2 are by-products);
//   * none of the reference's six n x m work arrays survive; two scratch panels in a
//     [position][sigma] layout are reused in place.
//
// Mapping: ONE LANE PER SINGULAR VALUE.  The recurrences are serial in the position index j
// (one divide per step) and independent across sigma, so a warp advances 32 sigmas in
// lock-step with fully coalesced scratch traffic; forward and backward sweeps (and later the
// two halves of the outward solve) run concurrently in different CTAs.  A CTA-per-sigma
// mapping would issue the same recurrences with 1 of 32 lanes active.
#include "common.cuh"
#include "twisted.cuh"
#include <cfloat>
#include <cstdlib>

namespace svdgpu {

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

// blockIdx.y == 0: dstqds  s_0 = -tau ; d+_j = q_j + s_j ; s_{j+1} = s_j e_j / d+_j - tau
// blockIdx.y == 1: dpqds   p_{m-1} = q_{m-1} - tau ; d-_{j+1} = e_j + p_{j+1} ; p_j = p_{j+1} q_j / d-_{j+1} - tau
__global__ void __launch_bounds__(32)
tw_qd_kernel(int mb, int ns, const double *__restrict__ q, const double *__restrict__ e,
             const double *__restrict__ tau, const double *__restrict__ pivmin_p,
             double *__restrict__ S, double *__restrict__ P)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ns) return;
    const double tv = tau[t], pivmin = *pivmin_p;
    if (blockIdx.y == 0) {
        double s = -tv;
#pragma unroll 4
        for (int j = 0; j < mb; ++j) {
            S[(size_t)j * ns + t] = s;
            double dp = guard_pivot(__ldg(q + j) + s, pivmin);
            s = s * (__ldg(e + j) / dp) - tv;
        }
    } else {
        double p = __ldg(q + mb - 1) - tv;
        P[(size_t)(mb - 1) * ns + t] = p;
#pragma unroll 4
        for (int j = mb - 2; j >= 0; --j) {
            double dm = guard_pivot(__ldg(e + j) + p, pivmin);
            p = p * (__ldg(q + j) / dm) - tv;
            P[(size_t)j * ns + t] = p;
        }
    }
}

// gamma_j = s_j + p_j + tau; twist index = argmin |gamma_j|, later index on ties
// (which_min_gamma, parallel-twisted.c:277-284).  32 sigmas (lanes) x 8 position slices (warps);
// every slice scans its positions with coalesced loads, the slices are merged in a fixed order.
// TWS_SL position slices per CTA: 8 by default; SVD_GPU_TW_SL=32 is an experiment (more loads in flight)
template <int TWS_SL>
__global__ void __launch_bounds__(32 * TWS_SL)
tw_select_kernel(int mb, int ns, const double *__restrict__ tau, const double *__restrict__ S,
                 const double *__restrict__ P, int *__restrict__ kidx, double *__restrict__ gk)
{
    __shared__ double s_best[TWS_SL][32], s_g[TWS_SL][32];
    __shared__ int s_k[TWS_SL][32];
    const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;
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

// z_k = 1; j < k: z_j = -(ab_j / d+_j) z_{j+1};  j >= k: z_{j+1} = -(ab_j / d-_{j+1}) z_j
// (TwistedFactorization, parallel-twisted.c:495-521).  z overwrites S in place.
// All lanes walk the SAME rows (coalesced, loads batched 8 rows ahead of the serial recurrence);
// a lane simply stays idle (z = 1) until the walk reaches its own twist index.
// UB = rows whose loads are in flight ahead of the recurrence (8 by default; SVD_GPU_TW_UB=16/32 is an experiment:
// 128 warps x 8 x 256 B in flight cannot cover the HBM latency at n = 4096)
template <int UB>
__global__ void __launch_bounds__(32)
tw_solve_kernel(int mb, int ns, const double *__restrict__ q, const double *__restrict__ e,
                const double *__restrict__ ab, const double *__restrict__ pivmin_p,
                const int *__restrict__ kidx, double *__restrict__ S, const double *__restrict__ P,
                double *__restrict__ nrm2)
{
    const int tt = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = tt < ns;
    const int t = valid ? tt : ns - 1;                      // idle lanes shadow a valid column, never store
    const int k = kidx[t];
    const double pivmin = *pivmin_p;
    double z = 1.0, acc = 0.0;
    if (blockIdx.y == 0) {
        // downward: rows kmax-1 .. 0 (warp-uniform start), a lane is active where j < k
        const int kmax = __reduce_max_sync(0xffffffffu, k);
        for (int jb = min(mb - 2, kmax - 1); jb >= 0; jb -= UB) {
            double sv[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) { const int j = jb - u; sv[u] = (j >= 0) ? S[(size_t)j * ns + t] : 0.0; }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int j = jb - u;
                if (j >= 0 && j < k) {
                    double dp = guard_pivot(__ldg(q + j) + sv[u], pivmin);
                    z = -(__ldg(ab + j) / dp) * z;
                    if (valid) S[(size_t)j * ns + t] = z;
                    acc += z * z;
                }
            }
        }
        if (valid) { S[(size_t)k * ns + t] = 1.0; nrm2[t] = acc + 1.0; }
    } else {
        // upward: rows kmin .. mb-2 write z_{j+1}, a lane is active where j >= k
        const int kmin = __reduce_min_sync(0xffffffffu, k);
        for (int jb = kmin; jb < mb - 1; jb += UB) {
            double pv[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) { const int j = jb + u; pv[u] = (j < mb - 1) ? P[(size_t)(j + 1) * ns + t] : 0.0; }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int j = jb + u;
                if (j < mb - 1 && j >= k) {
                    double dm = guard_pivot(__ldg(e + j) + pv[u], pivmin);
                    z = -(__ldg(ab + j) / dm) * z;
                    if (valid) S[(size_t)(j + 1) * ns + t] = z;
                    acc += z * z;
                }
            }
        }
        if (valid) nrm2[ns + t] = acc;
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
    double nn = nrm2[t] + nrm2[ns + t];
    double tn = tau[t] + gk[t] / nn;
    int gi = i0 + t;
    double s = sigma_all[gi];
    double lo = (gi > 0) ? 0.5 * (sigma_all[gi - 1] + s) : 0.0;
    double hi = (gi + 1 < ntot) ? 0.5 * (sigma_all[gi + 1] + s) : DBL_MAX;
    if (tn > 0.0) {
        double sn = sqrt(tn);
        if (sn >= lo && sn <= hi) tau[t] = tn;
    }
}

// normalise, transpose to the reference's layout X[i*mb + j] (NormalizeVectors :106-120) and
// form y = B x / sigma, Y[i*n + j] (BidiagMatVec :75-84, RighttoLeftSingularVectors :545-549)
__global__ void __launch_bounds__(256)
tw_finalize_kernel(int n, int mb, int ns, const double *__restrict__ a, const double *__restrict__ b,
                   const double *__restrict__ tau, const double *__restrict__ Z,
                   const double *__restrict__ nrm2, double *__restrict__ X, long ldx,
                   double *__restrict__ Y, long ldy, double *__restrict__ sigma_out)
{
    __shared__ double tile[33][33];
    const int t0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    // load 33 positions (one halo row) x 32 sigmas, scaled
    for (int r = ty; r < 33; r += 8) {
        int j = j0 + r, t = t0 + tx;
        double v = 0.0;
        if (j < mb && t < ns) v = Z[(size_t)j * ns + t] * rsqrt(nrm2[t] + nrm2[ns + t]);
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int t = t0 + r, j = j0 + tx;
        if (t >= ns) continue;
        double xj = tile[tx][r], xj1 = tile[tx + 1][r];
        if (j < mb) X[(size_t)t * ldx + j] = xj;
        if (Y != nullptr && j < n) {
            double sg = sqrt(tau[t]);
            double bj = (j < mb - 1) ? b[j] : 0.0;
            double y = a[j] * xj + bj * xj1;
            Y[(size_t)t * ldy + j] = (sg > 0.0) ? y / sg : 0.0;
        }
    }
    if (blockIdx.y == 0 && threadIdx.x < 32 && sigma_out != nullptr) {
        int t = t0 + threadIdx.x;
        if (t < ns) sigma_out[t] = sqrt(tau[t]);
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

__global__ void __launch_bounds__(256)
tw_left_finalize_kernel(int n, int ns, const double *__restrict__ Z, const double *__restrict__ nrm2,
                        const double *__restrict__ sgn, double *__restrict__ Y, long ldy)
{
    __shared__ double tile[32][33];
    const int t0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int j = j0 + r, t = t0 + tx;
        double v = 0.0;
        if (j < n && t < ns) v = Z[(size_t)j * ns + t] * rsqrt(nrm2[t] + nrm2[ns + t]) * sgn[t];
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int t = t0 + r, j = j0 + tx;
        if (t < ns && j < n) Y[(size_t)t * ldy + (n - 1 - j)] = tile[tx][r];
    }
}

// ---- near-pair orthogonalization -------------------------------------------------------------------
// Vectors of a twisted factorization are computed independently, so two neighbours with a small relative
// gap come out orthogonal only to ~eps / relgap: at n = 16384 (relative gaps down to 3e-6) the adjacent
// pairs alone are 99 % of ||X^T X - I||_F (bench/orth_probe.py, profiles/r02_orth_probe_16384.log).  The
// reference has no counterpart (its vectors are orthogonal to 1e-2 .. 1e-5).  For the pair
// (t, t+k), g = x_t . x_{t+k} is tiny, and the symmetric first-order (Loewdin) correction
//     x_t <- x_t - (g/2) x_{t+k},   x_{t+k} <- x_{t+k} - (g/2) x_t
// removes it at O(g^2), changes the norms at O(g^2) and the residual by ~sigma * g * relgap.  One CTA per
// pair; the pairs of one launch are disjoint (phase = which residue of t mod 2k starts a pair), g is taken
// from the current vectors.  Pairs that straddle two ranks' blocks are left alone.
__global__ void __launch_bounds__(256)
tw_pairfix_kernel(double *__restrict__ Z, long ld, int len, int ns, int k, int phase)
{
    __shared__ double red[8];
    __shared__ double s_g;
    // pairs start at t with (t / k) % 2 == phase: t in [0,k) pairs with [k,2k) ... disjoint within a launch
    const int p = blockIdx.x;
    const int t = (p / k) * 2 * k + phase * k + (p % k);
    if (t + k >= ns) return;
    double *x = Z + (size_t)t * ld, *y = Z + (size_t)(t + k) * ld;
    double acc = 0.0;
    for (int j = threadIdx.x; j < len; j += 256) acc += x[j] * y[j];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double g = 0.0;
        for (int w = 0; w < 8; ++w) g += red[w];
        s_g = g;
    }
    __syncthreads();
    const double h = 0.5 * s_g;
    if (!(fabs(h) > 1e-15) || fabs(h) > 1e-6) return;       // nothing to do / not a "nearly orthogonal" pair: leave it
    for (int j = threadIdx.x; j < len; j += 256) {
        const double a = x[j], b = y[j];
        x[j] = a - h * b;
        y[j] = b - h * a;
    }
}

static void tw_pairfix(double *Z, long ld, int len, int ns, cudaStream_t st)
{
    static const int kmax = getenv("SVD_GPU_PAIRFIX") ? atoi(getenv("SVD_GPU_PAIRFIX")) : 2;
    for (int k = 1; k <= kmax && k < ns; ++k)
        for (int phase = 0; phase < 2; ++phase) {
            // number of pair slots: groups of 2k vectors, k pairs each (the kernel drops pairs beyond ns)
            const int groups = (ns + 2 * k - 1) / (2 * k);
            tw_pairfix_kernel<<<groups * k, 256, 0, st>>>(Z, ld, len, ns, k, phase);
            SVD_KERNEL_CHECK();
        }
}

static void launch_tw_select(int sl, int grid, cudaStream_t st, int mb, int ns, const double *tau, const double *S,
                             const double *P, int *kidx, double *gk)
{
    if (sl == 32) tw_select_kernel<32><<<grid, 32 * 32, 0, st>>>(mb, ns, tau, S, P, kidx, gk);
    else if (sl == 16) tw_select_kernel<16><<<grid, 32 * 16, 0, st>>>(mb, ns, tau, S, P, kidx, gk);
    else tw_select_kernel<8><<<grid, 32 * 8, 0, st>>>(mb, ns, tau, S, P, kidx, gk);
}

static void launch_tw_solve(int ub, dim3 grid, cudaStream_t st, int mb, int ns, const double *q, const double *e,
                            const double *ab, const double *pivmin, const int *kidx, double *S, const double *P, double *nrm2)
{
    if (ub == 32) tw_solve_kernel<32><<<grid, 32, 0, st>>>(mb, ns, q, e, ab, pivmin, kidx, S, P, nrm2);
    else if (ub == 16) tw_solve_kernel<16><<<grid, 32, 0, st>>>(mb, ns, q, e, ab, pivmin, kidx, S, P, nrm2);
    else tw_solve_kernel<8><<<grid, 32, 0, st>>>(mb, ns, q, e, ab, pivmin, kidx, S, P, nrm2);
}

static int tw_chunk(int mb, int ns)
{
    const size_t budget = (size_t)4 << 30;                 // bytes for the two scratch panels
    long c = (long)(budget / (2 * sizeof(double) * (size_t)mb));
    c = c / 32 * 32;
    if (c < 32) c = 32;
    if (c > ns) c = ns;
    return (int)c;
}

size_t twisted_workspace_bytes(int n, int mb, int ns)
{
    int c = tw_chunk(mb, ns);
    size_t d = 0;
    d += 6 * (size_t)mb + 8;            // q, e, ab (+ the reversed set for the left vectors), pivmin
    d += 2 * (size_t)mb * c;            // S, P
    d += 4 * (size_t)c + 8;             // tau, gk, nrm2[2]
    return d * sizeof(double) + (size_t)c * sizeof(int) + 4096;
}

void twisted_vectors_device(int n, int mb, const double *a, const double *b, const double *sigma_all,
                            int ntot, int i0, int ns, double *X, long ldx, double *Y, long ldy,
                            double *sigma_out, int rqi_steps, void *workspace, cudaStream_t st)
{
    if (ns <= 0) return;
    const int cmax = tw_chunk(mb, ns);
    double *w = (double *)workspace;
    double *q = w;       w += mb;
    double *e = w;       w += mb;
    double *ab = w;      w += mb;
    double *q2 = w;      w += mb;
    double *e2 = w;      w += mb;
    double *ab2 = w;     w += mb;
    double *pivmin = w;  w += 8;
    double *S = w;       w += (size_t)mb * cmax;
    double *P = w;       w += (size_t)mb * cmax;
    double *tau = w;     w += cmax;
    double *gk = w;      w += cmax;
    double *nrm2 = w;    w += 2 * (size_t)cmax + 8;
    int *kidx = (int *)w;

    tw_prep_kernel<<<1, 1024, 0, st>>>(n, mb, a, b, q, e, ab, pivmin);
    SVD_KERNEL_CHECK();
    // left vectors from their own twisted factorization when B is square (SVD_GPU_LEFT=bx restores y = Bx/sigma)
    const char *uenv = getenv("SVD_GPU_TW_UB");
    const int tw_ub = uenv ? atoi(uenv) : 8;
    const char *senv = getenv("SVD_GPU_TW_SL");
    const int tw_sl = senv ? atoi(senv) : 8;
    const char *lenv = getenv("SVD_GPU_LEFT");
    const bool left_by_twist = (Y != nullptr) && (mb == n) && (n > 1) && !(lenv && lenv[0] == 'b');
    if (left_by_twist) {
        tw_prep_rev_kernel<<<ceil_div(n, 256), 256, 0, st>>>(n, a, b, q2, e2, ab2);
        SVD_KERNEL_CHECK();
    }
    for (int c0 = 0; c0 < ns; c0 += cmax) {
        const int c = (ns - c0 < cmax) ? ns - c0 : cmax;
        const int gi0 = i0 + c0;
        tw_tau_init_kernel<<<ceil_div(c, 256), 256, 0, st>>>(c, sigma_all + gi0, tau);
        SVD_KERNEL_CHECK();
        for (int sweep = 0; sweep <= rqi_steps; ++sweep) {
            tw_qd_kernel<<<dim3(ceil_div(c, 32), 2), 32, 0, st>>>(mb, c, q, e, tau, pivmin, S, P);
            SVD_KERNEL_CHECK();
            launch_tw_select(tw_sl, ceil_div(c, 32), st, mb, c, tau, S, P, kidx, gk);
            SVD_KERNEL_CHECK();
            launch_tw_solve(tw_ub, dim3(ceil_div(c, 32), 2), st, mb, c, q, e, ab, pivmin, kidx, S, P, nrm2);
            SVD_KERNEL_CHECK();
            if (sweep < rqi_steps) {
                tw_rqi_kernel<<<ceil_div(c, 256), 256, 0, st>>>(c, gi0, ntot, sigma_all, gk, nrm2, tau);
                SVD_KERNEL_CHECK();
            }
        }
        dim3 grid(ceil_div(c, 32), ceil_div(mb, 32));
        double *Xc = X + (size_t)c0 * ldx;
        double *Yc = Y ? Y + (size_t)c0 * ldy : nullptr;
        tw_finalize_kernel<<<grid, 256, 0, st>>>(n, mb, c, a, b, tau, S, nrm2, Xc, ldx, left_by_twist ? nullptr : Yc,
                                                 ldy, sigma_out ? sigma_out + c0 : nullptr);
        SVD_KERNEL_CHECK();
        if (left_by_twist) {
            // y_i from B B^T - sigma_i^2 I (reversed bidiagonal), with the polished sigma_i^2 already in tau
            tw_qd_kernel<<<dim3(ceil_div(c, 32), 2), 32, 0, st>>>(n, c, q2, e2, tau, pivmin, S, P);
            SVD_KERNEL_CHECK();
            launch_tw_select(tw_sl, ceil_div(c, 32), st, n, c, tau, S, P, kidx, gk);
            SVD_KERNEL_CHECK();
            launch_tw_solve(tw_ub, dim3(ceil_div(c, 32), 2), st, n, c, q2, e2, ab2, pivmin, kidx, S, P, nrm2);
            SVD_KERNEL_CHECK();
            tw_left_sign_kernel<<<ceil_div(c, 256), 256, 0, st>>>(n, c, a, b, kidx, Xc, ldx, gk);
            SVD_KERNEL_CHECK();
            tw_left_finalize_kernel<<<dim3(ceil_div(c, 32), ceil_div(n, 32)), 256, 0, st>>>(n, c, S, nrm2, gk, Yc, ldy);
            SVD_KERNEL_CHECK();
        }
    }
    tw_pairfix(X, ldx, mb, ns, st);
    if (Y != nullptr) tw_pairfix(Y, ldy, n, ns, st);
}

} // namespace svdgpu
