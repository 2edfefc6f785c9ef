// ddc.cu — "double divide and conquer" singular values of an upper bidiagonal on sm_100a.
//
// Replaces the host/OpenMP phase GetSingularValues_Parallel -> DivideAndConquer_Parallel
// (Calculations-Parallel.c:852-874, :706-850) and its pieces SolveSmallMatrices_Parallel
// (:590-703), SolveSecularEquation_Parallel (:48-348), GetTopAndBottomRows_Parallel (:351-588).
//
// Same algorithm (Konda-Nakamura dDC: split the N x (N+1) bidiagonal at row K = N/2, solve
// both halves, couple them through row K; only singular values and the first/last rows of
// the right singular vector matrices travel up the tree), re-organised for the GPU:
//   * the recursion is unrolled bottom-up: the host enumerates the tree once (sizes only)
//     and every tree level is a handful of launches that process ALL merges of that level;
//   * secular equation: ONE WARP PER ROOT; the 32 lanes stride the N poles and combine
//     with shuffles; the root is found as the offset eta = sigma^2 - d_o^2 from the
//     nearer pole with pole distances (d_j-d_o)(d_j+d_o), by a bracketed fixed-weight /
//     middle-way rational iteration (the reference iterates on gamma = 1/t and forms
//     sigma^2 = d^2 + 1/gamma, :255,:342, which loses the small singular values);
//   * Loewner/Gu-Eisenstat z-hat: one warp per pole, product of ratios (reference: sums of
//     logs, :462-481); row rotation: one warp per root;
//   * negligible coupling entries are deflated (reference: the c<1e-20 shortcut, :143);
//   * merge of the two sorted halves (:65-104) is a rank computation by binary search.
#include "common.cuh"
#include "ddc.cuh"
#include <vector>
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

namespace svdgpu {

#define DDC_EPS 2.220446049250313e-16

struct DdcDev {
    // problem
    const double *b1, *b2;
    // per position (length N)
    double *sig, *frow, *lrow;
    double *dS, *zS, *fS, *lS;          // merged order
    double *dA, *zA, *fA, *lA;          // active (non-deflated), compact per node
    double *dD, *fD, *lD;               // deflated, compact per node
    double *eta, *sigA, *zh, *fN, *lN;
    int *org;
    // per node
    const int *n_off, *n_N, *n_K, *n_c1, *n_c2;
    double *phi, *psi, *c0, *s0, *zsum;
    int *nact;
};

// ------------------------------------------------------------------------------ leaves
// closed forms for the 1x2 and 2x3 blocks (what Calculations-Parallel.c:590-703 computes),
// written through the 2x2 eigenproblem of B B^T so that the small singular value does not
// suffer the cancellation of the reference's n1 - sqrt(n1^2 - n2^2).
__global__ void ddc_leaves_kernel(DdcDev D, int first, int count)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int id = first + t, off = D.n_off[id], N = D.n_N[id];
    const double *b1 = D.b1 + off, *b2 = D.b2 + off;
    if (N == 1) {
        double a = b1[0], b = b2[0];
        double s = hypot(a, b);
        D.sig[off] = s;
        if (s > 0.0) { D.frow[off] = a / s; D.lrow[off] = b / s; D.phi[id] = b / s; D.psi[id] = -a / s; }
        else { D.frow[off] = 1.0; D.lrow[off] = 0.0; D.phi[id] = 0.0; D.psi[id] = 1.0; }
        return;
    }
    // B = [a b 0; 0 c e]
    const double a = b1[0], b = b2[0], c = b1[1], e = b2[1];
    const double p = a * a + b * b, s = c * c + e * e, r = b * c;
    const double h = 0.5 * (p - s);
    const double rad = hypot(h, r);
    const double lmax = 0.5 * (p + s) + rad;
    const double det = a * a * c * c + a * a * e * e + b * b * e * e;     // det(B B^T), no cancellation
    const double lmin = (lmax > 0.0) ? det / lmax : 0.0;
    const double smax = sqrt(lmax), smin = sqrt(lmin);
    // eigenvector of [[p r][r s]] for lmax
    double u1, u2;
    if (h >= 0.0) { u1 = h + rad; u2 = r; } else { u1 = r; u2 = rad - h; }
    double un = hypot(u1, u2);
    if (un > 0.0) { u1 /= un; u2 /= un; } else { u1 = 1.0; u2 = 0.0; }
    // right vectors v = B^T u / sigma ; first component a*u1/sigma, last e*u2/sigma
    // ascending order: index 0 <- smin (u = (-u2, u1)), index 1 <- smax (u = (u1, u2))
    D.sig[off] = smin;
    D.sig[off + 1] = smax;
    if (smin > 0.0) { D.frow[off] = a * (-u2) / smin; D.lrow[off] = e * u1 / smin; }
    else { D.frow[off] = 0.0; D.lrow[off] = 0.0; }
    if (smax > 0.0) { D.frow[off + 1] = a * u1 / smax; D.lrow[off + 1] = e * u2 / smax; }
    else { D.frow[off + 1] = 0.0; D.lrow[off + 1] = 0.0; }
    // null vector of B: (b e, -a e, a c)
    double n0 = b * e, n1 = -a * e, n2 = a * c;
    double nn = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
    if (nn > 0.0) { D.phi[id] = n0 / nn; D.psi[id] = n2 / nn; }
    else { D.phi[id] = 1.0; D.psi[id] = 0.0; }
}

// ------------------------------------------------------------------------------ prepare
__device__ __forceinline__ int lower_bound_dev(const double *x, int n, double v)
{   // number of elements < v
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (x[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int upper_bound_dev(const double *x, int n, double v)
{   // number of elements <= v
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (x[mid] <= v) lo = mid + 1; else hi = mid; }
    return lo;
}

__device__ double block_reduce_max(double v, double *red)
{
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmax(r, red[w]);
    return r;
}
__device__ double block_reduce_sum(double v, double *red)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += red[w];
    return r;
}

// One CTA per merge node: form d, z and the un-rotated first/last rows in merged (sorted)
// order, deflate, compact.  (Calculations-Parallel.c:731-782 + :65-104 + :371-448)
__global__ void __launch_bounds__(256) ddc_prepare_kernel(DdcDev D, int first)
{
    __shared__ double red[8];
    __shared__ double s_sc[4];
    __shared__ int s_cnt[2];
    __shared__ int s_wsum[8], s_wsumD[8];
    const int id = first + blockIdx.x;
    const int off = D.n_off[id], N = D.n_N[id], K = D.n_K[id];
    const int c1 = D.n_c1[id], c2 = D.n_c2[id];
    const int N2 = N - K - 1;
    const double bK = D.b1[off + K], cK = D.b2[off + K];
    if (threadIdx.x == 0) {
        double p = bK * D.psi[c1], q = cK * D.phi[c2];
        double r0 = hypot(p, q);
        double c0 = (r0 > 0.0) ? p / r0 : 1.0, s0 = (r0 > 0.0) ? q / r0 : 0.0;
        s_sc[0] = r0; s_sc[1] = c0; s_sc[2] = s0;
        D.c0[id] = c0; D.s0[id] = s0;
    }
    __syncthreads();
    const double r0 = s_sc[0], c0 = s_sc[1], s0 = s_sc[2];
    const double *run1 = D.sig + off, *run2 = D.sig + off + K + 1;
    double zmax = 0.0;
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
        double d, z, fo, lo;
        int rank;
        if (t == K) { d = 0.0; z = r0; fo = c0 * D.phi[c1]; lo = s0 * D.psi[c2]; rank = 0; }
        else if (t < K) {
            d = run1[t]; z = bK * D.lrow[off + t]; fo = D.frow[off + t]; lo = 0.0;
            rank = 1 + t + lower_bound_dev(run2, N2, d);
        } else {
            int b = t - K - 1;
            d = run2[b]; z = cK * D.frow[off + t]; fo = 0.0; lo = D.lrow[off + t];
            rank = 1 + b + upper_bound_dev(run1, K, d);
        }
        D.dS[off + rank] = d; D.zS[off + rank] = z; D.fS[off + rank] = fo; D.lS[off + rank] = lo;
        zmax = fmax(zmax, fabs(z));
    }
    zmax = block_reduce_max(zmax, red);
    __syncthreads();
    const double dmax = D.dS[off + N - 1];
    const double tol = 8.0 * DDC_EPS * fmax(dmax, zmax);

    // compaction: stable split into active / deflated, chunk by chunk
    if (threadIdx.x == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
    __syncthreads();
    double zs = 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < N; base += blockDim.x) {
        int t = base + threadIdx.x;
        bool in = t < N;
        double d = 0, z = 0, fo = 0, lo = 0;
        if (in) { d = D.dS[off + t]; z = D.zS[off + t]; fo = D.fS[off + t]; lo = D.lS[off + t]; }
        bool act = in && (fabs(z) > tol);
        bool defl = in && !act;
        unsigned ma = __ballot_sync(0xffffffffu, act), md = __ballot_sync(0xffffffffu, defl);
        int pa = __popc(ma & ((1u << lane) - 1)), pd = __popc(md & ((1u << lane) - 1));
        if (lane == 0) { s_wsum[warp] = __popc(ma); s_wsumD[warp] = __popc(md); }
        __syncthreads();
        int oa = s_cnt[0], od = s_cnt[1];
        for (int w = 0; w < warp; ++w) { oa += s_wsum[w]; od += s_wsumD[w]; }
        if (act) {
            int p = off + oa + pa;
            D.dA[p] = d; D.zA[p] = z; D.fA[p] = fo; D.lA[p] = lo;
            zs += z * z;
        } else if (defl) {
            int p = off + od + pd;
            D.dD[p] = d; D.fD[p] = fo; D.lD[p] = lo;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int ta = 0, td = 0;
            for (int w = 0; w < nw; ++w) { ta += s_wsum[w]; td += s_wsumD[w]; }
            s_cnt[0] += ta; s_cnt[1] += td;
        }
        __syncthreads();
    }
    zs = block_reduce_sum(zs, red);
    if (threadIdx.x == 0) { D.nact[id] = s_cnt[0]; D.zsum[id] = zs; }
}

// ------------------------------------------------------------------------------ secular
// One warp per root.  d ascending (active poles), z2_j = zA_j^2, zsum = sum z2.
// Returns origin o (index into the active list) and eta with sigma^2 = d_o^2 + eta.
__device__ void secular_root_warp(int i, int Na, const double *__restrict__ d,
                                  const double *__restrict__ zA, double zsum, int lane, int &org,
                                  double &eta_out)
{
    if (Na == 1) { org = 0; eta_out = zA[0] * zA[0]; return; }
    int o, plo, phi_i;
    double lo, hi, eta;
    const bool last = (i == Na - 1);
    {
        double z2a, z2b, delq, mid, dref;
        if (!last) {
            plo = i; phi_i = i + 1;
            z2a = zA[i] * zA[i]; z2b = zA[i + 1] * zA[i + 1];
            delq = (d[i + 1] - d[i]) * (d[i + 1] + d[i]);
            mid = 0.5 * delq; dref = d[i];
        } else {
            plo = Na - 2; phi_i = Na - 1;
            z2a = zA[Na - 2] * zA[Na - 2]; z2b = zA[Na - 1] * zA[Na - 1];
            delq = (d[Na - 1] - d[Na - 2]) * (d[Na - 1] + d[Na - 2]);
            mid = 0.5 * zsum; dref = d[Na - 1];
        }
        double csum = 0.0;
        for (int j = lane; j < Na; j += 32) {
            if (j == plo || j == phi_i) continue;
            double Dm = (d[j] - dref) * (d[j] + dref) - mid;
            csum += zA[j] * zA[j] / Dm;
        }
        double c = 1.0 + warp_sum(csum);
        if (!last) {
            double w = c - z2a / mid + z2b / mid;        // poles at 0 and delq, evaluated at mid
            if (w > 0.0) {
                o = i;
                double a = c * delq + z2a + z2b, b = z2a * delq;
                double disc = sqrt(fabs(a * a - 4.0 * b * c));
                eta = (a > 0.0) ? 2.0 * b / (a + disc) : (a - disc) / (2.0 * c);
                lo = 0.0; hi = mid;
            } else {
                o = i + 1;
                double a = c * delq - z2a - z2b, b = z2b * delq;
                double disc = sqrt(fabs(a * a + 4.0 * b * c));
                eta = (a < 0.0) ? 2.0 * b / (a - disc) : -(a + disc) / (2.0 * c);
                lo = -mid; hi = 0.0;
            }
            if (!(lo < eta && eta < hi)) eta = 0.5 * (lo + hi);
        } else {
            o = Na - 1;
            // poles at -delq (N-2) and 0 (N-1), evaluated at mid = zsum/2
            double w = c + z2a / (-delq - mid) + z2b / (-mid);
            double a = -c * delq + z2a + z2b, b = z2b * delq;
            double disc = sqrt(a * a + 4.0 * b * c);
            eta = (a < 0.0) ? 2.0 * b / (disc - a) : (a + disc) / (2.0 * c);
            lo = 0.0; hi = zsum;
            if (w <= 0.0) lo = mid; else hi = mid;
            if (!(lo < eta && eta <= hi)) eta = 0.5 * (lo + hi);
        }
    }
    const double d_o = d[o];
    const double Dlo = (d[plo] - d_o) * (d[plo] + d_o), Dhi = (d[phi_i] - d_o) * (d[phi_i] + d_o);
    const double z2lo = zA[plo] * zA[plo], z2hi = zA[phi_i] * zA[phi_i];
    bool use_middle = false;
    double w_prev = 0.0;
    bool have_prev = false;
    for (int it = 0; it < 64; ++it) {
        double psi = 0.0, ph = 0.0, dpsi = 0.0, dphi = 0.0, asum = 0.0;
        for (int j = lane; j < Na; j += 32) {
            double dj = d[j];
            double del = (dj - d_o) * (dj + d_o) - eta;
            double zz = zA[j];
            double t = zz * zz / del;
            double td = t / del;
            if (j <= plo) { psi += t; dpsi += td; } else { ph += t; dphi += td; }
            asum += fabs(t);
        }
        psi = warp_sum(psi); ph = warp_sum(ph); dpsi = warp_sum(dpsi); dphi = warp_sum(dphi);
        asum = warp_sum(asum);
        const double w = 1.0 + psi + ph, dw = dpsi + dphi;
        const double err = 8.0 * (1.0 + asum) + fabs(eta) * dw;
        if (fabs(w) <= DDC_EPS * err) break;
        if (w < 0.0) lo = fmax(lo, eta); else hi = fmin(hi, eta);
        if (have_prev && fabs(w) > 0.1 * fabs(w_prev)) use_middle = !use_middle;
        w_prev = w; have_prev = true;
        const double dl = Dlo - eta, dh = Dhi - eta;
        double step;
        if (!last) {
            double c;
            if (!use_middle) {
                if (o == plo) c = w - dh * dw - (Dlo - Dhi) * (z2lo / dl / dl);
                else          c = w - dl * dw - (Dhi - Dlo) * (z2hi / dh / dh);
            } else {
                c = w - dl * dpsi - dh * dphi;
            }
            double a = (dl + dh) * w - dl * dh * dw;
            double b = dl * dh * w;
            if (c == 0.0) step = (a != 0.0) ? b / a : 0.0;
            else {
                double disc = sqrt(fabs(a * a - 4.0 * b * c));
                step = (a <= 0.0) ? (a - disc) / (2.0 * c) : 2.0 * b / (a + disc);
            }
        } else {
            double c = w - dl * dpsi - dh * dphi;
            double a = (dl + dh) * w - dl * dh * dw;
            double b = dl * dh * w;
            if (c < 0.0) c = fabs(c);
            if (c == 0.0) step = hi - eta;
            else {
                double disc = sqrt(fabs(a * a - 4.0 * b * c));
                step = (a >= 0.0) ? (a + disc) / (2.0 * c) : 2.0 * b / (a - disc);
            }
        }
        if (w * step >= 0.0) step = -w / dw;
        double nw = eta + step;
        if (!(lo < nw && nw < hi)) nw = 0.5 * (lo + hi);
        if (nw == eta) break;
        eta = nw;
    }
    org = o;
    eta_out = eta;
}

__global__ void __launch_bounds__(256)
ddc_secular_kernel(DdcDev D, const int *__restrict__ pos2node, int first, int Ntot)
{
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= Ntot) return;
    const int nl = pos2node[g];
    if (nl < 0) return;
    const int id = first + nl, off = D.n_off[id];
    const int i = g - off, Na = D.nact[id];
    if (i >= Na) return;
    int o; double eta;
    secular_root_warp(i, Na, D.dA + off, D.zA + off, D.zsum[id], lane, o, eta);
    if (lane == 0) {
        double d_o = D.dA[off + o];
        D.org[g] = o;
        D.eta[g] = eta;
        D.sigA[g] = d_o + eta / (d_o + sqrt(d_o * d_o + eta));
    }
}

// ------------------------------------------------------------------------------ z-hat
// zhat_j^2 = prod_k (sigma_k^2 - d_j^2) / prod_{k != j} (d_k^2 - d_j^2)   (Loewner),
// grouped into positive ratios exactly as Calculations-Parallel.c:470-479 groups its logs.
__global__ void __launch_bounds__(256)
ddc_zhat_kernel(DdcDev D, const int *__restrict__ pos2node, int first, int Ntot)
{
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= Ntot) return;
    const int nl = pos2node[g];
    if (nl < 0) return;
    const int id = first + nl, off = D.n_off[id];
    const int j = g - off, Na = D.nact[id];
    if (j >= Na) return;
    const double *d = D.dA + off, *eta = D.eta + off;
    const int *org = D.org + off;
    const double dj = d[j];
    double p = 1.0;
    for (int k = lane; k < Na; k += 32) {
        double dok = d[org[k]];
        double S = eta[k] - (dj - dok) * (dj + dok);          // sigma_k^2 - d_j^2
        double ratio;
        if (k == Na - 1) ratio = S;
        else {
            int kk = (k < j) ? k : k + 1;
            double Dd = (d[kk] - dj) * (d[kk] + dj);          // d_kk^2 - d_j^2
            ratio = S / Dd;
        }
        p *= ratio;
    }
    p = warp_prod(p);
    if (lane == 0) {
        double z = sqrt(fabs(p));
        D.zh[g] = (D.zA[g] >= 0.0) ? z : -z;
    }
}

// ------------------------------------------------------------------------------ rows
// new first/last row entry of root i: sum_j v_i(j) * old_row(j), v_i(j) ~ zhat_j/(d_j^2-sigma_i^2)
// (Calculations-Parallel.c:541-584)
__global__ void __launch_bounds__(256)
ddc_rows_kernel(DdcDev D, const int *__restrict__ pos2node, int first, int Ntot)
{
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= Ntot) return;
    const int nl = pos2node[g];
    if (nl < 0) return;
    const int id = first + nl, off = D.n_off[id];
    const int i = g - off, Na = D.nact[id];
    if (i >= Na) return;
    const double *d = D.dA + off, *zh = D.zh + off, *fA = D.fA + off, *lA = D.lA + off;
    const double d_o = d[D.org[g]], eta = D.eta[g];
    double sq = 0.0, sf = 0.0, sl = 0.0;
    for (int j = lane; j < Na; j += 32) {
        double den = (d[j] - d_o) * (d[j] + d_o) - eta;       // d_j^2 - sigma_i^2
        double q = zh[j] / den;
        sq += q * q; sf += q * fA[j]; sl += q * lA[j];
    }
    sq = warp_sum(sq); sf = warp_sum(sf); sl = warp_sum(sl);
    if (lane == 0) {
        double inv = 1.0 / sqrt(sq);
        D.fN[g] = sf * inv;
        D.lN[g] = sl * inv;
    }
}

// ------------------------------------------------------------------------------ finish
// One CTA per node: merge the new roots with the deflated poles into ascending order, and
// (unless this is the root) renormalise rows jointly with phi/psi (:812-845).
__global__ void __launch_bounds__(256) ddc_finish_kernel(DdcDev D, int first, int want_rows)
{
    __shared__ double red[8];
    const int id = first + blockIdx.x;
    const int off = D.n_off[id], N = D.n_N[id], Na = D.nact[id], Nd = N - Na;
    const double *sa = D.sigA + off, *dd = D.dD + off;
    double sf = 0.0, sl = 0.0;
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
        double s, f = 0.0, l = 0.0;
        int rank;
        if (t < Na) {
            s = sa[t];
            rank = t + lower_bound_dev(dd, Nd, s);
            if (want_rows) { f = D.fN[off + t]; l = D.lN[off + t]; }
        } else {
            int e = t - Na;
            s = dd[e];
            rank = e + upper_bound_dev(sa, Na, s);
            if (want_rows) { f = D.fD[off + e]; l = D.lD[off + e]; }
        }
        D.sig[off + rank] = s;
        if (want_rows) { D.frow[off + rank] = f; D.lrow[off + rank] = l; sf += f * f; sl += l * l; }
    }
    if (!want_rows) return;
    sf = block_reduce_sum(sf, red);
    sl = block_reduce_sum(sl, red);
    const double ph = -D.s0[id] * D.phi[D.n_c1[id]];
    const double ps = D.c0[id] * D.psi[D.n_c2[id]];
    const double nf = sqrt(sf + ph * ph), nl = sqrt(sl + ps * ps);
    __syncthreads();
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
        D.frow[off + t] /= nf;
        D.lrow[off + t] /= nl;
    }
    if (threadIdx.x == 0) { D.phi[id] = ph / nf; D.psi[id] = ps / nl; }
}

// ------------------------------------------------------------------------------ host side
namespace {
struct HNode { int off, N, K, c1, c2, level; };
int build_tree(std::vector<HNode> &nodes, int off, int N)
{
    HNode nd; nd.off = off; nd.N = N; nd.K = 0; nd.c1 = nd.c2 = -1; nd.level = 0;
    if (N > 2) {
        nd.K = N / 2;
        nd.c1 = build_tree(nodes, off, nd.K);
        nd.c2 = build_tree(nodes, off + nd.K + 1, N - nd.K - 1);
        nd.level = 1 + std::max(nodes[nd.c1].level, nodes[nd.c2].level);
    }
    nodes.push_back(nd);
    return (int)nodes.size() - 1;
}
} // namespace

size_t ddc_workspace_bytes(int N)
{
    // 21 double arrays + 1 int array per position, 5 double + 6 int arrays per node (<= N nodes),
    // pos2node per level (<= 40 levels)
    size_t per_pos = 21 * sizeof(double) + sizeof(int);
    size_t per_node = 5 * sizeof(double) + 6 * sizeof(int);
    return (size_t)N * (per_pos + per_node + 40 * sizeof(int)) + (64 << 10);
}

// The merge tree depends on N only: its tables (node offsets / sizes / children by level, position -> node
// per level) are built once per (device, N) and kept on the device, so that a call neither copies from
// pageable host memory nor synchronises the stream (the reference recurses on the host instead,
// Calculations-Parallel.c:731-766).
struct DdcTables {
    int nn = 0, nlev = 0;
    std::vector<int> lvl_first;
    int *d_tab = nullptr;            // off | N | K | c1 | c2 (nn each) | pos2node (nlev x N)
};
static std::map<std::pair<int, int>, DdcTables> g_ddc_tables;
static std::mutex g_ddc_mu;

static const DdcTables &ddc_tables(int N)
{
    int dev = 0;
    SVD_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_ddc_mu);
    auto it = g_ddc_tables.find({dev, N});
    if (it != g_ddc_tables.end()) return it->second;
    // ---- tree (sizes only; the recursion of Calculations-Parallel.c:731-766 unrolled)
    std::vector<HNode> nodes;
    nodes.reserve((size_t)N + 8);
    const int root = build_tree(nodes, 0, N);
    const int nn = (int)nodes.size();
    const int nlev = nodes[root].level + 1;
    // order nodes by level
    std::vector<int> order(nn), newid(nn);
    for (int i = 0; i < nn; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return nodes[a].level < nodes[b].level; });
    for (int i = 0; i < nn; ++i) newid[order[i]] = i;
    std::vector<int> tab((size_t)5 * nn + (size_t)nlev * N, -1), lvl_first(nlev + 1, 0);
    int *h_off = tab.data(), *h_N = h_off + nn, *h_K = h_N + nn, *h_c1 = h_K + nn, *h_c2 = h_c1 + nn, *h_p2n = h_c2 + nn;
    for (int i = 0; i < nn; ++i) {
        const HNode &nd = nodes[order[i]];
        h_off[i] = nd.off; h_N[i] = nd.N; h_K[i] = nd.K;
        h_c1[i] = nd.c1 >= 0 ? newid[nd.c1] : -1;
        h_c2[i] = nd.c2 >= 0 ? newid[nd.c2] : -1;
        lvl_first[nd.level + 1]++;
    }
    for (int l = 0; l < nlev; ++l) lvl_first[l + 1] += lvl_first[l];
    for (int l = 1; l < nlev; ++l)
        for (int i = lvl_first[l]; i < lvl_first[l + 1]; ++i)
            for (int t = 0; t < h_N[i]; ++t) h_p2n[(size_t)l * N + h_off[i] + t] = i - lvl_first[l];
    DdcTables T;
    T.nn = nn; T.nlev = nlev; T.lvl_first = lvl_first;
    SVD_CUDA_CHECK(cudaMalloc(&T.d_tab, sizeof(int) * tab.size()));
    SVD_CUDA_CHECK(cudaMemcpy(T.d_tab, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice));
    // a handful of sizes per process in practice; keep the cache from growing without bound
    if (g_ddc_tables.size() >= 64) {
        for (auto &kv : g_ddc_tables) cudaFree(kv.second.d_tab);
        (void)cudaGetLastError();
        g_ddc_tables.clear();
    }
    return g_ddc_tables.emplace(std::make_pair(dev, N), std::move(T)).first->second;
}

void ddc_values_device(int N, const double *b1, const double *b2, double *sigma, void *workspace,
                       cudaStream_t st)
{
    const DdcTables &T = ddc_tables(N);
    const int nn = T.nn, nlev = T.nlev;
    const std::vector<int> &lvl_first = T.lvl_first;

    // ---- carve the workspace
    char *w = (char *)workspace;
    auto take = [&](size_t bytes) { char *p = w; w += (bytes + 255) / 256 * 256; return (void *)p; };
    DdcDev D;
    D.b1 = b1; D.b2 = b2;
    double **arrs[] = {&D.sig, &D.frow, &D.lrow, &D.dS, &D.zS, &D.fS, &D.lS, &D.dA, &D.zA, &D.fA, &D.lA,
                       &D.dD, &D.fD, &D.lD, &D.eta, &D.sigA, &D.zh, &D.fN, &D.lN};
    for (auto a : arrs) *a = (double *)take(sizeof(double) * (size_t)N);
    D.org = (int *)take(sizeof(int) * (size_t)N);
    D.phi = (double *)take(sizeof(double) * nn); D.psi = (double *)take(sizeof(double) * nn);
    D.c0 = (double *)take(sizeof(double) * nn);  D.s0 = (double *)take(sizeof(double) * nn);
    D.zsum = (double *)take(sizeof(double) * nn);
    D.nact = (int *)take(sizeof(int) * nn);
    int *d_off = T.d_tab, *d_N = d_off + nn, *d_K = d_N + nn, *d_c1 = d_K + nn, *d_c2 = d_c1 + nn;
    int *d_p2n = d_c2 + nn;
    D.n_off = d_off; D.n_N = d_N; D.n_K = d_K; D.n_c1 = d_c1; D.n_c2 = d_c2;
    if ((size_t)(w - (char *)workspace) > ddc_workspace_bytes(N)) {
        fprintf(stderr, "ddc_values_device: workspace too small\n"); abort();
    }

    // ---- level 0: all leaves at once
    {
        int cnt = lvl_first[1] - lvl_first[0];
        ddc_leaves_kernel<<<ceil_div(cnt, 128), 128, 0, st>>>(D, lvl_first[0], cnt);
        SVD_KERNEL_CHECK();
    }
    // ---- merges, bottom-up; every level is five launches for all of its nodes
    const int wblocks = ceil_div(N, 8);
    for (int l = 1; l < nlev; ++l) {
        const int first = lvl_first[l], cnt = lvl_first[l + 1] - first;
        const int want_rows = (l != nlev - 1);
        const int *p2n = d_p2n + (size_t)l * N;
        ddc_prepare_kernel<<<cnt, 256, 0, st>>>(D, first);
        SVD_KERNEL_CHECK();
        ddc_secular_kernel<<<wblocks, 256, 0, st>>>(D, p2n, first, N);
        SVD_KERNEL_CHECK();
        if (want_rows) {
            ddc_zhat_kernel<<<wblocks, 256, 0, st>>>(D, p2n, first, N);
            SVD_KERNEL_CHECK();
            ddc_rows_kernel<<<wblocks, 256, 0, st>>>(D, p2n, first, N);
            SVD_KERNEL_CHECK();
        }
        ddc_finish_kernel<<<cnt, 256, 0, st>>>(D, first, want_rows);
        SVD_KERNEL_CHECK();
    }
    SVD_CUDA_CHECK(cudaMemcpyAsync(sigma, D.sig, sizeof(double) * (size_t)N, cudaMemcpyDeviceToDevice, st));
}

} // namespace svdgpu
