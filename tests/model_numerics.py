"""numpy model of the numerics the CUDA kernels implement (TEST INFRASTRUCTURE).

The CUDA path keeps the reference's *algorithm family* (dDC singular values,
Calculations-Parallel.c; twisted-factorization vectors, parallel-twisted.c) but
replaces the numerically fragile pieces so the north-star bounds
(c*eps*max(m,n)) can be met (SURVEY.md fact 3, section 7 "Hard parts"):

  * secular roots are solved for the offset eta = sigma^2 - d_o^2 from the
    nearest pole with pole distances formed as (d_j-d_o)(d_j+d_o)
    (instead of sigma^2 = d^2 + 1/gamma, Calculations-Parallel.c:255,342);
  * the recomputed z (Loewner / Gu-Eisenstat) is a running product of ratios
    (instead of sums of logs, :462-481), the row rotation uses it directly;
  * negligible z components are deflated (the reference's c<1e-20 shortcut, :143);
  * vectors come from differential qd transforms of B^T B (dstqds / dpqds) with
    one Rayleigh-quotient correction of sigma^2 (instead of LDL^T on the formed
    tridiagonal plus a fixed-index solve, parallel-twisted.c:339-424).

This file is the executable specification used by the CPU tests and by the GPU
parity tests (same inputs -> same outputs to rounding).  It is pure numpy and is
never imported by the product.
"""
import numpy as np

EPS = np.finfo(np.float64).eps


# --------------------------------------------------------------------------- dDC
def secular_root(i, d, z2, zsum):
    """Root i of 1 + sum_j z2_j/(d_j^2 - t) = 0, d ascending (all poles active).

    Returns (origin o, eta) with t = d_o^2 + eta.  Safeguarded two-pole rational
    iteration in the style of LAPACK dlaed4/dlasd4 (fixed-weight, then middle-way).
    """
    N = d.shape[0]
    if i < N - 1:
        ip1 = i + 1
        delq = (d[ip1] - d[i]) * (d[ip1] + d[i])
        mid = 0.5 * delq
        Dm = (d - d[i]) * (d + d[i]) - mid
        t = z2 / Dm
        c = 1.0 + (t[:i].sum() + t[ip1 + 1:].sum())
        w = c + t[i] + t[ip1]
        if w > 0:
            o = i
            a = c * delq + z2[i] + z2[ip1]
            b = z2[i] * delq
            disc = np.sqrt(abs(a * a - 4.0 * b * c))
            eta = 2.0 * b / (a + disc) if a > 0 else (a - disc) / (2.0 * c)
            lo, hi = 0.0, mid
        else:
            o = ip1
            a = c * delq - z2[i] - z2[ip1]
            b = z2[ip1] * delq
            disc = np.sqrt(abs(a * a + 4.0 * b * c))
            eta = 2.0 * b / (a - disc) if a < 0 else -(a + disc) / (2.0 * c)
            lo, hi = -mid, 0.0
        if not (lo < eta < hi):
            eta = 0.5 * (lo + hi)
        pole_lo, pole_hi = i, ip1
    else:
        o = N - 1
        im1 = N - 2
        delq = (d[o] - d[im1]) * (d[o] + d[im1])
        lo, hi = 0.0, zsum
        # start from the two-pole model with everything else lumped into c at eta = zsum/2
        mid = 0.5 * zsum
        Dm = (d - d[o]) * (d + d[o]) - mid
        t = z2 / Dm
        c = 1.0 + t[:im1].sum()
        w = c + t[im1] + t[o]
        a = -c * delq + z2[im1] + z2[o]
        b = z2[o] * delq
        disc = np.sqrt(a * a + 4.0 * b * c)
        eta = 2.0 * b / (disc - a) if a < 0 else (a + disc) / (2.0 * c)
        if w <= 0:
            lo = mid
        else:
            hi = mid
        if not (lo < eta <= hi):
            eta = 0.5 * (lo + hi)
        pole_lo, pole_hi = im1, o

    D = (d - d[o]) * (d + d[o])
    use_middle = False
    w_prev = None
    for it in range(60):
        delta = D - eta
        t = z2 / delta
        psi = t[:pole_lo + 1].sum()
        phi = t[pole_hi:].sum() if pole_hi > pole_lo else 0.0
        dpsi = (t[:pole_lo + 1] / delta[:pole_lo + 1]).sum()
        dphi = (t[pole_hi:] / delta[pole_hi:]).sum()
        w = 1.0 + psi + phi
        dw = dpsi + dphi
        err = 8.0 * (1.0 + np.abs(t).sum()) + abs(eta) * dw
        if abs(w) <= EPS * err:
            break
        if w < 0:
            lo = max(lo, eta)
        else:
            hi = min(hi, eta)
        if w_prev is not None and abs(w) > 0.1 * abs(w_prev):
            use_middle = not use_middle
        w_prev = w
        dl, dh = delta[pole_lo], delta[pole_hi]
        if i < N - 1:
            if not use_middle:
                if o == i:
                    c = w - dh * dw - (D[pole_lo] - D[pole_hi]) * (z2[pole_lo] / dl / dl)
                else:
                    c = w - dl * dw - (D[pole_hi] - D[pole_lo]) * (z2[pole_hi] / dh / dh)
            else:
                c = w - dl * dpsi - dh * dphi
            a = (dl + dh) * w - dl * dh * dw
            b = dl * dh * w
            if c == 0.0:
                step = b / a if a != 0 else 0.0
            else:
                disc = np.sqrt(abs(a * a - 4.0 * b * c))
                step = (a - disc) / (2.0 * c) if a <= 0 else 2.0 * b / (a + disc)
        else:
            # last root: psi = poles <= N-2 ... here pole_lo = N-2, pole_hi = N-1
            c = w - dl * dpsi - dh * dphi
            a = (dl + dh) * w - dl * dh * dw
            b = dl * dh * w
            if c < 0:
                c = abs(c)
            if c == 0.0:
                step = hi - eta
            else:
                disc = np.sqrt(abs(a * a - 4.0 * b * c))
                step = (a + disc) / (2.0 * c) if a >= 0 else 2.0 * b / (a - disc)
        if w * step >= 0:
            step = -w / dw
        new = eta + step
        if not (lo < new < hi) or not np.isfinite(new):
            new = 0.5 * (lo + hi)
        if new == eta:
            break
        eta = new
    return o, eta


def merge_node(d, z, first_old, last_old, want_rows=True):
    """One dDC merge on sorted poles d (d[0]=0) with coupling vector z.

    Returns sigma ascending and (if want_rows) the rotated first/last rows.
    """
    N = d.shape[0]
    tol = 8.0 * EPS * max(d[-1], np.abs(z).max())
    act = np.abs(z) > tol
    act_idx = np.nonzero(act)[0]
    da = d[act_idx]
    za = z[act_idx]
    z2 = za * za
    zsum = z2.sum()
    Na = act_idx.shape[0]
    sig_a = np.empty(Na)
    org = np.empty(Na, dtype=np.int64)
    eta = np.empty(Na)
    if Na == 1:
        org[0] = 0
        eta[0] = z2[0]
    else:
        for i in range(Na):
            org[i], eta[i] = secular_root(i, da, z2, zsum)
    do = da[org]
    sig_a = do + eta / (do + np.sqrt(do * do + eta))
    # merged, sorted output (deflated poles pass through unchanged)
    sigma_all = np.empty(N)
    sigma_all[act_idx] = sig_a
    sigma_all[~act] = d[~act]
    order = np.argsort(sigma_all, kind="stable")
    sigma = sigma_all[order]
    if not want_rows:
        return sigma, None, None, int(N - Na)
    # Loewner z-hat on the active set: S[k,j] = sigma_k^2 - d_j^2 = eta_k - (d_j-d_ok)(d_j+d_ok)
    S = eta[:, None] - (da[None, :] - do[:, None]) * (da[None, :] + do[:, None])
    Dd = (da[:, None] - da[None, :]) * (da[:, None] + da[None, :])      # d_k^2 - d_j^2
    zh = np.empty(Na)
    for j in range(Na):
        p = S[Na - 1, j]
        if j > 0:
            p *= np.prod(S[:j, j] / Dd[:j, j])
        if j < Na - 1:
            p *= np.prod(S[j:Na - 1, j] / Dd[j + 1:, j])
        zh[j] = np.sqrt(abs(p)) * (1.0 if za[j] >= 0 else -1.0)
    Q = zh[None, :] / (-S)                       # v_i(j) ~ zhat_j / (d_j^2 - sigma_i^2)
    nrm = np.sqrt((Q * Q).sum(axis=1))
    f_new = np.array(first_old, dtype=np.float64).copy()
    l_new = np.array(last_old, dtype=np.float64).copy()
    f_new[act_idx] = (Q @ first_old[act_idx]) / nrm
    l_new[act_idx] = (Q @ last_old[act_idx]) / nrm
    return sigma, f_new[order], l_new[order], int(N - Na)


def leaf(b1, b2):
    """Closed forms for 1x2 and 2x3 blocks (what Calculations-Parallel.c:590-703 computes),
    written through the 2x2 symmetric eigenproblem of B B^T / null vector of B."""
    N = b1.shape[0]
    if N == 1:
        s = np.hypot(b1[0], b2[0])
        return (np.array([s]), np.array([b1[0] / s]), np.array([b2[0] / s]), b2[0] / s, -b1[0] / s)
    # B = [[a, b, 0], [0, c, e]]
    a, b, c, e = b1[0], b2[0], b1[1], b2[1]
    B = np.array([[a, b, 0.0], [0.0, c, e]])
    # right singular vectors through the 3x3 SVD done explicitly with numpy (tiny)
    U, s, Vt = np.linalg.svd(B)
    sig = s[::-1].copy()
    V = Vt[:2][::-1]                       # rows: right singular vectors, ascending sigma
    nullv = Vt[2]
    return sig, V[:, 0].copy(), V[:, 2].copy(), nullv[0], nullv[2]


def ddc_values(b1, b2, stats=None):
    """Singular values (ascending) of the N x (N+1) upper bidiagonal (diag b1, super b2)."""
    def node(lo, N, want_rows):
        if N <= 2:
            return leaf(b1[lo:lo + N], b2[lo:lo + N])
        K = N // 2
        s1, f1, l1, phi1, psi1 = node(lo, K, True)
        s2, f2, l2, phi2, psi2 = node(lo + K + 1, N - K - 1, True)
        bk, ck = b1[lo + K], b2[lo + K]
        p, q = bk * psi1, ck * phi2
        r0 = np.hypot(p, q)
        c0, s0 = p / r0, q / r0
        d = np.concatenate(([0.0], s1, s2))
        z = np.concatenate(([r0], bk * l1, ck * f2))
        fo = np.concatenate(([c0 * phi1], f1, np.zeros(N - K - 1)))
        lo_ = np.concatenate(([s0 * psi2], np.zeros(K), l2))
        order = np.argsort(d, kind="stable")
        sig, fn, ln, ndefl = merge_node(d[order], z[order], fo[order], lo_[order], want_rows)
        if stats is not None:
            stats.append((N, ndefl))
        if not want_rows:
            return sig, None, None, 0.0, 0.0
        phi, psi = -s0 * phi1, c0 * psi2
        nf = np.sqrt((fn * fn).sum() + phi * phi)
        nl = np.sqrt((ln * ln).sum() + psi * psi)
        return sig, fn / nf, ln / nl, phi / nf, psi / nl
    return node(0, b1.shape[0], False)[0]


# ----------------------------------------------------------------- twisted vectors
def twisted_vectors(a, b, sigma, rqi_steps=1):
    """Right vectors X[i, :] of the square upper bidiagonal (diag a[n], super b[n-1]) and the
    polished sigma.  dstqds/dpqds on B^T B = L diag(a^2) L^T, twist at argmin|gamma|."""
    n = a.shape[0]
    q = a * a
    e = np.concatenate((b * b, [0.0]))
    ab = np.concatenate((a[:-1] * b, [0.0]))
    ns = sigma.shape[0]
    tau = sigma * sigma
    X = np.empty((ns, n))
    for sweep in range(rqi_steps + 1):
        s = np.empty((n, ns))
        p = np.empty((n, ns))
        dplus = np.empty((n, ns))
        dminus = np.empty((n, ns))
        sj = -tau
        for j in range(n):
            s[j] = sj
            dplus[j] = q[j] + sj
            sj = sj * (e[j] / dplus[j]) - tau
        pj = q[n - 1] - tau
        p[n - 1] = pj
        dminus[n - 1] = pj                     # unused
        for j in range(n - 2, -1, -1):
            dm = e[j] + pj                     # d-_{j+1}
            dminus[j + 1] = dm
            pj = pj * (q[j] / dm) - tau
            p[j] = pj
        gamma = s + p + tau[None, :]
        k = n - 1 - np.argmin(np.abs(gamma[::-1]), axis=0)       # later index on ties
        zvec = np.zeros((n, ns))
        cols = np.arange(ns)
        zvec[k, cols] = 1.0
        # downwards j<k : z_j = -(ab_j/d+_j) z_{j+1} ; upwards j>=k : z_{j+1} = -(ab_j/d-_{j+1}) z_j
        for j in range(n - 2, -1, -1):
            m = j < k
            zvec[j, m] = -(ab[j] / dplus[j, m]) * zvec[j + 1, m]
        for j in range(0, n - 1):
            m = j >= k
            zvec[j + 1, m] = -(ab[j] / dminus[j + 1, m]) * zvec[j, m]
        nrm2 = (zvec * zvec).sum(axis=0)
        gk = gamma[k, cols]
        if sweep < rqi_steps:
            tau = tau + gk / nrm2
    X = (zvec / np.sqrt(nrm2)[None, :]).T
    return X, np.sqrt(tau)


def left_from_right(a, b, sigma, X):
    Y = X * a[None, :]
    Y[:, :-1] += X[:, 1:] * b[None, :]
    return Y / sigma[:, None]
