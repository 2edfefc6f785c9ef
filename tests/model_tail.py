"""Executable specification of the on-chip bidiagonalization tail (ddc_svd_b200/csrc/bidiag_tail.cuh).

A numpy walk through exactly the data flow of `bidiag_tail_kernel`: G "CTAs" own the trailing columns
round-robin, every phase between two exchanges is evaluated for all CTAs before the next one starts
(the kernel's tagged-slot exchanges are bulk-synchronous in effect), the cross-CTA buffers are the
kernel's (W partials, RR, R1, A1, X, C).  The kernel was written from this model; tests/test_host_logic.py
checks the model against the oracle (bidiag.c:73-183 restated in oracle/svd_oracle.c), the GPU tests check
the kernel against the same oracle.  Test infrastructure only.
"""
import numpy as np


def make_refl(x0, nrm2):
    """update_scale_matcol.cl:58-82 / bidiag.cu:make_refl: s*nu and 1/(sqrt(2) sqrt(nu^2 + |nu x0|))."""
    nu = np.sqrt(nrm2)
    s = -1.0 if x0 < 0 else 1.0
    sc = np.sqrt(2.0) * np.sqrt(nu * nu + abs(nu * x0))
    return s * nu, (1.0 / sc if sc > 0 else 0.0)


def tail_model(Ain, i0=0, G=7, rows_per_slice_max=16):
    """Bidiagonalize the trailing block (rows, columns >= i0) of the m x n matrix (m >= n) the way the
    kernel does; returns (A with the reflectors stored in place, alpha, beta[n-1])."""
    A = np.array(Ain, order="F", dtype=float)
    m, n = A.shape
    assert m >= n
    L0 = m - i0
    alpha = np.zeros(n)
    beta = np.zeros(max(n, 1))
    ncol = [(n - i0 - b + G - 1) // G for b in range(G)]           # own columns j = i0 + b + q G
    a = [np.zeros((max(ncol[b], 0), L0)) for b in range(G)]        # shared memory of CTA b
    for b in range(G):
        for q in range(ncol[b]):
            a[b][q, :] = A[i0:, i0 + b + q * G]
    s_u = [np.zeros(max(ncol[b], 1)) for b in range(G)]            # u of the previous step (pending right update)
    s_r = [np.zeros(max(ncol[b], 1)) for b in range(G)]
    s_t = [np.zeros(max(ncol[b], 1)) for b in range(G)]
    c = A[i0:, i0].copy()                                          # current column, identical in every CTA
    x = np.zeros(L0)
    cc = float(c @ c)
    ci = c[0]
    W = np.full((G, L0), np.nan)
    A1 = np.full(L0, np.nan)
    X = np.full(L0, np.nan)
    C = np.full(L0, np.nan)
    RR = np.full(G, np.nan)
    for i in range(i0, n):
        lo = i - i0
        snu, inv = make_refl(ci, cc)
        v = np.zeros(L0)
        v[lo:] = c[lo:] * inv
        v[lo] = (c[lo] + snu) * inv
        A[i:, i] = v[lo:]                                          # stored by the CTA that owned column i
        alpha[i] = -snu
        if i == n - 1:
            break
        vi = (ci + snu) * inv
        R1 = None
        for b in range(G):                                         # ---- sweeps, local to each CTA
            qlo = (lo - b + G) // G
            w = np.zeros(L0)
            rr = 0.0
            for q in range(qlo, ncol[b]):
                col = a[b][q]
                col[lo:] -= x[lo:] * s_u[b][q]                     # sweep 1: pending right update ...
                s_t[b][q] = v[lo:] @ col[lo:]                      # ... and column dot
                s_r[b][q] = col[lo] - 2.0 * vi * s_t[b][q]         # row i after H_i
            for q in range(qlo, ncol[b]):
                col = a[b][q]
                col[lo + 1:] -= v[lo + 1:] * (2.0 * s_t[b][q])     # sweep 2: left update
                w[lo + 1:] += col[lo + 1:] * s_r[b][q]             # partial A r
                rr += s_r[b][q] ** 2
            W[b, lo + 1:] = w[lo + 1:]
            RR[b] = rr
            if b == (lo + 1) % G:                                  # owner of column i+1
                q1 = (lo + 1) // G
                A1[lo + 1:] = a[b][q1][lo + 1:]
                R1 = s_r[b][q1]
        # ---- exchange 1; row reflector scalars, identical everywhere
        has_row = i < n - 2
        gsnu, ginv = make_refl(R1, float(RR.sum())) if has_row else (0.0, 0.0)
        u1 = (R1 + gsnu) * ginv
        beta[i] = -gsnu if has_row else R1
        for b in range(G):
            qlo = (lo - b + G) // G
            for q in range(qlo, ncol[b]):
                j = i0 + b + q * G
                s_u[b][q] = (s_r[b][q] + (gsnu if j == i + 1 else 0.0)) * ginv
                A[i, j] = s_u[b][q]                                # u_i in place
            Lb = m - i - 1
            S = (Lb + G - 1) // G
            assert S <= rows_per_slice_max
            for wrp in range(S):                                   # one warp per row of this CTA's slice
                rl = b * S + wrp
                if rl < Lb:
                    r = lo + 1 + rl
                    xx = 2.0 * ginv * (W[:, r].sum() + gsnu * A1[r])
                    X[r] = xx
                    C[r] = A1[r] - xx * u1
        # ---- exchange 2: every CTA takes x and c'
        x = np.zeros(L0)
        c = np.zeros(L0)
        x[lo + 1:] = X[lo + 1:]
        c[lo + 1:] = C[lo + 1:]
        cc = float(c @ c)
        ci = c[lo + 1]
    return A, alpha, beta[: n - 1]
