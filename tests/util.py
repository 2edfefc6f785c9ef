"""Shared test helpers: the oracle (CPU restatement), the optional reference build, metrics."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
P = ctypes.POINTER(ctypes.c_double)
EPS = np.finfo(np.float64).eps


def p(a):
    return a.ctypes.data_as(P)


_oracle = None
_ref = False


def oracle():
    """oracle/libddcoracle.so, built on demand with gcc (test infrastructure only)."""
    global _oracle
    if _oracle is None:
        so = os.path.join(ORACLE_DIR, "libddcoracle.so")
        src = os.path.join(ORACLE_DIR, "svd_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(so)
        L.orc_fill_rand.argtypes = [P, ctypes.c_long, ctypes.c_double, ctypes.c_double, ctypes.c_int]
        L.orc_bidiag.argtypes = [ctypes.c_int, ctypes.c_int, P, P, P]
        L.orc_ddc_values.argtypes = [ctypes.c_int, P, P, P]
        L.orc_right_vectors.argtypes = [ctypes.c_int, ctypes.c_int, P, P, P, P]
        L.orc_left_vectors.argtypes = [ctypes.c_int, ctypes.c_int, P, P, P, P, P]
        L.orc_apply_left.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, P, P, P]
        L.orc_apply_right.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, P, P, P]
        L.orc_form_u.argtypes = [ctypes.c_int, ctypes.c_int, P, P]
        L.orc_form_v.argtypes = [ctypes.c_int, ctypes.c_int, P, P]
        L.orc_svd.argtypes = [ctypes.c_int, ctypes.c_int, P, P, P, P]
        L.orc_last_timings.argtypes = [P]
        _oracle = L
    return _oracle


def reference():
    """oracle/_ref/libddcref.so (the unmodified reference host code) or None when absent."""
    global _ref
    if _ref is False:
        so = os.path.join(ORACLE_DIR, "_ref", "libddcref.so")
        if not os.path.exists(so) and os.path.isdir("/root/reference"):
            subprocess.call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL,
                            stderr=subprocess.DEVNULL)
        _ref = ctypes.CDLL(so) if os.path.exists(so) else None
    return _ref


def rand_matrix(m, n, lo=1.0, hi=4.0, seed=1):
    """The reference drivers' input recipe (test-whole-svd.c:18-24,69-73 with glibc's default
    seed 1; bidiag_dr.c uses [1,2) with srand(4)).  Returned column-major (Fortran order)."""
    buf = np.empty(m * n)
    oracle().orc_fill_rand(p(buf), m * n, lo, hi, seed)
    return np.asfortranarray(buf.reshape(n, m).T)


def oracle_bidiag(A):
    Af = np.array(A, order="F", dtype=np.float64, copy=True)
    m, n = Af.shape
    mn = min(m, n)
    alpha = np.zeros(mn)
    beta = np.zeros(mn + 1)
    oracle().orc_bidiag(m, n, p(Af), p(alpha), p(beta))
    return Af, alpha, beta[: (n - 1 if m >= n else m)]


def oracle_svd(A):
    Af = np.array(A, order="F", dtype=np.float64, copy=True)
    m, n = Af.shape
    mn = min(m, n)
    sigma = np.zeros(mn)
    U = np.zeros((m, m), order="F")
    V = np.zeros((n, n), order="F")
    oracle().orc_svd(m, n, p(Af), p(sigma), p(U), p(V))
    return sigma, U[:, :mn], V[:, :mn], Af


def reference_svd(ref, A):
    """The reference's svd_gpu() sequence (svd_gpu.c:100-121) called phase by phase on the
    reference build, with beta explicitly zero-padded to min(m,n) entries: svd_gpu.c:76,106
    lets the dDC recursion read beta[mn-1] from whatever follows the malloc'd block (SURVEY.md
    fact 4), which makes a direct call to its svd_gpu() depend on heap state.  Square only."""
    Af = np.array(A, order="F", dtype=np.float64, copy=True)
    n = Af.shape[0]
    assert Af.shape == (n, n)
    alpha = np.zeros(n); beta = np.zeros(n)
    ref.bidiag_seq(n, n, p(Af), p(alpha), p(beta))
    beta[n - 1] = 0.0
    AT = np.asfortranarray(Af.T)
    sigma = np.zeros(n)
    ref.GetSingularValues_Parallel(n, p(alpha), p(beta), p(sigma))
    X = np.zeros(n * n); Y = np.zeros(n * n)
    ref.CalcRightSingularVectors(n, n, p(alpha), p(beta), p(sigma), p(X))
    ref.RighttoLeftSingularVectors(n, n, p(alpha), p(beta), p(sigma), p(X), p(Y))
    U = np.zeros((n, n), order="F"); V = np.zeros((n, n), order="F")
    for i in range(n):
        u = np.zeros(n); v = np.zeros(n)
        ref.multU(n, n, i, p(Af), p(Y), p(u))
        ref.multV(n, n, i, p(AT), p(X), p(v))
        U[:, i] = u; V[:, i] = v
    return dict(A_mod=Af, alpha=alpha, beta=beta, sigma=sigma, X=X.reshape(n, n), Y=Y.reshape(n, n), U=U, V=V)


def bidiag_dense(alpha, beta, ncols=None):
    n = len(alpha)
    ncols = n if ncols is None else ncols
    B = np.zeros((n, ncols))
    B[np.arange(n), np.arange(n)] = alpha
    k = min(len(beta), ncols - 1)
    B[np.arange(k), np.arange(k) + 1] = beta[:k]
    return B


def svd_metrics(A, sigma, U, V):
    """north-star metrics against LAPACK: normwise sigma error, orthogonality, residual."""
    A = np.asarray(A)
    sv = np.linalg.svd(A, compute_uv=False)[::-1]
    mn = len(sigma)
    out = {
        "sigma_abs_over_max": float(np.abs(sigma - sv).max() / sv.max()),
        "sigma_rel": float(np.abs(sigma / sv - 1).max()),
        "orthU": float(np.linalg.norm(U.T @ U - np.eye(mn))),
        "orthV": float(np.linalg.norm(V.T @ V - np.eye(mn))),
        "resid": float(np.linalg.norm(A - (U * sigma) @ V.T) / np.linalg.norm(A)),
        "ascending": bool(np.all(np.diff(sigma) >= 0)),
    }
    return out
