"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against
  (1) the oracle (CPU restatement of the reference) on the same seeded inputs,
  (2) the committed golden vectors produced by the unmodified reference,
  (3) LAPACK (numpy) for the north-star bounds  c * eps * max(m, n),
  (4) size-independent properties at the benchmark sizes.

Tolerances (written here once):
  * bidiagonalization: overwritten A, alpha, beta within 1e-9 elementwise of the oracle on the
    reference's own check input (uniform [1,2), srand(4)) — the criterion of bidiag_dr.c:94,166-172;
  * singular values vs the reference: |sigma - sigma_ref| <= 1e-6 * sigma_max (the reference's own
    error band; it is ~1e-7 * sigma_max off LAPACK);
  * vs LAPACK: max|sigma - sigma_L| / sigma_max <= 10 eps max(m,n); Frobenius ||U^T U - I||,
    ||V^T V - I||, ||A - U S V^T|| / ||A|| <= 100 eps max(m,n).
"""
import os

import numpy as np
import pytest

import util
from util import EPS

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def D():
    import ddc_svd_b200 as D
    L = D.lib()
    assert L.svdgpu_device_count() >= 1
    return D


def check_lapack_bounds(A, sigma, U, V, c_sig=10.0, c_vec=100.0):
    m, n = A.shape
    met = util.svd_metrics(A, sigma, U, V)
    e = EPS * max(m, n)
    assert met["ascending"], "sigma must be ascending (Calculations-Parallel.c:56)"
    assert met["sigma_abs_over_max"] <= c_sig * e, met
    assert met["orthU"] <= c_vec * e and met["orthV"] <= c_vec * e, met
    assert met["resid"] <= c_vec * e, met
    return met


# ------------------------------------------------------------------ phase 1: bidiagonalization
@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (3, 3), (8, 8), (33, 33), (64, 64), (65, 65), (200, 200),
                                   (513, 512), (300, 200), (200, 300), (97, 3), (3, 97), (40, 1), (1, 40),
                                   (129, 128), (128, 129)])
def test_bidiag_vs_oracle(D, shape):
    m, n = shape
    A = util.rand_matrix(m, n, 1.0, 2.0, 4)                   # bidiag_dr.c:77,92-93
    Ao, ao, bo = util.oracle_bidiag(A)
    Ag, ag, bg = D.bidiag_par(A)
    assert not np.isnan(Ag).any()
    assert np.abs(Ag - Ao).max() <= 1e-9                       # bidiag_dr.c:94 tol, elementwise
    assert np.abs(ag - ao).max() <= 1e-9
    if len(bo):
        assert np.abs(bg - bo).max() <= 1e-9


@pytest.mark.parametrize("name", ["ref_bidiag_80x80", "ref_bidiag_90x60", "ref_bidiag_60x90"])
def test_bidiag_vs_golden_reference(D, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    Ag, ag, bg = D.bidiag_par(g["A"])
    assert np.abs(Ag - g["A_mod"]).max() <= 1e-9
    assert np.abs(ag - g["alpha"]).max() <= 1e-9 and np.abs(bg - g["beta"]).max() <= 1e-9


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (3, 3), (5, 4), (33, 33), (64, 64), (200, 200), (513, 512),
                                   (300, 200), (97, 3), (40, 1), (1500, 1400), (2040, 1100), (2100, 2000),
                                   (2500, 2500)])
@pytest.mark.parametrize("tail", ["1", "2"])       # version 2 is the default since round 2, version 1 stays as a cross-check
def test_bidiag_on_chip_tail(D, shape, tail, monkeypatch):
    # bidiag_tail.cuh: once the trailing block fits the SMs' shared memory the rest of the factorization runs
    # in one cooperative launch (small inputs entirely, larger ones from the first panel boundary that fits)
    monkeypatch.setenv("SVD_GPU_TAIL", tail)
    m, n = shape
    A = util.rand_matrix(m, n, 1.0, 2.0, 4)
    Ao, ao, bo = util.oracle_bidiag(A)
    Ag, ag, bg = D.bidiag_par(A)
    assert not np.isnan(Ag).any()
    assert np.abs(Ag - Ao).max() <= 1e-9
    assert np.abs(ag - ao).max() <= 1e-9
    if bo.size:
        assert np.abs(bg - bo).max() <= 1e-9
    monkeypatch.setenv("SVD_GPU_TAIL", "0")
    A0, a0, b0 = D.bidiag_par(A)
    assert np.abs(Ag - A0).max() <= 1e-9 and np.abs(ag - a0).max() <= 1e-9


def test_svd_gpu_with_on_chip_tail(D, monkeypatch):
    monkeypatch.setenv("SVD_GPU_TAIL", "1")
    for shape in [(700, 700), (2304, 2304), (3000, 1200)]:
        A = util.rand_matrix(*shape)
        sigma, U, V, _ = D.svd_gpu(A)
        check_lapack_bounds(A, sigma, U, V)


@pytest.mark.parametrize("nb", [1, 7, 32, 64])
def test_bidiag_panel_width_invariance(D, nb):
    # the deferred-update panel width is an implementation detail: results must not depend on it
    A = util.rand_matrix(150, 130, 1.0, 2.0, 4)
    Ao, ao, bo = util.oracle_bidiag(A)
    os.environ["SVD_GPU_NB"] = str(nb)
    try:
        Ag, ag, bg = D.bidiag_par(A)
    finally:
        del os.environ["SVD_GPU_NB"]
    assert np.abs(Ag - Ao).max() <= 1e-9 and np.abs(ag - ao).max() <= 1e-9 and np.abs(bg - bo).max() <= 1e-9


@pytest.mark.parametrize("shape", [(1500, 1400), (4200, 300), (8704, 256), (16500, 130), (33000, 100)])
@pytest.mark.parametrize("xw", ["1", "0", "2"])     # finish_xw (128 rows per CTA at once): long columns / never / always
def test_bidiag_fused_pass_vs_oracle(D, shape, xw, monkeypatch):
    # tall enough for the fused single-read pass (bidiag_fused.cuh): 1 CTA per column tile up to
    # 8192 rows, clusters of 2 / 4 / 8 CTAs (DSMEM exchange) above
    monkeypatch.setenv("SVD_GPU_XW", xw)
    m, n = shape
    A = util.rand_matrix(m, n, 1.0, 2.0, 4)
    Ao, ao, bo = util.oracle_bidiag(A)
    os.environ["SVD_GPU_FUSED"] = "1"
    try:
        Ag, ag, bg = D.bidiag_par(A)                           # fused single-read pass
    finally:
        del os.environ["SVD_GPU_FUSED"]
    assert np.abs(Ag - Ao).max() <= 1e-9 and np.abs(ag - ao).max() <= 1e-9 and np.abs(bg - bo).max() <= 1e-9
    os.environ["SVD_GPU_FUSED"] = "0"
    try:
        As, a_s, b_s = D.bidiag_par(A)                         # split gemvT/gemvN passes
    finally:
        del os.environ["SVD_GPU_FUSED"]
    assert np.abs(Ag - As).max() <= 1e-10 and np.abs(ag - a_s).max() <= 1e-10 * max(1.0, np.abs(ag).max())


@pytest.mark.parametrize("shape", [(9000, 3000), (6144, 6144), (17000, 2048)])
def test_bidiag_fused_vs_split_many_tiles(D, shape):
    # long pipelines (tens of tiles per cluster, clusters of 1/2/4/8): the fused pass must agree with
    # the split passes and conserve the Frobenius norm (||B||_F = ||A||_F) — catches stage-reuse races
    m, n = shape
    rng = np.random.default_rng(3)
    A = np.asfortranarray(rng.uniform(1.0, 2.0, size=(m, n)))
    fro2 = float(np.sum(A * A))
    out = {}
    for mode in ("1", "0"):
        os.environ["SVD_GPU_FUSED"] = mode
        try:
            out[mode] = D.bidiag_par(A)
        finally:
            del os.environ["SVD_GPU_FUSED"]
    (Af, af, bf), (As, a_s, b_s) = out["1"], out["0"]
    for al, be in ((af, bf), (a_s, b_s)):
        assert abs(np.sum(al * al) + np.sum(be * be) - fro2) <= 1e-12 * fro2
    # the two paths add the same terms in different orders; over thousands of steps the reflectors
    # drift apart like n * eps * (growth), a few 1e-9 at n = 6144 — a race shows up as >= 1e-6
    assert np.abs(Af - As).max() <= 1e-8
    assert np.abs(af - a_s).max() <= 1e-9 * np.abs(a_s).max() and np.abs(bf - b_s).max() <= 1e-9 * np.abs(a_s).max()


@pytest.mark.parametrize("shape", [(300, 200), (200, 300), (513, 512), (97, 80)])
def test_bidiag_fused_pass_small_tiles(D, shape):
    # force the fused pass on small trailing blocks too (all tile shapes, ragged last tiles)
    m, n = shape
    A = util.rand_matrix(m, n, 1.0, 2.0, 4)
    Ao, ao, bo = util.oracle_bidiag(A)
    os.environ["SVD_GPU_FUSED_MIN_ROWS"] = "8"; os.environ["SVD_GPU_FUSED_MIN_COLS"] = "3"
    os.environ["SVD_GPU_FUSED"] = "1"
    try:
        Ag, ag, bg = D.bidiag_par(A)
    finally:
        del os.environ["SVD_GPU_FUSED_MIN_ROWS"]; del os.environ["SVD_GPU_FUSED_MIN_COLS"]; del os.environ["SVD_GPU_FUSED"]
    assert not np.isnan(Ag).any()
    assert np.abs(Ag - Ao).max() <= 1e-9 and np.abs(ag - ao).max() <= 1e-9 and np.abs(bg - bo).max() <= 1e-9


def test_bidiag_reconstruction_property(D):
    # A = Q_L B Q_R^T with the stored reflectors: checked through the back-transform entry point
    m, n = 260, 200
    A = util.rand_matrix(m, n)
    Ag, al, be = D.bidiag_par(A)
    B = util.bidiag_dense(al, be)
    U, V = D.backtransform(Ag, np.eye(n), np.eye(n))           # Q_L[:, :n], Q_R
    assert np.linalg.norm(U.T @ U - np.eye(n)) < 100 * EPS * m
    assert np.linalg.norm(V.T @ V - np.eye(n)) < 100 * EPS * m
    assert np.linalg.norm(A - U @ B @ V.T) / np.linalg.norm(A) < 100 * EPS * m


def test_bidiag_zero_column_is_guarded(D):
    # the reference divides 0/0 here (SURVEY.md 8a2 "no guard for zero norm"); we define H = I
    A = util.rand_matrix(20, 20)
    A[:, 0] = 0.0
    Ag, al, be = D.bidiag_par(A)
    assert not np.isnan(Ag).any() and not np.isnan(al).any() and not np.isnan(be).any()
    sv = np.linalg.svd(A, compute_uv=False)
    svb = np.linalg.svd(util.bidiag_dense(al, be), compute_uv=False)
    assert np.abs(sv - svb).max() <= 100 * EPS * 20 * sv.max()


# ------------------------------------------------------------------ phase 2: dDC singular values
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 17, 64, 200, 257, 512, 1024])
def test_ddc_values_vs_lapack_and_oracle(D, n):
    A = util.rand_matrix(n, n)
    _, al, be = util.oracle_bidiag(A)
    sv = np.linalg.svd(util.bidiag_dense(al, be), compute_uv=False)[::-1]
    sg = D.get_singular_values(al, be)
    assert np.all(np.diff(sg) >= 0)
    assert np.abs(sg - sv).max() <= 10 * EPS * max(n, 8) * sv.max()
    # reference parity inside the reference's band
    bep = np.zeros(n); bep[: n - 1] = be
    so = np.zeros(n)
    util.oracle().orc_ddc_values(n, util.p(al.copy()), util.p(bep), util.p(so))
    assert np.abs(sg - so).max() <= 1e-6 * sv.max()


def test_ddc_values_vs_golden_reference(D):
    g = np.load(os.path.join(GOLD, "ref_ddc_257.npz"))
    sg = D.get_singular_values(g["alpha"], g["beta"])
    assert np.abs(sg - g["sigma"]).max() <= 1e-6 * g["sigma"].max()


def test_ddc_rectangular_bidiagonal(D):
    # N x (N+1): b2[N-1] != 0 — the general form of the reference's contract
    rng = np.random.default_rng(5)
    N = 300
    b1 = rng.uniform(0.5, 2.0, N); b2 = rng.uniform(0.5, 2.0, N)
    sv = np.linalg.svd(util.bidiag_dense(b1, b2, N + 1), compute_uv=False)[::-1]
    sg = D.get_singular_values(b1, b2)
    assert np.abs(sg - sv).max() <= 10 * EPS * N * sv.max()


def test_ddc_deflation_and_clusters(D):
    # glued blocks (tiny couplings -> deflation), graded entries and negative signs
    rng = np.random.default_rng(11)
    N = 256
    b1 = rng.uniform(1.0, 2.0, N) * rng.choice([-1.0, 1.0], N)
    b2 = rng.uniform(1.0, 2.0, N)
    b2[N - 1] = 0.0
    b2[63] = 1e-14; b2[127] = 0.0; b2[191] = 1e-9
    b1[:32] *= 1e-3
    sv = np.linalg.svd(util.bidiag_dense(b1, b2[: N - 1]), compute_uv=False)[::-1]
    sg = D.get_singular_values(b1, b2)
    assert not np.isnan(sg).any()
    assert np.abs(sg - sv).max() <= 10 * EPS * N * sv.max()


# ------------------------------------------------------------------ phase 3: twisted vectors
@pytest.mark.parametrize("n", [2, 3, 8, 64, 200, 512, 1024])
def test_twisted_vectors(D, n):
    A = util.rand_matrix(n, n)
    _, al, be = util.oracle_bidiag(A)
    B = util.bidiag_dense(al, be)
    sv = D.get_singular_values(al, be)
    X, Y = D.singular_vectors(al, be, sv)
    I = np.eye(n)
    c = 100 * EPS * max(n, 8)
    assert np.linalg.norm(X @ X.T - I) <= c and np.linalg.norm(Y @ Y.T - I) <= c
    assert np.linalg.norm(B - Y.T @ np.diag(sv) @ X) / np.linalg.norm(B) <= c
    assert np.allclose(np.linalg.norm(X, axis=1), 1.0, atol=1e-14)        # NormalizeVectors :106-120


def test_twisted_vectors_vs_golden_reference(D):
    # vectors agree with the reference's up to sign for well separated sigma, inside the reference's
    # own band: its smallest sigmas carry relative errors up to ~1e-5 and its
    # y = B x / sigma inherits them, hence 1e-4 on the cosines
    g = np.load(os.path.join(GOLD, "ref_svd_96.npz"))
    al, be, sig = g["alpha"], g["beta"][:95], g["sigma_phase"]
    X, Y = D.singular_vectors(al, be, sig)
    gap = np.minimum(np.diff(sig, prepend=-np.inf), np.diff(sig, append=np.inf))
    sep = gap > 1e-3 * sig.max()
    assert sep.sum() >= 5
    cx = np.abs(np.sum(X * g["X"], axis=1))
    cy = np.abs(np.sum(Y * g["Y"], axis=1))
    assert np.all(cx[sep] >= 1 - 1e-4) and np.all(cy[sep] >= 1 - 1e-4)


def test_twisted_rectangular(D):
    # B is n x (n+1): vectors of length n+1 (CalcRightSingularVectors with m = n+1)
    rng = np.random.default_rng(2)
    n = 200
    a = rng.uniform(0.5, 2.0, n); b = rng.uniform(0.5, 2.0, n)
    B = util.bidiag_dense(a, b, n + 1)
    sv = np.linalg.svd(B, compute_uv=False)[::-1]
    X, Y = D.singular_vectors(a, b, sv, m=n + 1)
    c = 100 * EPS * n
    assert np.linalg.norm(X @ X.T - np.eye(n)) <= c and np.linalg.norm(Y @ Y.T - np.eye(n)) <= c
    assert np.linalg.norm(B - Y.T @ np.diag(sv) @ X) / np.linalg.norm(B) <= c


# ------------------------------------------------------------------ phase 4: back-transform
@pytest.mark.parametrize("shape", [(5, 5), (64, 64), (130, 130), (200, 200), (300, 200), (200, 300), (65, 2)])
def test_backtransform_vs_oracle(D, shape):
    m, n = shape
    mn = min(m, n)
    xl = n if m >= n else m + 1
    rng = np.random.default_rng(7)
    A = util.rand_matrix(m, n)
    Ao, _, _ = util.oracle_bidiag(A)
    X = rng.standard_normal((mn, xl)); Y = rng.standard_normal((mn, mn))
    U, V = D.backtransform(Ao, X, Y)
    AT = np.asfortranarray(Ao.T)
    Xc = np.ascontiguousarray(X); Yc = np.ascontiguousarray(Y)
    for i in range(mn):
        u = np.zeros(m); v = np.zeros(n)
        util.oracle().orc_apply_left(m, n, i, util.p(Ao), util.p(Yc), util.p(u))
        util.oracle().orc_apply_right(m, n, i, util.p(AT), util.p(Xc), util.p(v))
        assert np.abs(U[:, i] - u).max() <= 1e-12 * max(1.0, np.abs(u).max())
        assert np.abs(V[:, i] - v).max() <= 1e-12 * max(1.0, np.abs(v).max())


@pytest.mark.parametrize("shape", [(40, 40), (150, 90), (90, 150), (300, 300)])
def test_form_u_v_vs_oracle(D, shape):
    # form_u_par / form_v_par (bidiag_par.c:877-988) on the same WY kernels
    m, n = shape
    A = util.rand_matrix(m, n, 1.0, 2.0, 4)
    Am, al, be = util.oracle_bidiag(A)
    U, V = D.form_q(Am)
    U_o = np.zeros((m, m), order="F"); V_o = np.zeros((n, n), order="F")
    util.oracle().orc_form_u(m, n, util.p(Am), util.p(U_o)); util.oracle().orc_form_v(m, n, util.p(Am), util.p(V_o))
    assert np.abs(U - U_o).max() <= 1e-12 and np.abs(V - V_o).max() <= 1e-12
    assert np.linalg.norm(U.T @ U - np.eye(m)) <= 100 * EPS * m


def test_multU_multV_reference_signatures(D):
    # one vector at a time, with the reference's calling convention (multV takes the transpose)
    L = D.lib()
    n = 40
    g = np.load(os.path.join(GOLD, "ref_svd_48.npz"))
    n = 48
    A_mod = np.asfortranarray(g["A_mod"]); AT = np.asfortranarray(A_mod.T)
    X = np.ascontiguousarray(g["X"]); Y = np.ascontiguousarray(g["Y"])
    for i in (0, 17, 47):
        u = np.zeros(n); v = np.zeros(n)
        L.multU(n, n, i, util.p(A_mod), util.p(Y), util.p(u))
        L.multV(n, n, i, util.p(AT), util.p(X), util.p(v))
        assert np.abs(u - g["U"][:, i]).max() <= 1e-12
        assert np.abs(v - g["V"][:, i]).max() <= 1e-12


# ------------------------------------------------------------------ the whole path
@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (3, 3), (5, 5), (64, 64), (100, 100), (256, 256), (512, 512),
                                   (1024, 1024), (700, 500), (500, 700), (1025, 33), (33, 1025), (300, 40000)])
def test_svd_gpu_vs_lapack(D, shape):
    m, n = shape
    A = util.rand_matrix(m, n)                                 # test-whole-svd.c recipe
    sigma, U, V, A_mod = D.svd_gpu(A)
    assert U.shape == (m, min(m, n)) and V.shape == (n, min(m, n))
    check_lapack_bounds(A, sigma, U, V)


@pytest.mark.parametrize("n", [48, 96])
def test_svd_gpu_vs_golden_reference(D, n):
    g = np.load(os.path.join(GOLD, f"ref_svd_{n}.npz"))
    sigma, U, V, A_mod = D.svd_gpu(g["A"])
    assert np.abs(A_mod - g["A_mod"]).max() <= 1e-9            # A is overwritten with the reflectors
    assert np.abs(sigma - g["sigma"]).max() <= 1e-6 * g["sigma"].max()
    gap = np.minimum(np.diff(g["sigma"], prepend=-np.inf), np.diff(g["sigma"], append=np.inf))
    sep = gap > 1e-3 * g["sigma"].max()
    cu = np.abs(np.sum(U * g["U"], axis=0)); cv = np.abs(np.sum(V * g["V"], axis=0))
    assert np.all(cu[sep] >= 1 - 1e-4) and np.all(cv[sep] >= 1 - 1e-4)
    # U(:,i), V(:,i) pair with sigma[i]: the sign of u_i v_i^T is fixed even if each flips
    for i in np.nonzero(sep)[0][:10]:
        assert np.sign(U[:, i] @ g["U"][:, i]) == np.sign(V[:, i] @ g["V"][:, i])


def test_svd_gpu_vs_oracle_512(D):
    # BASELINE.json configs[0]: 512 x 512, the reference's CPU-runnable case
    A = util.rand_matrix(512, 512)
    sigma, U, V, A_mod = D.svd_gpu(A)
    s_o, U_o, V_o, A_o = util.oracle_svd(A)
    assert np.abs(A_mod - A_o).max() <= 1e-9
    assert np.abs(sigma - s_o).max() <= 1e-6 * s_o.max()
    check_lapack_bounds(A, sigma, U, V)
    # we are strictly more accurate than the reference on its own input
    sv = np.linalg.svd(A, compute_uv=False)[::-1]
    assert np.abs(sigma - sv).max() < 1e-3 * np.abs(s_o - sv).max()


# ------------------------------------------------------------------ the FP64 DMMA GEMM building block
@pytest.mark.parametrize("case", [(0, 0, 130, 70, 50), (1, 0, 64, 200, 333), (0, 1, 257, 129, 64), (1, 1, 65, 66, 67),
                                  (0, 1, 1000, 900, 64), (0, 0, 700, 513, 128), (0, 1, 256, 256, 16),
                                  (0, 0, 1531, 777, 100), (0, 1, 4100, 300, 64)])
@pytest.mark.parametrize("alpha_beta", [(-1.0, 1.0), (1.0, 1.0), (0.5, 0.0), (0.5, 2.0)])
@pytest.mark.parametrize("ws", ["0", "1"])
def test_dgemm_vs_numpy(D, case, alpha_beta, ws, monkeypatch):
    # (transA, transB, M, N, K).  ws = 1: updates (alpha = +-1, beta != 0) and pure products with M > 64
    # go to the persistent warp-specialised kernel (dgemm_ws.cu), everything else - and everything
    # with ws = 0 - to the one-tile-per-CTA kernel (dgemm_dmma.cu)
    monkeypatch.setenv("SVD_GPU_GEMM_WS", ws)
    ta, tb, M, N, K = case
    alpha, beta = alpha_beta
    L = D.lib()
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)); B = rng.standard_normal((K, N)); C = rng.standard_normal((M, N))
    As = np.asfortranarray(A.T if ta else A); Bs = np.asfortranarray(B.T if tb else B); Cs = np.asfortranarray(C)
    bufs = []
    def put(a):
        d = L.svdgpu_malloc(a.nbytes); bufs.append(d)
        L.svdgpu_h2d(d, util.p(a), a.nbytes, None)
        return d
    try:
        dA, dB, dC = put(As), put(Bs), put(Cs)
        L.svdgpu_dgemm(ta, tb, M, N, K, alpha, dA, As.shape[0], dB, Bs.shape[0], beta, dC, M, None)
        out = np.empty((M, N), order="F")
        L.svdgpu_d2h(util.p(out), dC, out.nbytes, None)
        L.svdgpu_stream_sync(None)
    finally:
        for d in bufs:
            L.svdgpu_free(d)
    ref = beta * C + alpha * (A @ B)
    assert np.abs(out - ref).max() <= 50 * EPS * K * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("shape", [(1500, 1300), (700, 900), (5000, 257), (2048, 2048)])
def test_svd_gpu_gemm_kernels_agree(D, shape, monkeypatch):
    # the whole path (panel updates of the bidiagonalization, split-K products and updates of the
    # back-transform, the QR-first route) with either GEMM kernel: same answer up to summation order
    m, n = shape
    A = util.rand_matrix(m, n)
    out = {}
    for ws in ("0", "1"):
        monkeypatch.setenv("SVD_GPU_GEMM_WS", ws)
        out[ws] = D.svd_gpu(A)
        check_lapack_bounds(A, *out[ws][:3])
    s0, U0, V0, _ = out["0"]; s1, U1, V1, _ = out["1"]
    assert np.abs(s0 - s1).max() <= 50 * EPS * max(m, n) * s0.max()
    # vectors of well separated singular values agree up to sign
    gap = np.minimum(np.diff(s0, prepend=-np.inf), np.diff(s0, append=np.inf))
    sep = gap > 1e-3 * s0.max()
    if sep.any():
        assert np.all(np.abs(np.sum(U0[:, sep] * U1[:, sep], axis=0)) >= 1 - 1e-8)
        assert np.all(np.abs(np.sum(V0[:, sep] * V1[:, sep], axis=0)) >= 1 - 1e-8)


@pytest.mark.parametrize("shape,world", [((900, 900), 3), ((1200, 700), 2), ((2304, 2304), 4)])
def test_sharded_entry_points_on_one_gpu(D, shape, world):
    # SURVEY 8e on a single device: svd_gpu_values_dev (what rank 0 runs), then svd_gpu_vectors_dev once per
    # "rank" for its block of singular values (ddc_svd_b200/sharding.py:shard_range) - the concatenated blocks
    # must be the SVD, and the polished singular values must come back block by block
    from ddc_svd_b200.sharding import shard_range
    m, n = shape
    mn = min(m, n)
    L = D.lib()
    A = util.rand_matrix(m, n)
    Af = np.asfortranarray(A)
    bufs = []

    def dev(nbytes):
        d = L.svdgpu_malloc(nbytes); bufs.append(d)
        L.svdgpu_memset(d, 0, nbytes, None)
        return d
    try:
        dA = dev(Af.nbytes)
        L.svdgpu_h2d(dA, util.p(Af), Af.nbytes, None)
        dal, dbe, dsg = dev(8 * mn), dev(8 * (mn + 1)), dev(8 * mn)
        L.svd_gpu_values_dev(m, n, dA, m, dal, dbe, dsg, None)
        U = np.zeros((m, mn), order="F"); V = np.zeros((n, mn), order="F"); sig = np.zeros(mn)
        for rank in range(world):
            blk, i0, ns = shard_range(mn, world, rank)
            if ns == 0:
                continue
            dU, dV, dso = dev(8 * m * blk), dev(8 * n * blk), dev(8 * blk)
            L.svd_gpu_vectors_dev(m, n, dA, m, dal, dbe, dsg, i0, ns, dU, m, dV, n, dso, None)
            Ub = np.zeros((m, ns), order="F"); Vb = np.zeros((n, ns), order="F"); sb = np.zeros(ns)
            L.svdgpu_d2h(util.p(Ub), dU, Ub.nbytes, None)
            L.svdgpu_d2h(util.p(Vb), dV, Vb.nbytes, None)
            L.svdgpu_d2h(util.p(sb), dso, sb.nbytes, None)
            L.svdgpu_stream_sync(None)
            U[:, i0:i0 + ns] = Ub; V[:, i0:i0 + ns] = Vb; sig[i0:i0 + ns] = sb
    finally:
        for d in bufs:
            L.svdgpu_free(d)
    check_lapack_bounds(A, sig, U, V)


# ------------------------------------------------------------------ QR first (m >> n)
@pytest.mark.parametrize("shape", [(5, 2), (300, 100), (1025, 33), (5000, 257), (70000, 130), (4097, 512)])
def test_qr_tall_vs_lapack(D, shape):
    m, n = shape
    A = util.rand_matrix(m, n)
    A_qr, R, Q1 = D.qr_tall(A)
    e = EPS * m
    assert np.all(np.tril(R, -1) == 0.0)
    assert np.array_equal(np.triu(A_qr[:n], 1), np.triu(R, 1))         # R's strict upper triangle stays in place
    assert np.linalg.norm(Q1.T @ Q1 - np.eye(n)) <= 100 * e
    assert np.linalg.norm(Q1 @ R - A) / np.linalg.norm(A) <= 100 * e
    # R is unique up to row signs: same as LAPACK's
    R_l = np.linalg.qr(A, mode="r")
    sg = np.sign(np.diag(R)) * np.sign(np.diag(R_l))
    assert np.abs(R - sg[:, None] * R_l).max() <= 1000 * e * np.abs(R_l).max()
    # unit-norm reflectors from the diagonal down (the library's convention, bidiag_par.c:92-131)
    for j in (0, n // 2, n - 1):
        assert abs(np.linalg.norm(A_qr[j:, j]) - 1.0) <= 1e-13


@pytest.mark.parametrize("shape", [(1025, 33), (4000, 300), (20000, 512), (9001, 1000), (40, 2), (7, 3)])
def test_svd_gpu_qr_first_route(D, shape):
    m, n = shape
    A = util.rand_matrix(m, n)
    D.set_option("qr_first", 0)
    try:
        s0, U0, V0, _ = D.svd_gpu(A)
    finally:
        D.set_option("qr_first", 1)
    s1, U1, V1, _ = D.svd_gpu(A)
    check_lapack_bounds(A, s1, U1, V1)
    check_lapack_bounds(A, s0, U0, V0)
    assert np.abs(s1 - s0).max() <= 10 * EPS * m * s0.max()
    s2, _, _, _ = D.svd_gpu(A, vectors=False)
    assert np.abs(s2 - s1).max() <= 10 * EPS * m * s1.max()


def test_svd_gpu_values_only(D):
    A = util.rand_matrix(300, 300)
    s1, _, _, _ = D.svd_gpu(A, vectors=False)
    sv = np.linalg.svd(A, compute_uv=False)[::-1]
    assert np.abs(s1 - sv).max() <= 10 * EPS * 300 * sv.max()


def test_svd_gpu_only_first_mn_columns_written(D):
    # svd_gpu.c:118-121: the caller hands m x m / n x n buffers, only the first min(m,n) columns change
    m, n = 60, 40
    A = np.array(util.rand_matrix(m, n), order="F")
    U = np.full((m, m), 7.0, order="F"); V = np.full((n, n), 7.0, order="F"); s = np.zeros(n)
    D.lib().svd_gpu(m, n, util.p(A), util.p(s), util.p(U), util.p(V))
    assert np.all(U[:, n:] == 7.0)
    assert not np.any(U[:, :n] == 7.0)


def test_svd_gpu_repeatable(D):
    A = util.rand_matrix(200, 200)
    r1 = D.svd_gpu(A); r2 = D.svd_gpu(A)
    for a, b in zip(r1, r2):
        assert np.array_equal(a, b)                            # deterministic reductions, no atomics


def test_svd_gpu_scaling_and_structured_inputs(D):
    rng = np.random.default_rng(0)
    n = 200
    for A in (1e-150 * util.rand_matrix(n, n), 1e150 * util.rand_matrix(n, n),
              np.diag(np.arange(1.0, n + 1)), np.triu(rng.standard_normal((n, n))),
              rng.standard_normal((n, 3)) @ rng.standard_normal((3, n)) + 1e-8 * rng.standard_normal((n, n))):
        sigma, U, V, _ = D.svd_gpu(A)
        assert not (np.isnan(sigma).any() or np.isnan(U).any() or np.isnan(V).any())
        sv = np.linalg.svd(A, compute_uv=False)[::-1]
        assert np.abs(sigma - sv).max() <= 10 * EPS * n * sv.max()
        assert np.linalg.norm(A - (U * sigma) @ V.T) <= 100 * EPS * n * np.linalg.norm(A)


def test_svd_gpu_graded_spectrum_keeps_U_orthogonal(D):
    # singular values spread over 12 decades: y = B x / sigma (parallel-twisted.c:545-549) would lose the
    # orthogonality of U like eps * sigma_max / sigma_i; the left vectors come from their own twisted
    # factorization of B B^T instead
    rng = np.random.default_rng(0)
    n = 300
    Q1, _ = np.linalg.qr(rng.standard_normal((n, n))); Q2, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sv = np.logspace(0, -12, n)
    A = (Q1 * sv) @ Q2.T
    sigma, U, V, _ = D.svd_gpu(A)
    c = 100 * EPS * n
    assert np.abs(sigma - sv[::-1]).max() <= 10 * EPS * n
    assert np.linalg.norm(U.T @ U - np.eye(n)) <= c and np.linalg.norm(V.T @ V - np.eye(n)) <= c
    assert np.linalg.norm(A - (U * sigma) @ V.T) / np.linalg.norm(A) <= c


@pytest.mark.parametrize("shape", [(200, 300), (300, 200), (120, 900), (900, 120)])
def test_svd_gpu_graded_spectrum_rectangular(D, shape):
    # same, for tall and wide inputs: a wide matrix is solved as the SVD of its transpose, so its U
    # (the transposed problem's V) is as orthogonal as any other; the old wide route (option
    # "wide_transpose" = 0) takes y = B x / sigma and is only held to the residual bound
    m, n = shape
    mn = min(m, n)
    rng = np.random.default_rng(1)
    Q1, _ = np.linalg.qr(rng.standard_normal((m, mn))); Q2, _ = np.linalg.qr(rng.standard_normal((n, mn)))
    sv = np.logspace(0, -10, mn)
    A = (Q1 * sv) @ Q2.T
    sigma, U, V, _ = D.svd_gpu(A)
    c = 100 * EPS * max(m, n)
    assert np.abs(sigma - sv[::-1]).max() <= 10 * EPS * max(m, n)
    assert np.linalg.norm(U.T @ U - np.eye(mn)) <= c and np.linalg.norm(V.T @ V - np.eye(mn)) <= c
    assert np.linalg.norm(A - (U * sigma) @ V.T) / np.linalg.norm(A) <= c
    if m < n:
        D.set_option("wide_transpose", 0)
        try:
            s0, U0, V0, A0 = D.svd_gpu(A)
        finally:
            D.set_option("wide_transpose", 1)
        assert np.abs(s0 - sigma).max() <= 10 * EPS * max(m, n)
        assert np.linalg.norm(A - (U0 * s0) @ V0.T) / np.linalg.norm(A) <= c


def test_svd_gpu_wide_direct_route_keeps_reference_reflectors(D):
    # option "wide_transpose" = 0: a wide input is bidiagonalized as it is and A leaves as the
    # reference's reflector storage (bidiag.c:33-186); the default route returns the transposed
    # problem's storage instead
    A = util.rand_matrix(90, 130)
    D.set_option("wide_transpose", 0)
    try:
        s0, U0, V0, A0 = D.svd_gpu(A)
    finally:
        D.set_option("wide_transpose", 1)
    Ao, _, _ = util.oracle_bidiag(A)
    assert np.abs(A0 - Ao).max() <= 1e-9
    check_lapack_bounds(A, s0, U0, V0)
    s1, U1, V1, A1 = D.svd_gpu(A)
    check_lapack_bounds(A, s1, U1, V1)
    At, _, _ = util.oracle_bidiag(np.ascontiguousarray(A.T))
    assert np.abs(A1 - At.T).max() <= 1e-9


# ------------------------------------------------------------------ benchmark sizes: properties
def test_svd_gpu_4096_properties(D):
    # BASELINE.json configs[1]; LAPACK for sigma, size-independent properties for the vectors
    n = 4096
    A = util.rand_matrix(n, n)
    sigma, U, V, A_mod = D.svd_gpu(A)
    sv = np.linalg.svd(A, compute_uv=False)[::-1]
    e = EPS * n
    assert np.all(np.diff(sigma) >= 0)
    assert np.abs(sigma - sv).max() / sv.max() <= 10 * e
    assert np.linalg.norm(U.T @ U - np.eye(n)) <= 100 * e
    assert np.linalg.norm(V.T @ V - np.eye(n)) <= 100 * e
    assert np.linalg.norm(A - (U * sigma) @ V.T) / np.linalg.norm(A) <= 100 * e
    # checksum of checksums: ||A||_F^2 = sum sigma^2
    assert abs(np.sum(sigma ** 2) - np.sum(A ** 2)) <= 100 * e * np.sum(A ** 2)


@pytest.mark.parametrize("shape", [(1500, 1400), (2500, 2500), (3000, 1000), (2000, 3000), (4096, 4096), (4100, 1800)])
@pytest.mark.parametrize("tail", ["0", "2"])
def test_bidiag_persistent_panel_kernel(D, shape, tail, monkeypatch):
    # bidiag_panel.cuh: all the steps of a panel in ONE cooperative launch (pass -> grid barrier -> finish -> grid
    # barrier), the TMA producer streaming across the barriers.  Same arithmetic as the two-kernel path: the results
    # must agree with it to rounding, and with the oracle where the CPU restatement finishes in seconds.
    m, n = shape
    A = util.rand_matrix(m, n, 1.0, 2.0, 4)
    monkeypatch.setenv("SVD_GPU_TAIL", tail)
    monkeypatch.setenv("SVD_GPU_PPK", "1")
    Ag, ag, bg = D.bidiag_par(A)
    monkeypatch.setenv("SVD_GPU_PPK", "0")
    A0, a0, b0 = D.bidiag_par(A)
    assert not np.isnan(Ag).any()
    # (the finish splits its 2k-term corrections 16 ways instead of 32: last-bit differences, amplified by the
    #  factorization like any rounding error)
    assert np.abs(Ag - A0).max() <= 2e-10 and np.abs(ag - a0).max() <= 2e-10 * np.abs(a0).max()
    assert np.abs(bg - b0).max() <= 2e-10 * max(1.0, np.abs(b0).max())
    fro2 = float(np.sum(A * A))
    assert abs(float(np.sum(ag * ag) + np.sum(bg * bg)) - fro2) <= 1e-12 * fro2      # ||B||_F = ||A||_F
    if m * n <= 2500 * 2500:
        Ao, ao, bo = util.oracle_bidiag(A)
        assert np.abs(Ag - Ao).max() <= 1e-9 and np.abs(ag - ao).max() <= 1e-9 and np.abs(bg - bo).max() <= 1e-9
