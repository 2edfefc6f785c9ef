"""CPU tests: the oracle restatement against (a) the committed golden vectors produced by the
unmodified reference and (b) the reference build itself when oracle/_ref is present."""
import glob
import os

import numpy as np
import pytest

import util
from util import p

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_input_recipe_matches_golden():
    # the fixtures store the inputs; regenerating them pins the rand() recipe of test-whole-svd.c:18-24
    g = np.load(os.path.join(GOLD, "ref_svd_48.npz"))
    assert np.array_equal(util.rand_matrix(48, 48, 1.0, 4.0, 1), g["A"])
    g = np.load(os.path.join(GOLD, "ref_bidiag_90x60.npz"))
    assert np.array_equal(util.rand_matrix(90, 60, 1.0, 2.0, 4), g["A"])


@pytest.mark.parametrize("name", ["ref_bidiag_80x80", "ref_bidiag_90x60", "ref_bidiag_60x90"])
def test_oracle_bidiag_vs_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    A_mod, alpha, beta = util.oracle_bidiag(g["A"])
    # same arithmetic in the same order as bidiag.c: bit-for-bit
    assert np.array_equal(A_mod, g["A_mod"])
    assert np.array_equal(alpha, g["alpha"])
    assert np.array_equal(beta, g["beta"])


@pytest.mark.parametrize("n", [48, 96])
def test_oracle_phases_vs_golden(n):
    g = np.load(os.path.join(GOLD, f"ref_svd_{n}.npz"))
    alpha, beta = g["alpha"].copy(), g["beta"].copy()
    sig = np.zeros(n)
    util.oracle().orc_ddc_values(n, p(alpha), p(beta), p(sig))
    assert np.array_equal(sig, g["sigma_phase"])
    X = np.zeros(n * n)
    Y = np.zeros(n * n)
    util.oracle().orc_right_vectors(n, n, p(alpha), p(beta), p(sig), p(X))
    util.oracle().orc_left_vectors(n, n, p(alpha), p(beta), p(sig), p(X), p(Y))
    assert np.array_equal(X.reshape(n, n), g["X"])
    assert np.array_equal(Y.reshape(n, n), g["Y"])


@pytest.mark.parametrize("n", [48, 96])
def test_oracle_svd_vs_golden(n):
    g = np.load(os.path.join(GOLD, f"ref_svd_{n}.npz"))
    sigma, U, V, A_mod = util.oracle_svd(g["A"])
    assert np.array_equal(sigma, g["sigma"])
    assert np.array_equal(A_mod, g["A_mod"])
    assert np.array_equal(U, g["U"][:, :n])
    assert np.array_equal(V, g["V"][:, :n])
    assert np.all(np.diff(sigma) > 0)           # ascending (Calculations-Parallel.c:56)


def test_oracle_ddc_257_vs_golden():
    g = np.load(os.path.join(GOLD, "ref_ddc_257.npz"))
    sig = np.zeros(257)
    util.oracle().orc_ddc_values(257, p(g["alpha"].copy()), p(g["beta"].copy()), p(sig))
    assert np.array_equal(sig, g["sigma"])


def test_golden_reference_accuracy_band():
    # documents the reference's own error band: sigma is only good to ~1e-7*sigma_max
    g = np.load(os.path.join(GOLD, "ref_ddc_257.npz"))
    sv = np.linalg.svd(util.bidiag_dense(g["alpha"], g["beta"][:256]), compute_uv=False)[::-1]
    err = np.abs(g["sigma"] - sv).max() / sv.max()
    assert 1e-12 < err < 1e-6


@pytest.mark.skipif(util.reference() is None, reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("n", [8, 33, 130])
def test_oracle_vs_reference_build(n):
    ref = util.reference()
    A = util.rand_matrix(n, n, 1.0, 4.0, 1)
    s_o, U_o, V_o, A_o = util.oracle_svd(A)
    r = util.reference_svd(ref, A)
    assert np.array_equal(s_o, r["sigma"]) and np.array_equal(A_o, r["A_mod"])
    assert np.array_equal(U_o, r["U"]) and np.array_equal(V_o, r["V"])


def test_oracle_rectangular_fixes():
    # the two documented deviations: results for m != n are a valid SVD up to the reference
    # algorithm's own accuracy (the reference itself is wrong there)
    for (m, n) in [(70, 50), (50, 70)]:
        A = util.rand_matrix(m, n, 1.0, 4.0, 1)
        s, U, V, _ = util.oracle_svd(A)
        met = util.svd_metrics(A, s, U, V)
        assert met["resid"] < 1e-3 and met["sigma_abs_over_max"] < 1e-5


@pytest.mark.skipif(util.reference() is None, reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("shape", [(20, 20), (33, 21), (21, 33)])
def test_oracle_form_uv_vs_reference_build(shape):
    # explicit Q formation (bidiag.c:252-359): bit-for-bit, and A = U B V^T
    m, n = shape
    ref = util.reference()
    A = util.rand_matrix(m, n, 1.0, 2.0, 4)
    Am, al, be = util.oracle_bidiag(A)
    U_o = np.zeros((m, m), order="F"); V_o = np.zeros((n, n), order="F")
    U_r = np.zeros((m, m), order="F"); V_r = np.zeros((n, n), order="F")
    util.oracle().orc_form_u(m, n, p(Am), p(U_o)); util.oracle().orc_form_v(m, n, p(Am), p(V_o))
    ref.form_u(m, n, p(Am), p(U_r)); ref.form_v(m, n, p(Am), p(V_r))
    assert np.array_equal(U_o, U_r) and np.array_equal(V_o, V_r)
    B = np.zeros((m, n)); k = min(m, n)
    B[np.arange(k), np.arange(k)] = al
    B[np.arange(len(be)), np.arange(len(be)) + 1] = be
    assert np.abs(U_o @ B @ V_o.T - A).max() < 1e-12
