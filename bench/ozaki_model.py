"""Feasibility model (numpy, CPU) of an int8 slice scheme for the FP64 GEMMs of the back-transform.

Question for the next round: how many 7-bit slices does `W = V^T C` / `C -= (VT) W` need on int8 tensor cores
(tcgen05 kind::i8, int32 accumulation) to match the FP64 DMMA result on THIS path's operands (unit-norm
reflector panels, orthonormal-column C)?  Scheme: every row of A (column of B) is scaled by a power of two so
that max |entry| < 1, then cut into signed slices of `bits` bits; slice products with the same total shift
are summed exactly in int32 (K * 2^(2 bits) must stay below 2^31) and combined in FP64.

    python bench/ozaki_model.py [--rows 4096] [--k 128] [--cols 512]
"""
import argparse
import numpy as np


def slices(M, axis, nsl, bits):
    """Power-of-two row/column scaling + signed fixed-point slices (exact: sum_s S_s 2^{-bits (s+1)} == M/scale)."""
    mx = np.abs(M).max(axis=axis, keepdims=True)
    e = np.where(mx > 0, np.ceil(np.log2(np.where(mx > 0, mx, 1.0))) + 1, 0.0)
    scale = 2.0 ** e
    R = M / scale                                      # |R| < 1/2 ... < 1
    out = []
    for s in range(nsl):
        R = R * (2.0 ** bits)
        S = np.round(R)                                # |S| <= 2^(bits-1) after the first slice
        out.append(S.astype(np.int64))
        R = R - S
    return out, scale


def ozaki_gemm(A, B, nsl, bits):
    """A (M x K) @ B (K x N) with nsl slices each; products with total shift > nsl are dropped (triangular)."""
    SA, sa = slices(A, 1, nsl, bits)
    SB, sb = slices(B, 0, nsl, bits)
    K = A.shape[1]
    assert K * (2 ** (bits - 1)) ** 2 * nsl < 2 ** 31, "int32 accumulation would overflow"
    C = np.zeros((A.shape[0], B.shape[1]))
    ngemm = 0
    for tot in range(2 * nsl - 1):
        if tot >= nsl:                                 # below the last kept bit
            continue
        acc = np.zeros((A.shape[0], B.shape[1]), dtype=np.int64)
        for i in range(nsl):
            j = tot - i
            if 0 <= j < nsl:
                acc += SA[i] @ SB[j]
                ngemm += 1
        assert np.abs(acc).max() < 2 ** 31
        C += acc.astype(np.float64) * 2.0 ** (-bits * (tot + 2))
    return C * sa * sb, ngemm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--k", type=int, default=128)
    ap.add_argument("--cols", type=int, default=512)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    # a reflector panel: unit-norm columns, trapezoidal; C: orthonormal columns
    V = np.tril(rng.standard_normal((a.rows, a.k)))
    V /= np.linalg.norm(V, axis=0)
    Cm = np.linalg.qr(rng.standard_normal((a.rows, a.cols)))[0]
    W_ref = (V.T.astype(np.longdouble) @ Cm.astype(np.longdouble)).astype(np.float64)
    W_f64 = V.T @ Cm
    err64 = np.abs(W_f64 - W_ref).max()
    print(f"W = V^T C, K = {a.rows}: FP64 GEMM error vs long double {err64:.2e} (max |W| {np.abs(W_ref).max():.2e})")
    for bits in (7, 6):
        for nsl in (6, 7, 8, 9, 10):
            if a.rows * (2 ** (bits - 1)) ** 2 * nsl >= 2 ** 31:
                continue
            W_o, ng = ozaki_gemm(V.T, Cm, nsl, bits)
            print(f"  bits {bits} slices {nsl:2d}: {ng:3d} int8 GEMMs, error {np.abs(W_o - W_ref).max():.2e}"
                  f"  ({np.abs(W_o - W_ref).max() / max(err64, 1e-300):.1f} x FP64)")
    # the short-K update C -= (VT) W
    VT = V @ np.triu(rng.standard_normal((a.k, a.k)))
    U_ref = (VT.astype(np.longdouble) @ W_ref.astype(np.longdouble)).astype(np.float64)
    err64 = np.abs(VT @ W_ref - U_ref).max()
    print(f"(VT) W, K = {a.k}: FP64 GEMM error {err64:.2e} (max |.| {np.abs(U_ref).max():.2e})")
    for nsl in (7, 8, 9):
        U_o, ng = ozaki_gemm(VT, W_ref, nsl, 7)
        print(f"  bits 7 slices {nsl:2d}: {ng:3d} int8 GEMMs, error {np.abs(U_o - U_ref).max():.2e}"
              f"  ({np.abs(U_o - U_ref).max() / max(err64, 1e-300):.1f} x FP64)")


if __name__ == "__main__":
    main()
