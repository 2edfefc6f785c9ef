"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libddcref.so).

Run in the build container only (needs /root/reference, from which oracle/Makefile compiles
the reference's host C in place):   python tests/golden/make_golden.py
The reference ships no golden vectors; these pin its behaviour so that the
oracle restatement and the CUDA path can be checked on machines without /root/reference.
Inputs follow the reference drivers' recipes and are stored alongside the outputs.
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import util  # noqa: E402

P = util.P
p = util.p


def main():
    ref = util.reference()
    if ref is None:
        raise SystemExit("oracle/_ref/libddcref.so is not available (needs /root/reference)")
    # 1) whole svd_gpu path, test-whole-svd.c recipe: uniform [1,4), glibc default seed (1)
    for n in (48, 96):
        A = util.rand_matrix(n, n, 1.0, 4.0, 1)
        r = util.reference_svd(ref, A)      # svd_gpu.c:100-121 order, beta zero-padded
        np.savez_compressed(os.path.join(HERE, f"ref_svd_{n}.npz"), A=A, A_mod=r["A_mod"], sigma=r["sigma"],
                            U=r["U"], V=r["V"], alpha=r["alpha"], beta=r["beta"], sigma_phase=r["sigma"],
                            X=r["X"], Y=r["Y"])
    # 2) bidiagonalization alone, bidiag_dr.c recipe: uniform [1,2), srand(4); square only for the
    #    whole path, but bidiag_seq itself is valid for rectangular inputs
    for (m, n) in ((80, 80), (90, 60), (60, 90)):
        A = util.rand_matrix(m, n, 1.0, 2.0, 4)
        Ab = A.copy(order="F"); mn = min(m, n)
        alpha = np.zeros(mn); beta = np.zeros(mn + 1)
        ref.bidiag_seq(m, n, p(Ab), p(alpha), p(beta))
        lb = n - 1 if m >= n else m
        np.savez_compressed(os.path.join(HERE, f"ref_bidiag_{m}x{n}.npz"), A=A, A_mod=Ab, alpha=alpha,
                            beta=beta[:lb])
    # 3) singular values of a larger case (where the reference's own accuracy has degraded)
    n = 257
    A = util.rand_matrix(n, n, 1.0, 4.0, 1)
    Ab = A.copy(order="F"); alpha = np.zeros(n); beta = np.zeros(n)
    ref.bidiag_seq(n, n, p(Ab), p(alpha), p(beta)); beta[n - 1] = 0.0
    sig = np.zeros(n)
    ref.GetSingularValues_Parallel(n, p(alpha), p(beta), p(sig))
    np.savez_compressed(os.path.join(HERE, "ref_ddc_257.npz"), alpha=alpha, beta=beta, sigma=sig)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
