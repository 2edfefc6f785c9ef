"""Small end-to-end calls for compute-sanitizer (not a pytest file):
    compute-sanitizer --tool memcheck python tests/sanitize_target.py
Covers the fused pass (clusters of 1 and 2 via SVD_GPU_FUSED_CS), the split passes, QR first, the
wide-as-transpose route, values only, the phase entry points, and - second loop - the kernels behind the
SVD_GPU_TAIL / SVD_GPU_GEMM_WS switches (on-chip tail on and off, persistent and one-tile-per-CTA GEMM), and
- round 2 - the on-device checker, the progressive panel set-up, a group of one rank, the tcgen05 update, the persistent per-panel kernel.
(racecheck does not follow cross-CTA traffic through global memory: the tail kernel's exchange protocol is
argued in bidiag_tail.cuh and exercised by test_bidiag_on_chip_tail.)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import util
import ddc_svd_b200 as D

EPS = 2.220446049250313e-16
for (m, n) in [(300, 300), (700, 120), (120, 700), (513, 257), (64, 64), (1, 1), (2, 5)]:
    A = util.rand_matrix(m, n)
    s, U, V, _ = D.svd_gpu(A)
    met = util.svd_metrics(A, s, U, V)
    assert met["resid"] <= 100 * EPS * max(m, n), (m, n, met)
    s2, _, _, _ = D.svd_gpu(A, vectors=False)
    assert np.abs(s2 - s).max() <= 10 * EPS * max(m, n) * max(s.max(), 1.0)
    print("svd_gpu", (m, n), "ok", flush=True)
for tail, ws in (("0", "1"), ("1", "0"), ("0", "0")):
    os.environ["SVD_GPU_TAIL"] = tail; os.environ["SVD_GPU_GEMM_WS"] = ws
    for (m, n) in [(300, 300), (700, 120), (257, 513)]:
        A = util.rand_matrix(m, n)
        s, U, V, _ = D.svd_gpu(A)
        met = util.svd_metrics(A, s, U, V)
        assert met["resid"] <= 100 * EPS * max(m, n), (tail, ws, m, n, met)
    print("svd_gpu with SVD_GPU_TAIL=%s SVD_GPU_GEMM_WS=%s ok" % (tail, ws), flush=True)
del os.environ["SVD_GPU_TAIL"], os.environ["SVD_GPU_GEMM_WS"]
A = util.rand_matrix(400, 350)
Am, al, be = D.bidiag_par(A)
Ao, ao, bo = util.oracle_bidiag(A)
assert np.abs(Am - Ao).max() <= 1e-9
Aq, R, Q1 = D.qr_tall(util.rand_matrix(900, 130))
print("phases ok", flush=True)
# round 2: the on-device checker, the progressive panel set-up (the path rank 0 of a multi-GPU group takes), a
# group of one rank, the tcgen05 (int8 slices in TMEM) update
A = util.rand_matrix(400, 300)
s, U, V, _ = D.svd_gpu(A)
chk = D.check(A, s, U, V)
assert chk["resid"] <= 100 * EPS * 400 and chk["ascending"], chk
D.set_option("wy_overlap", 1)
s2, U2, V2, _ = D.svd_gpu(util.rand_matrix(700, 650))
D.set_option("wy_overlap", 0)
g = D.Group.local(1)
s3, U3, V3, _ = g.svd(util.rand_matrix(300, 300))
g.destroy()
import ctypes
L = D.lib()
M, N, K = 300, 100, 128
rng = np.random.default_rng(0)
Am = np.asfortranarray(rng.standard_normal((M, K))); Bm = np.asfortranarray(rng.standard_normal((K, N))); Cm = np.asfortranarray(rng.standard_normal((M, N)))
bufs = [L.svdgpu_malloc(x.nbytes) for x in (Am, Bm, Cm)] + [L.svdgpu_malloc(L.svdgpu_ozaki_workspace(M, N))]
for d, x in zip(bufs, (Am, Bm, Cm)):
    L.svdgpu_h2d(d, util.p(x), x.nbytes, None)
L.svdgpu_ozaki_update(M, N, -1.0, bufs[0], M, bufs[1], K, bufs[2], M, bufs[3], None)
out = np.empty((M, N), order="F")
L.svdgpu_d2h(util.p(out), bufs[2], out.nbytes, None); L.svdgpu_stream_sync(None)
assert np.abs(out - (Cm - Am @ Bm)).max() <= 1e-11
for d in bufs:
    L.svdgpu_free(d)
# the persistent per-panel kernel (opt-in) and the 128-row finish kernel, streaming path to the end
os.environ["SVD_GPU_PPK"] = "1"; os.environ["SVD_GPU_TAIL"] = "0"; os.environ["SVD_GPU_XW"] = "2"
A = util.rand_matrix(1200, 1000)
Am, al, be = D.bidiag_par(A)
fro2 = float(np.sum(A * A))
assert abs(float(np.sum(al * al) + np.sum(be * be)) - fro2) <= 1e-12 * fro2
del os.environ["SVD_GPU_PPK"], os.environ["SVD_GPU_TAIL"], os.environ["SVD_GPU_XW"]
print("round-2 paths ok", flush=True)
