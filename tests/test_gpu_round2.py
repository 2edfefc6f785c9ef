"""GPU tests of what round 2 added around the path: the on-device checker (the reference driver's dormant
check, enabled), the executed drop-in binaries, pageable callers, the A_mod contract of the default routes,
the progressive panel set-up that the multi-GPU path uses on rank 0, thread safety, the BASELINE.json
configs at full size through size-independent properties, and — on boxes with more than one GPU — the
sharded path itself (one process driving N GPUs, and one process per GPU).  Everything goes through the C ABI."""
import ctypes
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
EPS = util.EPS
ROOT = util.ROOT


@pytest.fixture(scope="module")
def D():
    import ddc_svd_b200 as mod
    mod.lib()
    return mod


def ngpus(D):
    return D.lib().svdgpu_device_count()


def bounds_ok(A, sigma, U, V, c_sig=10.0, c_vec=100.0):
    m, n = A.shape
    met = util.svd_metrics(A, sigma, U, V)
    b = EPS * max(m, n)
    assert met["ascending"], met
    assert met["sigma_abs_over_max"] <= c_sig * b, met
    assert met["orthU"] <= c_vec * b and met["orthV"] <= c_vec * b and met["resid"] <= c_vec * b, met
    return met


# ------------------------------------------------------------------ the checker (test-whole-svd.c:81-96, enabled)
@pytest.mark.parametrize("shape", [(1, 1), (7, 7), (300, 200), (200, 300), (1100, 1100), (2500, 130)])
def test_device_checker_matches_numpy(D, shape):
    m, n = shape
    A = util.rand_matrix(m, n)
    s, U, V, _ = D.svd_gpu(A)
    met = util.svd_metrics(A, s, U, V)
    chk = D.check(A, s, U, V)
    assert chk["ascending"]
    # same quantities as numpy computes them, to rounding of the reductions
    for k_np, k_dev in (("orthU", "orthU"), ("orthV", "orthV"), ("resid", "resid")):
        assert abs(chk[k_dev] - met[k_np]) <= 0.2 * met[k_np] + 5 * EPS, (k_np, chk, met)
    assert abs(chk["normA"] - np.linalg.norm(A)) <= 1e-12 * np.linalg.norm(A)
    assert chk["checksum"] <= 100 * EPS * max(m, n)


def test_device_checker_sees_a_wrong_result(D):
    A = util.rand_matrix(400, 300)
    s, U, V, _ = D.svd_gpu(A)
    Ub = U.copy(); Ub[:, 5] = Ub[:, 6]                     # two equal columns
    assert D.check(A, s, Ub, V)["orthU"] > 0.5
    sb = s.copy(); sb[10] *= 1.0 + 1e-6
    assert D.check(A, sb, U, V)["resid"] > 1e-10
    sd = s[::-1].copy()
    assert not D.check(A, sd, U, V)["ascending"]


def test_device_checker_block_form(D):
    # what one rank of a sharded run can verify on its own: ||A V_b - U_b S_b||_F / ||A||_F and its block's Gram
    m, n = 900, 700
    A = util.rand_matrix(m, n)
    s, U, V, _ = D.svd_gpu(A)
    L = D.lib()
    i0, ns = 200, 170
    Af = np.asfortranarray(A); Ub = np.asfortranarray(U[:, i0:i0 + ns]); Vb = np.asfortranarray(V[:, i0:i0 + ns])
    sb = np.ascontiguousarray(s[i0:i0 + ns])
    bufs = [L.svdgpu_malloc(x.nbytes) for x in (Af, Ub, Vb, sb)]
    try:
        for d, x in zip(bufs, (Af, Ub, Vb, sb)):
            L.svdgpu_h2d(d, util.p(x), x.nbytes, None)
        out = np.zeros(6)
        L.svd_gpu_check_dev(m, n, bufs[0], m, bufs[3], bufs[1], m, bufs[2], n, ns, util.p(out), None)
    finally:
        for d in bufs:
            L.svdgpu_free(d)
    ref = np.linalg.norm(A @ Vb - Ub * sb) / np.linalg.norm(A)
    assert abs(out[2] - ref) <= 0.2 * ref + 5 * EPS
    assert out[0] <= 100 * EPS * m and out[1] <= 100 * EPS * m and out[5] == 1.0


# ------------------------------------------------------------------ the drop-in client, executed
def _run(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, **kw)


def test_reference_driver_binary_runs_on_the_gpu():
    # build/test-whole-svd is the reference's UNMODIFIED test-whole-svd.c linked against libsvdgpu.so
    # (make dropin, run by __graft_entry__.build() where /root/reference is mounted; the binary travels)
    exe = os.path.join(ROOT, "build", "test-whole-svd")
    assert os.path.exists(exe), "build/test-whole-svd missing: run __graft_entry__.build() where the reference is mounted"
    r = _run([exe, "512", "512"])
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.parametrize("args", [("512", "512"), ("1300", "1300"), ("900", "400"), ("300", "700"), ("3000", "300")])
def test_self_checking_driver(args):
    # tests/dropin_check.c: the same flow (malloc'd buffers, rand() inputs) with the driver's "#if 0" check enabled
    exe = os.path.join(ROOT, "build", "dropin_check")
    assert os.path.exists(exe), "build/dropin_check missing: make dropin"
    r = _run([exe, *args])
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert "OK" in r.stdout and "relative error in Frobenius norm" in r.stdout


# ------------------------------------------------------------------ pageable callers, options, threads
def test_pageable_and_page_locked_callers_agree(D):
    A = util.rand_matrix(1200, 1000)
    try:
        D.set_option("host_register", 0)
        s0, U0, V0, A0 = D.svd_gpu(A)
        D.set_option("host_register", 1)
        s1, U1, V1, A1 = D.svd_gpu(A)
    finally:
        D.set_option("host_register", 0)
    assert np.array_equal(s0, s1) and np.array_equal(U0, U1) and np.array_equal(V0, V1) and np.array_equal(A0, A1)
    bounds_ok(A, s1, U1, V1)


def test_progressive_panel_setup_matches_default(D):
    # "wy_overlap": compact-WY panels are prepared on the side stream WHILE the factorization runs (what rank 0
    # of a multi-GPU group always does).  Reflectors and singular values are untouched (bitwise); the panels'
    # Gram matrices are summed in a batch-dependent split, so U / V agree to rounding, not bitwise
    for shape in [(1500, 1300), (2304, 2304), (5000, 257), (700, 900), (130, 130), (64, 64), (3, 3), (1, 1)]:
        A = util.rand_matrix(*shape)
        s0, U0, V0, A0 = D.svd_gpu(A)
        try:
            D.set_option("wy_overlap", 1)
            s1, U1, V1, A1 = D.svd_gpu(A)
        finally:
            D.set_option("wy_overlap", 0)
        assert np.array_equal(s0, s1) and np.array_equal(A0, A1), shape
        assert np.abs(U0 - U1).max() <= 1e-11 and np.abs(V0 - V1).max() <= 1e-11, shape
        bounds_ok(A, s1, U1, V1)


def test_concurrent_callers_are_serialised_per_device(D):
    # two host threads call svd_gpu() on the same device: the per-device context is locked, both get their SVD
    mats = [util.rand_matrix(700, 600, seed=3), util.rand_matrix(640, 640, seed=5)]
    res = [None, None]

    def work(i):
        for _ in range(3):
            res[i] = D.svd_gpu(mats[i])
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for A, (s, U, V, _) in zip(mats, res):
        bounds_ok(A, s, U, V)


def test_a_mod_contract_of_the_default_routes(D):
    # include/svd_gpu_b200.h: what svd_gpu() leaves in A.  Square / mildly tall: the reference's bidiag_par
    # reflectors (1e-9, bidiag_dr.c:94).  m >= 2.5 n: reflectors of A = QR + R above the diagonal, unless
    # qr_first = 0.  m < n: transpose of the tall problem's storage, unless wide_transpose = 0.
    A = util.rand_matrix(600, 500, 1.0, 2.0, 4)
    _, _, _, Am = D.svd_gpu(A)
    Ao, _, _ = util.oracle_bidiag(A)
    assert np.abs(Am - Ao).max() <= 1e-9
    T = util.rand_matrix(1500, 200, 1.0, 2.0, 4)
    _, _, _, Tm = D.svd_gpu(T)                               # default: QR first
    Aqr, R, _ = D.qr_tall(T)
    assert np.abs(np.triu(Tm[:200], 1) - np.triu(R, 1)).max() <= 1e-9 * np.abs(R).max()
    assert np.abs(np.tril(Tm) - np.tril(Aqr)).max() <= 1e-9
    To, _, _ = util.oracle_bidiag(T)
    assert np.abs(Tm - To).max() > 1e-3                      # ... which is NOT the bidiagonalization's storage
    try:
        D.set_option("qr_first", 0)
        _, _, _, Tm0 = D.svd_gpu(T)
    finally:
        D.set_option("qr_first", 1)
    assert np.abs(Tm0 - To).max() <= 1e-9
    W = util.rand_matrix(200, 320, 1.0, 2.0, 4)
    _, _, _, Wm = D.svd_gpu(W)                               # default: SVD of the transpose
    _, _, _, Wt = D.svd_gpu(np.asfortranarray(W.T))
    assert np.array_equal(Wm, Wt.T)
    try:
        D.set_option("wide_transpose", 0)
        _, _, _, Wm0 = D.svd_gpu(W)
    finally:
        D.set_option("wide_transpose", 1)
    Wo, _, _ = util.oracle_bidiag(W)
    assert np.abs(Wm0 - Wo).max() <= 1e-9


def test_group_of_one_is_svd_gpu(D):
    A = util.rand_matrix(1000, 900)
    g = D.Group.local(1)
    try:
        s1, U1, V1, A1 = g.svd(A)
        ms = g.phase_ms(0)
    finally:
        g.destroy()
    s0, U0, V0, A0 = D.svd_gpu(A)
    assert np.array_equal(s0, s1) and np.array_equal(U0, U1) and np.array_equal(V0, V1) and np.array_equal(A0, A1)
    assert ms[1] > 0 and ms[4] > 0 and ms[6] >= ms[1]


# ------------------------------------------------------------------ BASELINE.json configs at full size
def _device_svd_with_check(D, m, n, vectors=True, seed=1):
    """svd_gpu_dev on a resident random matrix + svd_gpu_check_dev against a resident copy of the input."""
    L = D.lib()
    mn = min(m, n)
    rng = np.random.default_rng(seed)
    A = np.asfortranarray(rng.uniform(1.0, 4.0, size=(n, m)).T)     # column-major m x n
    nbA = A.nbytes
    bufs = []

    def dev(nbytes):
        d = L.svdgpu_malloc(nbytes); bufs.append(d)
        return d
    out = np.zeros(6)
    sig = np.zeros(mn)
    try:
        dA0, dA, dsg = dev(nbA), dev(nbA), dev(8 * mn)
        L.svdgpu_h2d(dA0, util.p(A), nbA, None)
        L.svdgpu_d2d(dA, dA0, nbA, None)
        if vectors:
            dU, dV = dev(8 * m * mn), dev(8 * n * mn)
            L.svd_gpu_dev(m, n, dA, m, dsg, dU, m, dV, n, None)
            L.svd_gpu_check_dev(m, n, dA0, m, dsg, dU, m, dV, n, mn, util.p(out), None)
        else:
            L.svd_gpu_dev(m, n, dA, m, dsg, None, m, None, n, None)
        L.svdgpu_d2h(util.p(sig), dsg, 8 * mn, None)
        L.svdgpu_stream_sync(None)
        normA2 = float(np.einsum("ij,ij->", A, A))
    finally:
        for d in bufs:
            L.svdgpu_free(d)
    return sig, out, normA2, D.last_phase_ms()


def test_config_c4_16384_full_vs_values_only(D):
    # BASELINE.json configs[3]: square 16384^2, singular values only (dDC path) vs full vectors — on the final
    # build, through size-independent properties: the checker's three Frobenius figures at c*eps*n, the checksum
    # of checksums sum sigma^2 = ||A||_F^2, ascending order, and values-only == the unpolished dDC values of the
    # full run to 10 eps n sigma_max
    n = 16384
    sig, out, normA2, ms = _device_svd_with_check(D, n, n, vectors=True)
    b = EPS * n
    assert out[5] == 1.0 and out[0] <= 100 * b and out[1] <= 100 * b and out[2] <= 100 * b, out
    assert out[3] <= 100 * b
    assert abs(np.dot(sig, sig) - normA2) <= 100 * b * normA2
    sig_v, _, _, ms_v = _device_svd_with_check(D, n, n, vectors=False)
    assert np.all(np.diff(sig_v) >= 0)
    assert np.abs(sig_v - sig).max() <= 10 * b * sig.max()
    assert ms_v[3] == 0.0 or ms_v[3] < 1.0                  # no vector phases in a values-only run
    print("C4 16384^2 full: phases ms", [round(x, 1) for x in ms], "values only:", [round(x, 1) for x in ms_v], "check", out)


def test_config_c3_tall_65536x4096(D):
    # BASELINE.json configs[2] on one GPU (QR first); the sharded run of the same config is in the multi-GPU tests
    m, n = 65536, 4096
    sig, out, normA2, ms = _device_svd_with_check(D, m, n, vectors=True)
    b = EPS * m
    assert out[5] == 1.0 and out[0] <= 100 * b and out[1] <= 100 * b and out[2] <= 100 * b, out
    assert abs(np.dot(sig, sig) - normA2) <= 100 * b * normA2
    print("C3 65536x4096: phases ms", [round(x, 1) for x in ms], "check", out)


# ------------------------------------------------------------------ more than one GPU
def _need(D, n):
    if ngpus(D) < n:
        pytest.skip(f"needs {n} GPUs, this box has {ngpus(D)} (covered by gpurun --gpus {n} and by SCALE)")


@pytest.mark.parametrize("shape,world", [((1500, 1500), 2), ((2304, 2000), 2), ((700, 700), 2), ((5000, 300), 2),
                                         ((600, 1400), 2), ((130, 130), 2), ((3, 3), 2), ((1, 1), 2),
                                         ((3000, 3000), 4), ((2500, 900), 3), ((4096, 4096), 8)])
def test_sharded_local_group_is_an_svd(D, shape, world):
    # ONE process drives `world` GPUs (svdgpu_group_create_local): rank 0 factorizes and broadcasts prepared WY
    # panels while it does, every rank solves and back-transforms its block of singular values, every block is
    # copied to the host from its own GPU.  Reflectors and singular values are bitwise those of one GPU; the
    # vectors agree to rounding (the K-split of the GEMMs depends on the block width).
    _need(D, world)
    A = util.rand_matrix(*shape)
    g = D.Group.local(world)
    try:
        s, U, V, Am = g.svd(A)
        ms = [g.phase_ms(lr) for lr in range(world)]
    finally:
        g.destroy()
    bounds_ok(A, s, U, V)
    s1, U1, V1, Am1 = D.svd_gpu(A)
    assert np.array_equal(Am, Am1) and np.array_equal(s, s1)
    assert np.abs(U - U1).max() <= 1e-11 and np.abs(V - V1).max() <= 1e-11
    assert ms[0][1] > 0 and all(x[1] == 0 for x in ms[1:])


def test_svd_gpu_honours_ngpus(D):
    # the drop-in entry point itself on 2 GPUs (SVD_GPU_NGPUS / svd_gpu_set_option("ngpus", 2))
    _need(D, 2)
    A = util.rand_matrix(1800, 1700)
    s1, U1, V1, A1 = D.svd_gpu(A)
    try:
        D.set_option("ngpus", 2)
        s2, U2, V2, A2 = D.svd_gpu(A)
        sv, _, _, _ = D.svd_gpu(A, vectors=False)
    finally:
        D.set_option("ngpus", 1)
    assert np.array_equal(s1, s2) and np.array_equal(A1, A2)
    assert np.abs(U1 - U2).max() <= 1e-11 and np.abs(V1 - V2).max() <= 1e-11
    assert np.abs(sv - s1).max() <= 10 * EPS * 1800 * s1.max()
    exe = os.path.join(ROOT, "build", "dropin_check")
    r = _run([exe, "1500", "1500", "2"])
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])


_RANK_SCRIPT = r"""
import sys, os, ctypes, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import ddc_svd_b200 as D, util
rank, world, m, n, idfile, outdir = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6]
L = D.lib()
L.svdgpu_set_device(rank)
import time
if rank == 0:
    open(idfile + ".tmp", "wb").write(D.Group.unique_id()); os.rename(idfile + ".tmp", idfile)
while not os.path.exists(idfile):
    time.sleep(0.05)
g = D.Group.rank(world, rank, open(idfile, "rb").read())
mn = min(m, n)
blk, i0, ns = D.shard_range(mn, world, rank)
A = util.rand_matrix(m, n) if rank == 0 else None
sig = np.zeros(mn)
Ub = np.zeros((m, max(ns, 1)), order="F"); Vb = np.zeros((n, max(ns, 1)), order="F")
ub = (ctypes.c_void_p * 1)(Ub.ctypes.data); vb = (ctypes.c_void_p * 1)(Vb.ctypes.data)
for rep in range(2):
    Aw = np.array(A, order="F") if rank == 0 else None
    L.svd_gpu_sharded(g.h, m, n, util.p(Aw) if rank == 0 else None, util.p(sig) if rank == 0 else None, ub, vb)
np.savez(os.path.join(outdir, "r%d.npz" % rank), U=Ub[:, :ns], V=Vb[:, :ns], sig=sig, i0=i0, ns=ns, ms=np.array(g.phase_ms(0)))
g.destroy()
"""


@pytest.mark.parametrize("shape,world", [((1600, 1400), 2), ((3000, 2800), 4)])
def test_sharded_one_process_per_gpu(D, shape, world, tmp_path):
    # one process per GPU (svdgpu_group_create_rank, the torchrun layout bench.py uses): the unique id travels
    # through a file, every process receives only its own blocks
    _need(D, world)
    m, n = shape
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT.format(root=ROOT))
    idfile = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), str(m), str(n), idfile, str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE) for r in range(world)]
    for pr in procs:
        out, err = pr.communicate(timeout=600)
        assert pr.returncode == 0, err.decode()[-3000:]
    mn = min(m, n)
    U = np.zeros((m, mn)); V = np.zeros((n, mn)); sig = None
    for r in range(world):
        z = np.load(tmp_path / ("r%d.npz" % r))
        i0, ns = int(z["i0"]), int(z["ns"])
        U[:, i0:i0 + ns] = z["U"]; V[:, i0:i0 + ns] = z["V"]
        if r == 0:
            sig = z["sig"]
    A = util.rand_matrix(m, n)
    bounds_ok(A, sig, U, V)
    s1, U1, V1, _ = D.svd_gpu(A)
    assert np.array_equal(sig, s1)
    assert np.abs(U - U1).max() <= 1e-11 and np.abs(V - V1).max() <= 1e-11


# ------------------------------------------------------------------ tcgen05: FP64-accurate update from int8 slice products
@pytest.mark.parametrize("shape", [(128, 32), (128, 512), (300, 70), (1000, 513), (4100, 2050), (16384, 1024)])
@pytest.mark.parametrize("sign", [-1.0, 1.0])
def test_ozaki_update_vs_numpy(D, shape, sign):
    # C += sign * A (M x 128) * B (128 x N) on the int8 tensor cores (ozaki.cu): 8 x 8 error-free slices, 36 products,
    # int32 accumulators in TMEM, FP64 recombination — must match the FP64 product at the level of its own rounding
    M, N = shape
    K = 128
    L = D.lib()
    rng = np.random.default_rng(M + N)
    A = np.asfortranarray(rng.standard_normal((M, K)) * np.exp(rng.uniform(-6, 0, size=(M, 1))))   # rows of very different size
    B = np.asfortranarray(rng.standard_normal((K, N)) / 16.0)
    C = np.asfortranarray(rng.standard_normal((M, N)))
    bufs = []

    def put(a):
        d = L.svdgpu_malloc(a.nbytes); bufs.append(d)
        L.svdgpu_h2d(d, util.p(a), a.nbytes, None)
        return d
    try:
        dA, dB, dC = put(A), put(B), put(C)
        work = L.svdgpu_malloc(L.svdgpu_ozaki_workspace(M, N)); bufs.append(work)
        L.svdgpu_ozaki_update(M, N, sign, dA, M, dB, K, dC, M, work, None)
        out = np.empty((M, N), order="F")
        L.svdgpu_d2h(util.p(out), dC, out.nbytes, None)
        L.svdgpu_stream_sync(None)
    finally:
        for d in bufs:
            L.svdgpu_free(d)
    ref = C + sign * (A @ B)
    # error budget: the slices resolve 2^-56 of each operand's max-abs; FP64's own bound is K eps |A||B|
    scale = np.abs(A).max() * np.abs(B).max()
    err = np.abs(out - ref).max()
    print("ozaki update", shape, "max err %.3e = %.2f eps*K*scale" % (err, err / (EPS * K * scale)))
    assert err <= 2 * EPS * K * scale + 4 * EPS * np.abs(ref).max()


@pytest.mark.parametrize("n", [1, 2, 31, 33, 63, 65, 100, 257, 300, 511, 1000, 1025, 2049])
def test_twisted_every_entry_written_once(D, n):
    # svdgpu_twisted_vectors writes the normalised vectors straight into X / Y from several warps per sigma (position
    # segments, tw_solve_scan_kernel).  Poison the outputs AND the workspace, run twice, and require (i) no poison left,
    # (ii) bitwise identical results, (iii) orthonormal vectors: a segment that writes outside its positions (the bug
    # compute-sanitizer's timing exposed at n = 257: an empty trailing segment overwrote position n-1) fails (ii)/(iii).
    L = D.lib()
    A = util.rand_matrix(max(2 * n - 1, 1), n)
    _, al, be = util.oracle_bidiag(A)
    B = util.bidiag_dense(al, be)
    sv = np.ascontiguousarray(np.linalg.svd(B, compute_uv=False)[::-1])
    bep = np.zeros(n); bep[:n - 1] = be
    outs = []
    for fill in (0xFF, 0x7F):
        bufs = [L.svdgpu_malloc(8 * (3 * n + 2)), L.svdgpu_malloc(8 * n * n), L.svdgpu_malloc(8 * n * n)]
        wb = L.svdgpu_twisted_workspace(n, n, n)
        bufs.append(L.svdgpu_malloc(wb))
        d, dX, dY, work = bufs
        try:
            L.svdgpu_memset(work, fill, wb, None)
            L.svdgpu_memset(dX, fill, 8 * n * n, None); L.svdgpu_memset(dY, fill, 8 * n * n, None)
            L.svdgpu_h2d(d, util.p(al), 8 * n, None); L.svdgpu_h2d(d + 8 * n, util.p(bep), 8 * n, None)
            L.svdgpu_h2d(d + 16 * n + 8, util.p(sv), 8 * n, None)
            L.svdgpu_twisted_vectors(n, n, d, d + 8 * n, d + 16 * n + 8, n, 0, n, dX, n, dY, n, None, 1, work, None)
            X = np.zeros((n, n)); Y = np.zeros((n, n))
            L.svdgpu_d2h(util.p(X), dX, 8 * n * n, None); L.svdgpu_d2h(util.p(Y), dY, 8 * n * n, None)
            L.svdgpu_stream_sync(None)
        finally:
            for b in bufs:
                L.svdgpu_free(b)
        assert np.isfinite(X).all() and np.isfinite(Y).all()
        assert np.abs(X).max() <= 1.0 + 1e-12 and np.abs(Y).max() <= 1.0 + 1e-12
        outs.append((X, Y))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    X, Y = outs[0]
    c = 100 * EPS * max(n, 8)
    assert np.linalg.norm(X @ X.T - np.eye(n)) <= c and np.linalg.norm(Y @ Y.T - np.eye(n)) <= c
    if n > 1:
        assert np.linalg.norm(B - Y.T @ np.diag(sv) @ X) / np.linalg.norm(B) <= c
