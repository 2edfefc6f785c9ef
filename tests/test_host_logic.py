"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the
headers declare what the reference's client includes, the numpy model of the device numerics
meets the north-star bounds, and the host helper functions behave like the reference's."""
import os
import re
import subprocess

import numpy as np
import pytest

import util
from util import p

ROOT = util.ROOT
INC = os.path.join(ROOT, "include")


def _declared_symbols():
    names = set()
    for h in sorted(os.listdir(INC)):
        txt = open(os.path.join(INC, h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        for m in re.finditer(r"^\s*(?:const\s+)?(?:void|int|float|double|size_t|char|unsigned long long|svdgpu_group)\s*\*?\s*\*?\s*(\w+)\s*\(", txt, re.M):
            names.add(m.group(1))
    return names


def test_library_exports_every_declared_symbol():
    import ddc_svd_b200 as D
    assert os.path.exists(D.LIB_PATH), "libsvdgpu.so not built: run make / __graft_entry__.build()"
    out = subprocess.check_output(["nm", "-D", "--defined-only", D.LIB_PATH]).decode()
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    declared = _declared_symbols()
    assert len(declared) > 50
    missing = sorted(declared - exported)
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    # the ctypes table binds the same set (AttributeError inside lib() would mean a gap)
    D.lib()
    assert {s[0] for s in D.SIGNATURES} == declared


def test_reference_entry_point_signature():
    txt = open(os.path.join(INC, "svd_gpu.h")).read()
    assert "#ifndef SVDGPU" in txt
    assert re.search(r"void\s+svd_gpu\(int m, int n, double\* A,double \* sigma, double \* U, double\* V\);", txt)
    for h in ("cl-helper.h", "matrix_helper.h", "svd_gpu.h"):          # test-whole-svd.c:4-7
        assert os.path.exists(os.path.join(INC, h))


@pytest.mark.skipif(not os.path.exists("/root/reference/test-whole-svd.c"), reason="reference not mounted")
def test_dropin_client_links_unmodified():
    subprocess.check_call(["make", "-C", ROOT, "dropin"], stdout=subprocess.DEVNULL)
    assert os.path.exists(os.path.join(ROOT, "build", "test-whole-svd"))


def test_host_helpers_match_reference_semantics():
    import ddc_svd_b200 as D
    L = D.lib()
    rng = np.random.default_rng(3)
    M, N, K = 7, 5, 4
    A = np.asfortranarray(rng.standard_normal((M, N)))
    AT = np.zeros((N, M), order="F")
    L.transpose(M, N, p(A), p(AT))
    assert np.array_equal(AT, A.T)
    B = np.asfortranarray(rng.standard_normal((N, K)))
    C = np.asfortranarray(rng.standard_normal((M, K)))
    C0 = C.copy()
    L.dgemm_simple(M, K, N, p(A), p(B), p(C))                 # accumulates into C
    assert np.allclose(C, C0 + A @ B)
    al = rng.standard_normal(4); be = rng.standard_normal(4)
    Bm = np.zeros((4, 6), order="F")
    L.form_bidiag(4, 6, p(al), p(be), p(Bm))
    assert np.array_equal(Bm, util.bidiag_dense(al, be, 6))
    Bm = np.zeros((6, 4), order="F")
    L.form_bidiag(6, 4, p(al), p(be), p(Bm))
    ref = np.zeros((6, 4)); ref[:4, :4] = util.bidiag_dense(al, be[:3])
    assert np.array_equal(Bm, ref)
    assert np.isclose(L.l2_norm_mat(M, N, p(A)), np.linalg.norm(A))
    assert np.isclose(L.dot_prod(5, p(al.copy()[:4].repeat(2)[:5].copy()), p(np.ones(5))), al[:4].repeat(2)[:5].sum())
    assert np.isclose(L.l2_norm_mat_row(M, N, N, p(A)), np.linalg.norm(A[0, :]))


def test_compute_entry_fails_loudly_without_gpu():
    # no CPU fallback: on a machine without CUDA the library must abort, not compute
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    code = ("import sys; sys.path.insert(0, %r); import numpy as np, ddc_svd_b200 as D; "
            "D.svd_gpu(np.ones((4,4)))") % ROOT
    r = subprocess.run(["python", "-c", code], capture_output=True)
    assert r.returncode != 0
    assert b"no usable CUDA device" in r.stderr or b"failed with error" in r.stderr


@pytest.mark.parametrize("n", [40, 150])
def test_device_numerics_model_meets_bounds(n):
    # executable specification of the kernels' numerics (tests/model_numerics.py)
    import model_numerics as mn
    A = util.rand_matrix(n, n)
    _, al, be = util.oracle_bidiag(A)
    bep = np.zeros(n); bep[: n - 1] = be
    B = util.bidiag_dense(al, be)
    sv = np.linalg.svd(B, compute_uv=False)[::-1]
    sig = mn.ddc_values(al, bep)
    assert np.abs(sig - sv).max() / sv.max() < 10 * util.EPS * n
    with np.errstate(all="ignore"):
        X, sig2 = mn.twisted_vectors(al, be, sig)
    Y = mn.left_from_right(al, be, sig2, X)
    c = 100 * util.EPS * n
    assert np.linalg.norm(X @ X.T - np.eye(n)) < c
    assert np.linalg.norm(Y @ Y.T - np.eye(n)) < c
    assert np.linalg.norm(B - Y.T @ np.diag(sig2) @ X) / np.linalg.norm(B) < c


def test_on_chip_tail_planning():
    # bidiag.cu:bidiag_tail_start - pure host logic: the tail kernel takes over at the first panel boundary
    # whose trailing block fits 148 x 28000 doubles of shared memory (<= 2048 rows, <= 16 columns per CTA)
    import ddc_svd_b200 as D
    f = D.lib().svdgpu_bidiag_tail_start
    assert f(512, 512, 32, 148) == 0 and f(1, 1, 32, 148) == 0          # small inputs: entirely on chip
    assert f(1500, 1400, 32, 148) == 0
    i0 = f(4096, 4096, 32, 148)
    assert i0 % 32 == 0 and 0 < i0 < 4096
    L = 4096 - i0
    assert L <= 2048 and -(-L // 148) * L <= 28000                      # fits ...
    Lprev = L + 32
    assert Lprev > 2048 or -(-Lprev // 148) * Lprev > 28000             # ... and the boundary before did not
    assert f(16384, 16384, 32, 148) == 16384 - L                        # same trailing size for any square input
    assert f(200, 300, 32, 148) == 200                                  # wide inputs never (they arrive transposed)
    assert f(65536, 4096, 32, 148) == 4096                              # too tall: 61k rows never fit
    # fewer CTAs (MIG slice, cut-down part): 8 x 16 columns would fit at 128 trailing rows, but a CTA has only
    # 15 row-owning warps, so ceil(127 / 8) = 16 rows per CTA do not: the hand-over waits for 96 rows
    assert f(4096, 4096, 32, 8) == 4096 - 96
    for ctas in (8, 16, 37, 64, 100, 132, 148):
        i0 = f(4096, 4096, 32, ctas)
        assert i0 == 4096 or -(-(4096 - i0 - 1) // ctas) <= 15


@pytest.mark.parametrize("shape,i0,G", [((1, 1), 0, 3), ((3, 3), 0, 5), ((30, 28), 0, 5), ((60, 60), 0, 7),
                                        ((64, 64), 0, 4), ((97, 3), 0, 6), ((40, 1), 0, 2), ((80, 70), 0, 11)])
def test_on_chip_tail_model_matches_oracle(shape, i0, G):
    # tests/model_tail.py is the numpy specification bidiag_tail_kernel was written from: same column
    # distribution, same phases, same exchanged buffers; it must reproduce the reference's bidiagonalization
    import model_tail
    m, n = shape
    A = util.rand_matrix(m, n, 1.0, 2.0, 4)
    Ao, ao, bo = util.oracle_bidiag(A)
    Ag, ag, bg = model_tail.tail_model(A, i0, G)
    assert np.abs(Ag - Ao).max() <= 1e-11 and np.abs(ag - ao).max() <= 1e-11
    if n > 1:
        assert np.abs(bg - bo).max() <= 1e-11


def test_shard_range_c_matches_python_model():
    # svdgpu_shard_range (svd_gpu.c) is the partitioning every rank of a sharded run derives for itself; the CPU
    # model in ddc_svd_b200/sharding.py (exercised over gloo) must be the same function
    import ddc_svd_b200 as D
    from ddc_svd_b200.sharding import shard_range
    for mn in (1, 2, 3, 7, 64, 100, 129, 4096, 16384, 32768):
        for world in (1, 2, 3, 4, 5, 8):
            cover = []
            for r in range(world):
                assert D.shard_range(mn, world, r) == shard_range(mn, world, r)
                _, i0, ns = D.shard_range(mn, world, r)
                cover += list(range(i0, i0 + ns))
            assert cover == list(range(mn))


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (5, 5), (130, 130), (4096, 4096), (16384, 16384), (5000, 4096),
                                   (65536, 4096), (300, 900), (4096, 32768), (700, 129)])
def test_panel_chunk_plan_is_the_same_on_every_rank(shape):
    # svdgpu_plan_chunks: the canonical order in which rank 0 prepares / broadcasts compact-WY panel chunks.  The other
    # ranks post their receives from this list before rank 0 has produced anything, so it must depend on (m, n) only,
    # cover every panel of every reflector set exactly once, and never ask for a chunk before an earlier one.
    import ctypes
    import ddc_svd_b200 as D
    L = D.lib()
    m, n = shape
    plans = []
    for world in (1, 2, 8):
        route = (ctypes.c_int * 4)()
        cap = 4096
        arr = [(ctypes.c_int * cap)() for _ in range(4)]
        cnt = L.svdgpu_plan_chunks(m, n, world, route, cap, *arr)
        assert 0 < cnt <= cap
        plans.append((list(route), [tuple(a[i] for a in arr) for i in range(cnt)]))
    assert plans[0] == plans[1] == plans[2]
    route, chunks = plans[0]
    wide, qr, tm, tn = route
    assert wide == (1 if m < n else 0) and (tm, tn) == ((n, m) if m < n else (m, n))
    assert qr == (1 if tm * 10 >= tn * 25 and tm > tn and tn >= 2 else 0)
    nL = min(tm, tn) if not qr else tn
    nR = (tn - 2 if tn >= 2 else 0)
    expect = {1: nL, 2: nR}
    if qr:
        expect[0] = tn
    seen = {}
    last_need = {0: 0, 1: 0}
    for st, pb, pe, need in chunks:
        assert pb == seen.get(st, 0) and pe > pb            # contiguous, in order, per set
        seen[st] = pe
        stage = 0 if st == 0 else 1
        assert need >= last_need[stage] or need == expect[st]   # the producing factorization only moves forward
        last_need[stage] = max(last_need[stage], need)
        assert need <= expect[st]
    for st, nref in expect.items():
        if nref > 0:
            assert seen[st] == -(-nref // 128), (st, seen, nref)
    if qr:      # the QR's chunks all come before the bidiagonalization's
        first_bd = min(i for i, c in enumerate(chunks) if c[0] != 0)
        assert all(c[0] == 0 for c in chunks[:first_bd]) and all(c[0] != 0 for c in chunks[first_bd:])
