"""ddc_svd_b200 — host-side Python mirror of the reference's C interface for the svd_gpu() path.

The product is ``libsvdgpu.so`` (C ABI, built by ``make`` / ``__graft_entry__.build()`` from
``csrc/*.cu`` + ``host/*.c`` for sm_100a).  This module only binds it with ctypes, with the
reference's names, argument meaning and layouts (column-major, ascending sigma):

    svd_gpu(m, n, A, sigma, U, V)                       svd_gpu.h:5
    bidiag_par(m, n, A, alpha, beta)                    bidiag_par.h:30
    GetSingularValues_Parallel(N, b1, b2, sigma)        Calculations-Parallel.h:41
    CalcRightSingularVectors / RighttoLeftSingularVectors   parallel-twisted.h:18-19
    multU / multV                                       bidiag_par.h:72-73

There is no CPU fallback: importing works anywhere, but every compute call needs the
library and a CUDA device and fails loudly otherwise.  Nothing here touches ``oracle/``.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVD_GPU_LIB") or os.path.join(_HERE, "libsvdgpu.so")   # SVD_GPU_LIB: A/B builds (bench/ab_build.sh)
_lib = None

c_double_p = ctypes.POINTER(ctypes.c_double)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_void_p = ctypes.c_void_p
c_void_pp = ctypes.POINTER(ctypes.c_void_p)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_int, c_long, c_size_t, c_double = ctypes.c_int, ctypes.c_long, ctypes.c_size_t, ctypes.c_double

# (name, restype, argtypes) for every symbol include/*.h declares
SIGNATURES = [
    # include/svd_gpu.h
    ("svd_gpu", None, [c_int, c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    # include/svd_gpu_b200.h
    ("svd_gpu_dev", None, [c_int, c_int, c_void_p, c_long, c_void_p, c_void_p, c_long, c_void_p, c_long, c_void_p]),
    ("svd_gpu_vectors_dev", None, [c_int, c_int, c_void_p, c_long, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                   c_void_p, c_long, c_void_p, c_long, c_void_p, c_void_p]),
    ("svd_gpu_values_dev", None, [c_int, c_int, c_void_p, c_long, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("svd_gpu_last_phase_ms", None, [c_float_p]),
    ("svd_gpu_set_option", None, [ctypes.c_char_p, c_int]),
    ("svdgpu_group_create_local", c_void_p, [c_int, c_int_p]),
    ("svdgpu_group_create_rank", c_void_p, [c_int, c_int, c_void_p]),
    ("svdgpu_group_destroy", None, [c_void_p]),
    ("svdgpu_group_size", c_int, [c_void_p]),
    ("svdgpu_group_nlocal", c_int, [c_void_p]),
    ("svdgpu_group_rank", c_int, [c_void_p, c_int]),
    ("svdgpu_group_device", c_int, [c_void_p, c_int]),
    ("svdgpu_shard_range", None, [c_int, c_int, c_int, c_int_p, c_int_p, c_int_p]),
    ("svdgpu_plan_chunks", c_int, [c_int, c_int, c_int, c_int_p, c_int, c_int_p, c_int_p, c_int_p, c_int_p]),
    ("svd_gpu_sharded_dev", None, [c_void_p, c_int, c_int, c_void_p, c_long, c_void_pp, c_void_pp, c_long,
                                   c_void_pp, c_long, c_void_pp]),
    ("svd_gpu_sharded", None, [c_void_p, c_int, c_int, c_double_p, c_double_p, c_void_pp, c_void_pp]),
    ("svd_gpu_group_phase_ms", None, [c_void_p, c_int, c_float_p]),
    ("svdgpu_fill_rand", None, [c_double_p, c_size_t, c_double, c_double, ctypes.c_uint]),
    ("svd_gpu_check", None, [c_int, c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    ("svd_gpu_check_dev", None, [c_int, c_int, c_void_p, c_long, c_void_p, c_void_p, c_long, c_void_p, c_long,
                                 c_int, c_double_p, c_void_p]),
    # include/bidiag_par.h
    ("bidiag_par", None, [c_int, c_int, c_double_p, c_double_p, c_double_p]),
    ("form_u_par", None, [c_int, c_int, c_double_p, c_double_p]),
    ("form_v_par", None, [c_int, c_int, c_double_p, c_double_p]),
    ("multU", None, [c_int, c_int, c_int, c_double_p, c_double_p, c_double_p]),
    ("multV", None, [c_int, c_int, c_int, c_double_p, c_double_p, c_double_p]),
    ("svd_gpu_backtransform", None, [c_int, c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    # include/Calculations-Parallel.h
    ("GetSingularValues_Parallel", None, [c_int, c_double_p, c_double_p, c_double_p]),
    # include/parallel-twisted.h
    ("CalcRightSingularVectors", None, [c_int, c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    ("RighttoLeftSingularVectors", None, [c_int, c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    # include/matrix_helper.h
    ("transpose", None, [c_int, c_int, c_double_p, c_double_p]),
    ("form_bidiag", None, [c_int, c_int, c_double_p, c_double_p, c_double_p]),
    ("dgemm_simple", None, [c_int, c_int, c_int, c_double_p, c_double_p, c_double_p]),
    ("l2_norm_mat", c_double, [c_int, c_int, c_double_p]),
    ("l2_normv", c_double, [c_int, c_double_p]),
    ("dot_prod", c_double, [c_int, c_double_p, c_double_p]),
    ("scale_vector", None, [c_int, c_double_p, c_double]),
    ("l2_norm_mat_row", c_double, [c_int, c_int, c_int, c_double_p]),
    ("scale_mat_row", None, [c_int, c_int, c_int, c_double_p, c_double]),
    ("dot_prod_mat_rows", c_double, [c_int, c_int, c_int, c_double_p, c_double_p]),
    ("dot_prod_mat_row_with_vec", c_double, [c_int, c_int, c_int, c_double_p, c_double_p]),
    ("set_vec_to_zero", None, [c_int, c_double_p]),
    ("print_matrix", None, [c_double_p, c_long, c_long, ctypes.c_char_p]),
    # include/cuda-helper.h
    ("svdgpu_device_count", c_int, []),
    ("svdgpu_set_device", None, [c_int]),
    ("svdgpu_get_device", c_int, []),
    ("svdgpu_device_name", ctypes.c_char_p, []),
    ("svdgpu_malloc", c_void_p, [c_size_t]),
    ("svdgpu_free", None, [c_void_p]),
    ("svdgpu_memset", None, [c_void_p, c_int, c_size_t, c_void_p]),
    ("svdgpu_h2d", None, [c_void_p, c_void_p, c_size_t, c_void_p]),
    ("svdgpu_d2h", None, [c_void_p, c_void_p, c_size_t, c_void_p]),
    ("svdgpu_d2d", None, [c_void_p, c_void_p, c_size_t, c_void_p]),
    ("svdgpu_h2d_2d", None, [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p]),
    ("svdgpu_d2h_2d", None, [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p]),
    ("svdgpu_stream_create", c_void_p, []),
    ("svdgpu_stream_create_priority", c_void_p, [c_int]),
    ("svdgpu_host_register", c_int, [c_void_p, c_size_t]),
    ("svdgpu_host_unregister", None, [c_void_p]),
    ("svdgpu_device_sync", None, []),
    ("svdgpu_enable_peer_access", c_int, [c_int, c_int]),
    ("svdgpu_range_push", None, [ctypes.c_char_p]),
    ("svdgpu_range_pop", None, []),
    ("svdgpu_nccl_version", c_int, []),
    ("svdgpu_nccl_unique_id", None, [c_void_p]),
    ("svdgpu_nccl_comm_init_rank", c_void_p, [c_int, c_int, c_void_p]),
    ("svdgpu_nccl_comm_init_all", None, [c_int, c_int_p, c_void_pp]),
    ("svdgpu_nccl_comm_destroy", None, [c_void_p]),
    ("svdgpu_nccl_comm_count", c_int, [c_void_p]),
    ("svdgpu_nccl_group_start", None, []),
    ("svdgpu_nccl_group_end", None, []),
    ("svdgpu_nccl_bcast", None, [c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    ("svdgpu_nccl_allgather", None, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    ("svdgpu_stream_destroy", None, [c_void_p]),
    ("svdgpu_stream_sync", None, [c_void_p]),
    ("svdgpu_stream_wait_event", None, [c_void_p, c_void_p]),
    ("svdgpu_event_create", c_void_p, []),
    ("svdgpu_event_destroy", None, [c_void_p]),
    ("svdgpu_event_record", None, [c_void_p, c_void_p]),
    ("svdgpu_event_elapsed_ms", ctypes.c_float, [c_void_p, c_void_p]),
    ("svdgpu_launch_count", ctypes.c_ulonglong, []),
    ("svdgpu_host_alloc", c_void_p, [c_size_t]),
    ("svdgpu_host_free", None, [c_void_p]),
    ("svdgpu_bidiag_workspace", c_size_t, [c_int, c_int, c_long]),
    ("svdgpu_bidiag", None, [c_int, c_int, c_void_p, c_long, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    ("svdgpu_bidiag_progress", None, [c_int, c_int, c_void_p, c_long, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                      c_void_p]),
    ("svdgpu_bidiag_tail_start", c_int, [c_int, c_int, c_int, c_int]),
    ("svdgpu_ddc_workspace", c_size_t, [c_int]),
    ("svdgpu_ddc_values", None, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("svdgpu_twisted_workspace", c_size_t, [c_int, c_int, c_int]),
    ("svdgpu_twisted_vectors", None, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                      c_long, c_void_p, c_long, c_void_p, c_int, c_void_p, c_void_p]),
    ("svdgpu_backtransform_workspace", c_size_t, [c_int, c_int, c_int]),
    ("svdgpu_transpose", None, [c_int, c_int, c_void_p, c_long, c_void_p, c_long, c_void_p]),
    ("svdgpu_qr_workspace", c_size_t, [c_int, c_int]),
    ("svdgpu_qr", None, [c_int, c_int, c_void_p, c_long, c_void_p, c_long, c_void_p, c_void_p]),
    ("svdgpu_wy_apply", None, [c_int, c_int, c_int, c_void_p, c_long, c_void_p, c_long, c_int, c_void_p, c_void_p]),
    ("svdgpu_qr_progress", None, [c_int, c_int, c_void_p, c_long, c_void_p, c_long, c_void_p, c_void_p, c_void_p]),
    ("svdgpu_wy_panel_width", c_int, []),
    ("svdgpu_wy_panel_count", c_int, [c_int]),
    ("svdgpu_wy_panels_bytes", c_size_t, [c_int, c_int]),
    ("svdgpu_wy_apply_workspace", c_size_t, [c_int]),
    ("svdgpu_wy_setup", None, [c_int, c_int, c_int, c_void_p, c_long, c_void_p, c_int, c_int, c_void_p]),
    ("svdgpu_wy_apply_prepared", None, [c_int, c_int, c_int, c_void_p, c_void_p, c_long, c_int, c_void_p, c_void_p]),
    ("svdgpu_wy_panel_slices", None, [c_void_p, c_int, c_int, c_int, c_int, c_void_pp, c_void_pp,
                                      ctypes.POINTER(c_size_t)]),
    ("svdgpu_check_workspace", c_size_t, [c_int, c_int, c_int]),
    ("svdgpu_check", None, [c_int, c_int, c_void_p, c_long, c_void_p, c_void_p, c_long, c_void_p, c_long, c_int,
                            c_void_p, c_void_p, c_void_p]),
    ("svdgpu_ozaki_workspace", c_size_t, [c_int, c_int]),
    ("svdgpu_ozaki_update", None, [c_int, c_int, c_double, c_void_p, c_long, c_void_p, c_long, c_void_p, c_long,
                                   c_void_p, c_void_p]),
    ("svdgpu_dgemm", None, [c_int, c_int, c_int, c_int, c_int, c_double, c_void_p, c_long, c_void_p, c_long,
                            c_double, c_void_p, c_long, c_void_p]),
    ("svdgpu_scale_matrix", None, [c_int, c_int, c_void_p, c_long, c_void_p, c_void_p, c_void_p]),
    ("svdgpu_scale_vector", None, [c_int, c_void_p, c_void_p, c_void_p]),
    ("svdgpu_bidiag_pass_probe", None, [c_int, c_int, c_void_p, c_long, c_void_p, c_int, c_void_p]),
]


def lib():
    """The loaded C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
                "ddc_svd_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, res, args in SIGNATURES:
            fn = getattr(L, name)            # AttributeError = a declared symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(c_double_p)


def _colmajor(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("expected a 2-D matrix")
    return np.array(a, dtype=np.float64, order="F", copy=True)


# ----------------------------------------------------------------- reference-shaped host API
def svd_gpu(A, vectors=True):
    """svd_gpu(m,n,A,sigma,U,V) on a 2-D array.  Returns (sigma ascending, U m x mn, V n x mn,
    A_mod) — A_mod is the overwritten A (the Householder reflectors), as the reference leaves it."""
    Af = _colmajor(A)
    m, n = Af.shape
    mn = min(m, n)
    sigma = np.zeros(mn)
    if vectors:
        U = np.zeros((m, mn), order="F")
        V = np.zeros((n, mn), order="F")
        lib().svd_gpu(m, n, _p(Af), _p(sigma), _p(U), _p(V))
        return sigma, U, V, Af
    lib().svd_gpu(m, n, _p(Af), _p(sigma), None, None)
    return sigma, None, None, Af


def bidiag_par(A):
    """bidiag_par(m,n,A,alpha,beta): returns (A_mod, alpha[min(m,n)], beta[n-1 | m])."""
    Af = _colmajor(A)
    m, n = Af.shape
    mn = min(m, n)
    alpha = np.zeros(mn)
    beta = np.zeros(max(n - 1 if m >= n else m, 1))
    lib().bidiag_par(m, n, _p(Af), _p(alpha), _p(beta))
    return Af, alpha, beta[: (n - 1 if m >= n else m)]


def get_singular_values(b1, b2):
    """GetSingularValues_Parallel(N,b1,b2,sigma); b2 is zero-padded to N entries."""
    b1 = np.ascontiguousarray(b1, dtype=np.float64)
    N = b1.shape[0]
    b2p = np.zeros(N)
    b2p[: len(b2)] = b2
    sigma = np.zeros(N)
    lib().GetSingularValues_Parallel(N, _p(b1), _p(b2p), _p(sigma))
    return sigma


def singular_vectors(a, b, sigma, m=None):
    """CalcRightSingularVectors + RighttoLeftSingularVectors.  Returns X (n x m, row i = x_i, the
    reference's X[i*m+j]) and Y (n x n, row i = y_i)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    m = n if m is None else m
    bp = np.zeros(max(m - 1, 1))
    bp[: min(len(b), m - 1)] = np.asarray(b, dtype=np.float64)[: m - 1]
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    X = np.zeros((n, m))
    Y = np.zeros((n, n))
    lib().RighttoLeftSingularVectors(n, m, _p(a), _p(bp), _p(sigma), _p(X), _p(Y))
    return X, Y


def backtransform(A_mod, X, Y):
    """All of multU/multV at once: U = Q_L [Y^T;0], V = Q_R [X^T;0] (X, Y as returned above)."""
    Af = _colmajor(A_mod)
    m, n = Af.shape
    mn = min(m, n)
    U = np.zeros((m, mn), order="F")
    V = np.zeros((n, mn), order="F")
    Xc = np.ascontiguousarray(X, dtype=np.float64)
    Yc = np.ascontiguousarray(Y, dtype=np.float64)
    lib().svd_gpu_backtransform(m, n, _p(Af), _p(Xc), _p(Yc), _p(U), _p(V))
    return U, V


def form_q(A_mod):
    """form_u_par / form_v_par: explicit Q_L (m x m) and Q_R (n x n) of the bidiagonalization."""
    Af = _colmajor(A_mod)
    m, n = Af.shape
    U = np.zeros((m, m), order="F")
    V = np.zeros((n, n), order="F")
    lib().form_u_par(m, n, _p(Af), _p(U))
    lib().form_v_par(m, n, _p(Af), _p(V))
    return U, V


def qr_tall(A):
    """Householder QR used by the QR-first route of svd_gpu() for m >> n (include/cuda-helper.h,
    svdgpu_qr): returns (A_qr, R, Q1) with A_qr the reflector storage, R (n x n) and the thin
    Q1 = Q [I_n; 0] (m x n) formed through the same compact-WY apply the SVD uses."""
    L = lib()
    Af = _colmajor(A)
    m, n = Af.shape
    nb = Af.nbytes
    dA = L.svdgpu_malloc(nb)
    dR = L.svdgpu_malloc(8 * n * n)
    dQ = L.svdgpu_malloc(nb)
    work = L.svdgpu_malloc(max(L.svdgpu_qr_workspace(m, n), L.svdgpu_backtransform_workspace(m, n, n)))
    try:
        L.svdgpu_h2d(dA, _p(Af), nb, None)
        L.svdgpu_qr(m, n, dA, m, dR, n, work, None)
        Q1 = np.zeros((m, n), order="F")
        Q1[np.arange(n), np.arange(n)] = 1.0
        L.svdgpu_h2d(dQ, _p(Q1), nb, None)
        L.svdgpu_wy_apply(1, m, n, dA, m, dQ, m, n, work, None)
        A_qr = np.empty((m, n), order="F")
        R = np.empty((n, n), order="F")
        L.svdgpu_d2h(_p(A_qr), dA, nb, None)
        L.svdgpu_d2h(_p(R), dR, 8 * n * n, None)
        L.svdgpu_d2h(_p(Q1), dQ, nb, None)
        L.svdgpu_stream_sync(None)
    finally:
        for d in (dA, dR, dQ, work):
            L.svdgpu_free(d)
    return A_qr, R, Q1


def shard_range(mn, world, rank):
    """(blk, i0, ns) of rank's block of singular values (svdgpu_shard_range)."""
    b, i0, ns = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib().svdgpu_shard_range(mn, world, rank, ctypes.byref(b), ctypes.byref(i0), ctypes.byref(ns))
    return b.value, i0.value, ns.value


class Group:
    """A group of ranks, one per GPU (include/svd_gpu_b200.h).  Group.local(n) drives n GPUs from this
    process; Group.rank(world, rank, id128) is one rank of a one-process-per-GPU job (id128 = the 128 bytes
    Group.unique_id() returns on rank 0, shared out of band, e.g. with torch.distributed.broadcast)."""

    def __init__(self, handle):
        self.h = handle

    @staticmethod
    def unique_id():
        buf = ctypes.create_string_buffer(128)
        lib().svdgpu_nccl_unique_id(buf)
        return buf.raw

    @classmethod
    def local(cls, ndev, devices=None):
        arr = (ctypes.c_int * ndev)(*devices) if devices is not None else None
        return cls(lib().svdgpu_group_create_local(ndev, arr))

    @classmethod
    def rank(cls, world, rank, id128):
        buf = ctypes.create_string_buffer(bytes(id128), 128)
        return cls(lib().svdgpu_group_create_rank(world, rank, buf))

    @property
    def world(self):
        return lib().svdgpu_group_size(self.h)

    @property
    def nlocal(self):
        return lib().svdgpu_group_nlocal(self.h)

    def global_rank(self, local):
        return lib().svdgpu_group_rank(self.h, local)

    def phase_ms(self, local=0):
        ms = (ctypes.c_float * 8)()
        lib().svd_gpu_group_phase_ms(self.h, local, ms)
        return list(ms)

    def destroy(self):
        if self.h:
            lib().svdgpu_group_destroy(self.h)
            self.h = None

    def svd(self, A):
        """svd_gpu_sharded on a host matrix with a LOCAL group: returns (sigma, U, V, A_mod), every rank's
        column block copied from its own GPU into place."""
        Af = _colmajor(A)
        m, n = Af.shape
        mn = min(m, n)
        sigma = np.zeros(mn)
        U = np.zeros((m, mn), order="F")
        V = np.zeros((n, mn), order="F")
        nl = self.nlocal
        ub = (ctypes.c_void_p * nl)()
        vb = (ctypes.c_void_p * nl)()
        for lr in range(nl):
            _, i0, _ = shard_range(mn, self.world, self.global_rank(lr))
            ub[lr] = U.ctypes.data + 8 * i0 * m
            vb[lr] = V.ctypes.data + 8 * i0 * n
        lib().svd_gpu_sharded(self.h, m, n, _p(Af), _p(sigma), ub, vb)
        return sigma, U, V, Af


def check(A, sigma, U, V):
    """svd_gpu_check: the reference driver's dormant residual check (test-whole-svd.c:81-96) run on the GPU.
    Returns dict(orthU, orthV, resid, checksum, normA, ascending)."""
    A0 = _colmajor(A)
    m, n = A0.shape
    out = np.zeros(6)
    lib().svd_gpu_check(m, n, _p(A0), _p(np.ascontiguousarray(sigma, dtype=np.float64)),
                        _p(_colmajor(U)), _p(_colmajor(V)), _p(out))
    return dict(orthU=out[0], orthV=out[1], resid=out[2], checksum=out[3], normA=out[4], ascending=bool(out[5] == 1.0))


def last_phase_ms():
    """[h2d, bidiag, dDC, twisted, back-transform, d2h, total] of the last svd_gpu/svd_gpu_dev call."""
    ms = (ctypes.c_float * 7)()
    lib().svd_gpu_last_phase_ms(ms)
    return list(ms)


def set_option(name, value):
    lib().svd_gpu_set_option(name.encode(), int(value))


# ----------------------------------------------------------------- device-resident API (torch)
def svd_gpu_dev(tA, tsigma, tU=None, tV=None, stream=None):
    """svd_gpu_dev on torch CUDA tensors.  tA holds the m x n matrix COLUMN-MAJOR, i.e. it is a
    contiguous (n, lda) float64 tensor whose row j is column j of A (lda even >= m); tU (mn, m) and
    tV (mn, n) likewise hold U and V column by column.  Enqueues on `stream` (default: torch's
    current stream) so torch.cuda.Event timing brackets it."""
    import torch
    n, lda = tA.shape
    mn = tsigma.shape[0]
    m = tU.shape[1] if tU is not None else None
    if m is None:
        raise ValueError("pass m via svd_gpu_dev_raw for values-only runs")
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    lib().svd_gpu_dev(m, n, tA.data_ptr(), lda, tsigma.data_ptr(), tU.data_ptr(), tU.shape[1],
                      tV.data_ptr(), tV.shape[1], st)


def svd_gpu_dev_raw(m, n, dA, lda, dsigma, dU, ldu, dV, ldv, stream):
    lib().svd_gpu_dev(m, n, dA, lda, dsigma, dU, ldu, dV, ldv, stream)
