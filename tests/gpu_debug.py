"""First-contact diagnostic run on the GPU box (not a pytest file): prints per-phase checks."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import util
import ddc_svd_b200 as D
import model_numerics as mn

def section(s): print("\n==== " + s, flush=True)

L = D.lib()
print("device:", L.svdgpu_device_name().decode())

section("dgemm")
import ctypes
rng = np.random.default_rng(0)
def dev_put(a):
    a = np.ascontiguousarray(a)
    d = L.svdgpu_malloc(a.nbytes); L.svdgpu_h2d(d, a.ctypes.data, a.nbytes, None); L.svdgpu_stream_sync(None); return d
def dev_get(d, shape):
    a = np.empty(shape); L.svdgpu_d2h(a.ctypes.data, d, a.nbytes, None); L.svdgpu_stream_sync(None); return a
for (ta, tb, M, N, K) in [(0,0,130,70,50),(1,0,64,200,333),(0,1,257,129,64),(1,1,65,66,67)]:
    A = rng.standard_normal((M, K)); B = rng.standard_normal((K, N)); C = rng.standard_normal((M, N))
    As = np.asfortranarray(A.T if ta else A); Bs = np.asfortranarray(B.T if tb else B); Cs = np.asfortranarray(C)
    dA = dev_put(As.T); dB = dev_put(Bs.T); dC = dev_put(Cs.T)   # .T of F-order -> C-contiguous with same memory
    L.svdgpu_dgemm(ta, tb, M, N, K, -1.0, dA, As.shape[0], dB, Bs.shape[0], 1.0, dC, M, None)
    out = dev_get(dC, (N, M)).T
    print((ta,tb,M,N,K), "max err", np.abs(out - (C - A @ B)).max())

section("bidiag_par vs oracle (srand(4), [1,2))")
for (m, n) in [(8,8),(64,64),(100,100),(200,200),(513,512),(300,200),(200,300),(33,1),(1,33),(2,2)]:
    A = util.rand_matrix(m, n, 1.0, 2.0, 4)
    Ao, ao, bo = util.oracle_bidiag(A)
    t=time.time(); Ag, ag, bg = D.bidiag_par(A); dt=time.time()-t
    print((m,n), "A err", np.abs(Ag-Ao).max(), "alpha", np.abs(ag-ao).max(), "beta", (np.abs(bg-bo).max() if len(bo) else 0.0), "nan", int(np.isnan(Ag).sum()), "t", round(dt,3))

section("dDC vs LAPACK / model")
for n in [3, 5, 17, 64, 200, 512, 1024]:
    A = util.rand_matrix(n, n)
    _, al, be = util.oracle_bidiag(A)
    sv = np.linalg.svd(util.bidiag_dense(al, be), compute_uv=False)[::-1]
    sg = D.get_singular_values(al, be)
    bep = np.zeros(n); bep[:n-1] = be
    sm = mn.ddc_values(al, bep) if n <= 512 else sv
    print(n, "gpu vs lapack abs/max", np.abs(sg-sv).max()/sv.max(), "rel", np.abs(sg/sv-1).max(), "vs model", np.abs(sg-sm).max()/sv.max())

section("twisted vectors")
for n in [8, 64, 200, 512, 1024]:
    A = util.rand_matrix(n, n)
    _, al, be = util.oracle_bidiag(A)
    B = util.bidiag_dense(al, be)
    sv = np.linalg.svd(B, compute_uv=False)[::-1]
    X, Y = D.singular_vectors(al, be, sv)
    I = np.eye(n)
    print(n, "orthX", np.linalg.norm(X@X.T-I), "orthY", np.linalg.norm(Y@Y.T-I), "resid", np.linalg.norm(B - Y.T@np.diag(sv)@X)/np.linalg.norm(B), "eps*n", n*2.2e-16)

section("backtransform vs oracle")
for (m, n) in [(64,64),(200,200),(300,200)]:
    A = util.rand_matrix(m, n)
    Ao, al, be = util.oracle_bidiag(A)
    mnn = min(m,n)
    X = rng.standard_normal((mnn, n if m>=n else m+1)); Y = rng.standard_normal((mnn, mnn))
    U, V = D.backtransform(Ao, X, Y)
    # oracle, vector by vector
    AT = np.asfortranarray(Ao.T)
    Uo = np.zeros((m, mnn), order='F'); Vo = np.zeros((n, mnn), order='F')
    Xc = np.ascontiguousarray(X); Yc = np.ascontiguousarray(Y)
    for i in range(mnn):
        u = np.zeros(m); v = np.zeros(n)
        util.oracle().orc_apply_left(m, n, i, util.p(Ao), util.p(Yc), util.p(u))
        util.oracle().orc_apply_right(m, n, i, util.p(AT), util.p(Xc), util.p(v))
        Uo[:, i] = u; Vo[:, i] = v
    print((m,n), "U err", np.abs(U-Uo).max(), "V err", np.abs(V-Vo).max())

section("svd_gpu end to end")
for (m, n) in [(64,64),(512,512),(1024,1024),(700,500),(500,700),(2048,2048),(4096,4096)]:
    A = util.rand_matrix(m, n)
    t = time.time(); sg, U, V, Amod = D.svd_gpu(A); dt = time.time()-t
    ms = D.last_phase_ms()
    if m <= 2048:
        met = util.svd_metrics(A, sg, U, V)
        print((m,n), {k:(float('%.3g'%v) if not isinstance(v,bool) else v) for k,v in met.items()}, "eps*max", max(m,n)*2.2e-16)
    print((m,n), "wall", round(dt,3), "phase ms [h2d,bidiag,ddc,tw,bt,d2h,tot]", [round(x,2) for x in ms], flush=True)
    if (m,n)==(4096,4096):
        t = time.time(); sg2,_,_,_ = D.svd_gpu(A); print("second call wall", round(time.time()-t,3), [round(x,2) for x in D.last_phase_ms()])
        sv = np.linalg.svd(A, compute_uv=False)[::-1]
        print("4096 sigma abs/max", np.abs(sg-sv).max()/sv.max(), "rel", np.abs(sg/sv-1).max(), "orthU(sample)", np.linalg.norm(U[:, :256].T@U[:, :256]-np.eye(256)))
        print("4096 orthU", np.linalg.norm(U.T@U-np.eye(n)), "orthV", np.linalg.norm(V.T@V-np.eye(n)), "resid", np.linalg.norm(A-(U*sg)@V.T)/np.linalg.norm(A))
