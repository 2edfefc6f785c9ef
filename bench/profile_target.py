"""One svd_gpu_dev() call on resident data — the short command profiled under ncu.
    python bench/profile_target.py --n 8192 [--values-only]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ddc_svd_b200 as D

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--m", type=int, default=0)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--values-only", action="store_true")
ap.add_argument("--check", action="store_true")
a = ap.parse_args()
n = a.n; m = a.m or n; mn = min(m, n)
L = D.lib()
dev = torch.device("cuda", 0)
torch.manual_seed(1)
A = torch.rand((n, m), dtype=torch.float64, device=dev) * 3 + 1       # column-major m x n
W = torch.empty_like(A)
sig = torch.empty(mn, dtype=torch.float64, device=dev)
U = torch.empty((1 if a.values_only else mn, m), dtype=torch.float64, device=dev)
V = torch.empty((1 if a.values_only else mn, n), dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(a.reps):
    W.copy_(A)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    if a.values_only:
        L.svd_gpu_dev(m, n, W.data_ptr(), m, sig.data_ptr(), None, m, None, n, st)
    else:
        L.svd_gpu_dev(m, n, W.data_ptr(), m, sig.data_ptr(), U.data_ptr(), m, V.data_ptr(), n, st)
    e1.record(); torch.cuda.synchronize()
    print("svd_gpu_dev %dx%d: %.2f ms, phases %s" % (m, n, e0.elapsed_time(e1), [round(x, 2) for x in D.last_phase_ms()]), flush=True)

if a.check and not a.values_only:
    # size-independent properties on the device (torch only as the checker): orthogonality of a
    # column sample, residual of a column sample, checksum of checksums ||A||_F^2 = sum sigma^2
    k = min(mn, 512)
    idx = torch.linspace(0, mn - 1, k, device=dev).long()
    Us, Vs, ss = U[idx], V[idx], sig[idx]                    # rows = vectors
    eye = torch.eye(k, dtype=torch.float64, device=dev)
    oU = torch.linalg.norm(Us @ Us.T - eye).item(); oV = torch.linalg.norm(Vs @ Vs.T - eye).item()
    # A v_i = sigma_i u_i  (A column-major m x n  ==  A.T as a row-major (n, m) tensor)
    AV = Vs @ A                                              # (k, m): (A v_i)^T
    res = (torch.linalg.norm(AV - ss[:, None] * Us) / torch.linalg.norm(ss)).item()
    chk = abs((sig ** 2).sum().item() - (A ** 2).sum().item()) / (A ** 2).sum().item()
    print("check: orthU(sample %d) %.3e orthV %.3e |Av-su|/|s| %.3e  |sum s^2 - |A|^2|/|A|^2 %.3e  eps*max %.3e"
          % (k, oU, oV, res, chk, 2.2e-16 * max(m, n)), flush=True)
