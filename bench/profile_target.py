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
a = ap.parse_args()
n = a.n; m = a.m or n; mn = min(m, n)
L = D.lib()
dev = torch.device("cuda", 0)
torch.manual_seed(1)
A = torch.rand((n, m), dtype=torch.float64, device=dev) * 3 + 1       # column-major m x n
W = torch.empty_like(A)
sig = torch.empty(mn, dtype=torch.float64, device=dev)
U = torch.empty((mn, m), dtype=torch.float64, device=dev)
V = torch.empty((mn, n), dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(a.reps):
    W.copy_(A)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    L.svd_gpu_dev(m, n, W.data_ptr(), m, sig.data_ptr(), U.data_ptr(), m, V.data_ptr(), n, st)
    e1.record(); torch.cuda.synchronize()
    print("svd_gpu_dev %dx%d: %.2f ms, phases %s" % (m, n, e0.elapsed_time(e1), [round(x, 2) for x in D.last_phase_ms()]), flush=True)
