"""Time the tcgen05 (int8 Ozaki) update against the FP64 DMMA GEMM on the back-transform's update shape:
C (M x N) -= A (M x 128) * B (128 x N).   python bench/ozaki_bench.py [M N]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ddc_svd_b200 as D
L = D.lib()
shapes = [(16384, 16384), (8192, 16384), (4096, 4096), (2048, 4096), (16384, 2048)]
if len(sys.argv) > 2:
    shapes = [(int(sys.argv[1]), int(sys.argv[2]))]
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
for M, N in shapes:
    K = 128
    A = torch.randn((K, M), dtype=torch.float64, device=dev) * 0.01      # column-major M x K
    B = torch.randn((N, K), dtype=torch.float64, device=dev) * 0.01      # column-major K x N
    C = torch.randn((N, M), dtype=torch.float64, device=dev)
    work = torch.empty(L.svdgpu_ozaki_workspace(M, N) // 8 + 64, dtype=torch.float64, device=dev)
    def t(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    t_oz = t(lambda: L.svdgpu_ozaki_update(M, N, -1.0, A.data_ptr(), M, B.data_ptr(), K, C.data_ptr(), M, work.data_ptr(), st))
    t_dm = t(lambda: L.svdgpu_dgemm(0, 0, M, N, K, -1.0, A.data_ptr(), M, B.data_ptr(), K, 1.0, C.data_ptr(), M, st))
    fl = 2.0 * M * N * K
    print(f"M={M} N={N}: ozaki(tcgen05 int8) {t_oz:.3f} ms = {fl/t_oz/1e9:.1f} TFLOP/s FP64-equivalent, C traffic {16.0*M*N/t_oz/1e6:.0f} GB/s | "
          f"DMMA {t_dm:.3f} ms = {fl/t_dm/1e9:.1f} TFLOP/s | speed-up {t_dm/t_oz:.2f}x", flush=True)
