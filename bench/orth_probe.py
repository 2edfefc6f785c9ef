"""Where does ||U^T U - I||_F come from at large n?  (analysis aid, torch only as the calculator)
    python bench/orth_probe.py --n 16384
Prints the Frobenius norm, the share of the first off-diagonals, and the worst pairs with their relative gaps."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ddc_svd_b200 as D

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16384)
a = ap.parse_args()
n = a.n
L = D.lib()
dev = torch.device("cuda", 0)
torch.manual_seed(1)
A = torch.rand((n, n), dtype=torch.float64, device=dev) * 3 + 1
W = A.clone()
sig = torch.empty(n, dtype=torch.float64, device=dev)
U = torch.empty((n, n), dtype=torch.float64, device=dev)
V = torch.empty((n, n), dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
L.svd_gpu_dev(n, n, W.data_ptr(), n, sig.data_ptr(), U.data_ptr(), n, V.data_ptr(), n, st)
torch.cuda.synchronize()
print("phases", [round(x, 2) for x in D.last_phase_ms()])
del W
for name, Q in (("U", U), ("V", V)):
    G = Q @ Q.T
    G.diagonal().sub_(1.0)
    tot = torch.linalg.norm(G).item()
    print(f"{name}: ||G||_F {tot:.3e} = {tot / (2.2e-16 * n):.1f} eps n ; diag part {torch.linalg.norm(G.diagonal()).item():.3e}")
    for k in (1, 2, 3, 8):
        band = sum((torch.linalg.norm(G.diagonal(d)).item() ** 2) * 2 for d in range(1, k + 1)) ** 0.5
        print(f"   |i-j| <= {k}: {band:.3e}")
    Ga = G.abs()
    Ga.diagonal().zero_()
    vals, idx = torch.topk(Ga.flatten(), 12)
    for v, ix in zip(vals.tolist(), idx.tolist()):
        i, j = ix // n, ix % n
        if i < j:
            s_i, s_j = sig[i].item(), sig[j].item()
            print(f"   |g[{i},{j}]| = {v:.3e}  sigma {s_i:.6e} {s_j:.6e} relgap {(s_j - s_i) / s_j:.2e}")
    # without the 100 worst rows
    rowmax = Ga.max(dim=1).values
    print(f"   rows with an entry > 1e-12: {(rowmax > 1e-12).sum().item()}, > 1e-11: {(rowmax > 1e-11).sum().item()}, > 1e-10: {(rowmax > 1e-10).sum().item()}")
    del G, Ga
