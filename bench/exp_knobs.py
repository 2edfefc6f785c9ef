"""Knob sweep in ONE process (one box, one CUDA context): svd_gpu_dev on resident data under different
environment knobs / leading-dimension paddings, per-phase device times, with the size-independent
property checks of bench/profile_target.py after every configuration.

    python bench/exp_knobs.py --n 16384 --configs "base;ld=16;nb=64;nbbig=64:8192;tile=64"
A configuration is a comma-separated list of  ld=<pad doubles> | nb=<w> | nbbig=<w>:<min rows> | tile=64 | ws=0/1 | env:NAME=VALUE.
"""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ddc_svd_b200 as D

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--configs", default="base")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--values-only", action="store_true")
a = ap.parse_args()
n = m = a.n
L = D.lib()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(1)
A = torch.rand((n, m), dtype=torch.float64, device=dev) * 3 + 1       # column-major m x n, ld m
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
KNOB_ENVS = ["SVD_GPU_NB_BIG", "SVD_GPU_NB_BIG_MIN", "SVD_GPU_GEMM_TILE", "SVD_GPU_GEMM_WS", "SVD_GPU_TAIL"]


def padded(rows, ld):
    t = torch.zeros((rows, ld), dtype=torch.float64, device=dev)
    return t


def run(cfg):
    for e in KNOB_ENVS:
        os.environ.pop(e, None)
    extra = []
    pad, nb = 0, 32
    for tok in [t for t in cfg.split(",") if t and t != "base"]:
        if tok.startswith("env:"):
            k, v = tok[4:].split("=", 1); os.environ[k] = v; extra.append(k)
            continue
        k, v = tok.split("=")
        if k == "ld": pad = int(v)
        elif k == "nb": nb = int(v)
        elif k == "nbbig":
            w, mn_ = v.split(":"); os.environ["SVD_GPU_NB_BIG"] = w; os.environ["SVD_GPU_NB_BIG_MIN"] = mn_
        elif k == "tile": os.environ["SVD_GPU_GEMM_TILE"] = v
        elif k == "ws": os.environ["SVD_GPU_GEMM_WS"] = v
        elif k == "tail": os.environ["SVD_GPU_TAIL"] = v
        else: raise SystemExit("unknown knob " + tok)
    L.svd_gpu_set_option(b"nb", nb)
    ld = m + pad
    W = padded(n, ld); U = None; V = None
    sig = torch.empty(n, dtype=torch.float64, device=dev)
    if not a.values_only:
        U = padded(n, ld); V = padded(n, ld)
    best = None
    for _ in range(a.reps):
        W[:, :m].copy_(A)
        flush.zero_()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        L.svd_gpu_dev(m, n, W.data_ptr(), ld, sig.data_ptr(), U.data_ptr() if U is not None else None, ld,
                      V.data_ptr() if V is not None else None, ld, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ph = [round(x, 2) for x in D.last_phase_ms()]
        if best is None or ms < best[0]:
            best = (ms, ph)
    out = {"n": n, "cfg": cfg, "ms": round(best[0], 2), "bidiag": best[1][1], "ddc": best[1][2], "twisted": best[1][3],
           "backxf": best[1][4]}
    chk = abs((sig ** 2).sum().item() - (A ** 2).sum().item()) / (A ** 2).sum().item()
    out["sumsq"] = float("%.2e" % chk)
    if U is not None:
        k = min(n, 512)
        idx = torch.linspace(0, n - 1, k, device=dev).long()
        Us, Vs, ss = U[idx][:, :m], V[idx][:, :n], sig[idx]
        eye = torch.eye(k, dtype=torch.float64, device=dev)
        out["orthU"] = float("%.2e" % torch.linalg.norm(Us @ Us.T - eye).item())
        out["orthV"] = float("%.2e" % torch.linalg.norm(Vs @ Vs.T - eye).item())
        out["res"] = float("%.2e" % (torch.linalg.norm(Vs @ A - ss[:, None] * Us) / torch.linalg.norm(ss)).item())
    for e in extra:
        os.environ.pop(e, None)
    del W, U, V
    print(json.dumps(out), flush=True)


run("base")          # warm-up (arena, attributes); printed too
for c in a.configs.split(";"):
    run(c)
