"""BASELINE.json configs that shard (C3 tall-skinny 65536x4096 over 2/4/8 GPUs, C5 32768^2 over 8 GPUs) through the
library's own group API, ONE process driving N GPUs (svdgpu_group_create_local), device-resident data:

    python bench/configs_multi.py --m 65536 --n 4096 --gpus 2 4 8
    python bench/configs_multi.py --m 32768 --n 32768 --gpus 8

Per run: seconds per svd_gpu_sharded_dev step (CUDA events on rank 0's stream, after a warm-up), per-rank phase
times, and the result of every rank's block checked on its own GPU with svd_gpu_check_dev (orthogonality of the
block, ||A V_b - U_b S_b||_F / ||A||_F) plus ascending order and sum sigma^2 = ||A||_F^2.  One JSON line per run."""
import argparse, ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddc_svd_b200 as D

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, required=True)
ap.add_argument("--n", type=int, required=True)
ap.add_argument("--gpus", type=int, nargs="+", default=[2])
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
L = D.lib()
m, n = a.m, a.n
mn = min(m, n)
EPS = 2.220446049250313e-16
c_dp = ctypes.POINTER(ctypes.c_double)
A = np.empty((n, m))
L.svdgpu_fill_rand(A.ctypes.data_as(c_dp), m * n, 1.0, 4.0, 1)
normA2 = float(np.einsum("ij,ij->", A, A))
have = L.svdgpu_device_count()

for world in a.gpus:
    if world > have:
        print(json.dumps({"config": f"{m}x{n}", "n_gpus": world, "skipped": f"only {have} GPUs visible"}), flush=True)
        continue
    g = D.Group.local(world) if world > 1 else D.Group.local(1)
    blk, _, _ = D.shard_range(mn, world, 0)
    dA0, dA, dS, dU, dV, streams = [], None, [], [], [], []
    for r in range(world):
        L.svdgpu_set_device(r)
        d0 = L.svdgpu_malloc(8 * m * n); dA0.append(d0)
        L.svdgpu_h2d(d0, A.ctypes.data_as(c_dp), 8 * m * n, None)
        dS.append(L.svdgpu_malloc(8 * mn))
        dU.append(L.svdgpu_malloc(8 * m * blk)); dV.append(L.svdgpu_malloc(8 * n * blk))
        streams.append(L.svdgpu_stream_create())
        if r == 0:
            dA = L.svdgpu_malloc(8 * m * n)
    arr = lambda xs: (ctypes.c_void_p * world)(*xs)
    L.svdgpu_set_device(0)
    e0, e1 = L.svdgpu_event_create(), L.svdgpu_event_create()
    times = []
    for step in range(a.steps + 1):
        L.svdgpu_set_device(0)
        L.svdgpu_d2d(dA, dA0[0], 8 * m * n, streams[0])
        L.svdgpu_event_record(e0, streams[0])
        L.svd_gpu_sharded_dev(g.h, m, n, dA, m, arr(dS), arr(dU), m, arr(dV), n, arr(streams))
        L.svdgpu_set_device(0)
        L.svdgpu_event_record(e1, streams[0])
        for r in range(world):
            L.svdgpu_set_device(r); L.svdgpu_stream_sync(streams[r])
        L.svdgpu_set_device(0)
        if step > 0:
            times.append(L.svdgpu_event_elapsed_ms(e0, e1))
    phases = [g.phase_ms(r) for r in range(world)]
    chk = []
    sig = np.zeros(mn)
    L.svdgpu_set_device(0)
    L.svdgpu_d2h(sig.ctypes.data_as(c_dp), dS[0], 8 * mn, None); L.svdgpu_stream_sync(None)
    for r in range(world):
        _, i0, ns = D.shard_range(mn, world, r)
        out = np.zeros(6)
        if ns > 0:
            L.svdgpu_set_device(r)
            L.svd_gpu_check_dev(m, n, dA0[r], m, ctypes.c_void_p(dS[r] + 8 * i0), dU[r], m, dV[r], n, ns, out.ctypes.data_as(c_dp), None)
        chk.append(out)
    chk = np.array(chk)
    bound = 100 * EPS * max(m, n)
    r0 = phases[0]
    line = {"config": f"{m}x{n} full U/Sigma/V, one process driving {world} GPU(s)", "n_gpus": world,
            "seconds": float(np.mean(times)) * 1e-3, "steps": a.steps,
            "rank0_ms": {"factorization": round(r0[1], 2), "ddc": round(r0[2], 2), "panel_wait": round(r0[7], 3),
                         "twisted": round(r0[3], 2), "backtransform": round(r0[4], 2)},
            "twisted_ms_by_rank": [round(p[3], 2) for p in phases], "backtransform_ms_by_rank": [round(p[4], 2) for p in phases],
            "vector_phases_ms": round(r0[6] - r0[1] - r0[2], 2),
            "check": {"orthU_block_max": float(chk[:, 0].max()), "orthV_block_max": float(chk[:, 1].max()),
                      "resid_AV_minus_US_max": float(chk[:, 2].max()), "ascending": bool(np.all(np.diff(sig) >= 0)),
                      "checksum": abs(float(np.dot(sig, sig)) - normA2) / normA2, "bound_100_eps_max": bound,
                      "ok": bool(chk[:, :3].max() <= bound and np.all(np.diff(sig) >= 0)
                                 and abs(float(np.dot(sig, sig)) - normA2) <= bound * normA2)}}
    print(json.dumps(line), flush=True)
    for r in range(world):
        L.svdgpu_set_device(r)
        for d in (dA0[r], dS[r], dU[r], dV[r]):
            L.svdgpu_free(d)
        L.svdgpu_stream_destroy(streams[r])
    L.svdgpu_set_device(0)
    L.svdgpu_free(dA)
    g.destroy()
    D.set_option("release", 0)
