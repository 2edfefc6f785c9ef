#!/usr/bin/env python
"""bench.py — svd_gpu() seconds on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W [--size 16384] [--impl reference]

A "step" is one full svd_gpu() of the workload matrix (n x n, FP64, uniform [1,4) from glibc rand() with the
default seed: bit for bit the matrix the reference's driver feeds, test-whole-svd.c:18-24,69-73; full U, Sigma, V).
BASELINE.json quotes the metric at n = 4096 and n = 16384: the headline is the north-star target 16384^2
(BASELINE.json configs[3], fits one GPU), and the 4096^2 line (configs[1]) rides along under "secondary"
(N = 1 only), measured the same way with the same K and W.
  value : seconds per step with the input already resident in HBM (svd_gpu_dev / svd_gpu_sharded_dev through the
          C ABI, timed with CUDA events on the launching stream, K steps bracketed by barrier + synchronize).
  e2e   : seconds per step through the reference-facing host-pointer call (svd_gpu() / svd_gpu_sharded()) with
          page-locked HOST buffers; host->device and device->host copies are inside the timed region.
          e2e_pageable: the same with malloc'd buffers, as the reference's callers pass them.
  phases: per-phase device milliseconds of one more step; at N > 1 rank 0's factorization phases, the slowest
          rank's vector phases, the exposed wait for the last panels, and the vector-phase speed-up against one
          more step run on rank 0 alone (same box, same build).
  check : the LAST step's result verified on the device (svd_gpu_check_dev: ||U^T U - I||_F, ||V^T V - I||_F,
          ||A - U S V^T||_F / ||A||_F, sum sigma^2 = ||A||_F^2, ascending) at every N; LAPACK too for n <= 4096.
  roofline     : the bidiagonalization (HBM bound): SURVEY 8(d) B_alg and the bytes the scheme really moves,
                 both over the measured phase time; single full-size pass probes; back-transform vs the DMMA
                 peak calibrated on this box (build/calib).
  cpu_baseline : the reference's own CPU path (oracle/_ref, else the oracle port) on bounded samples.
N > 1 (torchrun, one rank per GPU): rank 0 factorizes and broadcasts prepared compact-WY panels over NCCL while
it does, every rank solves / back-transforms its block of singular values ("scaling": "strong").
--impl reference : times the reference's CPU implementation on the box's host cores (rank 0 only).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum of ONE captured launch of the dominant kernel
# (ncu --set full, round 2 capture, profiles/r02_ncu_summary.md): fused_pass_kernel at step i=1500 of n=16384
# (algorithmic single-read bytes of that launch: 8*(16384-1500)*(16384-1501) = 1.772e9) and at step 1000 of n=4096
NCU_TRAFFIC = {16384: 1.780513e9 + 4.743680e6, 4096: 77.3888e6 + 1.939456e6}
NCU_TRAFFIC_NOTE = ("dram__bytes_read.sum + dram__bytes_write.sum of ONE fused_pass_kernel launch (ncu --set full): "
                    "n=16384 step 1500 (single-read algorithmic bytes of that launch 1.772e9, 263.3 us), "
                    "n=4096 step 1000 (76.7e6, 23.8 us); profiles/r02_ncu_summary.md")
METRIC = "svd_gpu seconds"
UNIT = "s"
EPS = float(np.finfo(np.float64).eps)
c_dp = ctypes.POINTER(ctypes.c_double)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def dmma_peak():
    """FP64 DMMA peak of THIS box from build/calib (bench/calib.cu, a few seconds); the committed calibration
    record of the pool otherwise."""
    exe = os.path.join(ROOT, "build", "calib")
    best, src = None, None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        vals = [json.loads(l)["tflops"] for l in out.splitlines() if '"dmma"' in l]
        if vals:
            best, src = max(vals), "build/calib run on this box in this job"
    except Exception:
        pass
    if best is None:
        try:
            rec = [json.loads(l) for l in open(os.path.join(ROOT, "profiles", "r01_calib_b200.jsonl"))]
            best = max(r["tflops"] for r in rec if r.get("probe") == "dmma")
            src = "profiles/r01_calib_b200.jsonl (bench/calib.cu on this pool's B200)"
        except Exception:
            best, src = 37.1, "spec (148 SMs x 128 flop/clk x 1.965 GHz)"
    return best, src


def bidiag_bytes(m, n, nb, fused=True, tail=True):
    """Algorithmic HBM bytes of the panel bidiagonalization (DESIGN.md 3.1 'bytes per unit').
    Per step: the fused pass reads the trailing block (m-i) x (n-i-1) once; the split passes read
    it twice; plus one read+write of the trailing block per panel for the deferred rank-2nb update.
    Returns (bytes our scheme has to move, SURVEY 8d B_alg = 12*S for the unblocked 3-transfer scheme)."""
    mn = min(m, n)
    i = np.arange(mn - 1, dtype=np.float64)
    t1 = (m - i) * (n - i - 1)
    t2 = np.where(i < n - 2, (m - i - 1) * (n - i - 1), 0.0)
    is_fused = (m - i >= 16) & (n - i - 1 >= 8) & (i < n - 2) if fused else np.zeros_like(i, dtype=bool)
    per_step = np.where(is_fused, t1, t1 + t2)
    i_tail = mn
    if tail:
        import ddc_svd_b200 as D
        i_tail = D.lib().svdgpu_bidiag_tail_start(m, n, nb, 148)      # the library's own planning rule
    reads = per_step[:i_tail].sum() + (float(m - i_tail) * (n - i_tail) if i_tail < mn else 0.0)
    ends = np.arange(nb - 1, min(mn - 1, i_tail), nb, dtype=np.float64)
    upd = ((m - ends - 1) * (n - ends - 1)).sum()
    return 8.0 * reads + 16.0 * upd, 12.0 * (t1 + t2).sum()


def backxf_flops(m, n):
    mn = min(m, n)
    j = np.arange(mn, dtype=np.float64)
    fu = 4.0 * mn * (m - j).sum()
    jr = np.arange(max(min(mn, n - 2), 0), dtype=np.float64)
    fv = 4.0 * mn * (n - jr - 1).sum()
    return fu + fv


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_input(m, n, seed=1):
    """The reference driver's matrix: uniform [1,4) FP64 from glibc rand() after srand(seed), fill order
    i = 0 .. m*n-1 into column-major storage (test-whole-svd.c:18-24,69-73; svdgpu_fill_rand is that loop in C).
    Returned as the (n, m) row-major array whose memory IS the column-major m x n matrix."""
    import ddc_svd_b200 as D
    buf = np.empty((n, m), dtype=np.float64)
    D.lib().svdgpu_fill_rand(buf.ctypes.data_as(c_dp), m * n, 1.0, 4.0, seed)
    return buf


# ------------------------------------------------------------------------------ the reference's CPU path
def _cpu_svd_seconds(ns, reps=1):
    """Seconds of the reference's svd_gpu.c:100-121 sequence on an n x n sample (oracle/_ref when it was built
    where the reference is mounted, else the oracle port) — the one place bench.py executes oracle/."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    ref = util.reference()
    A = np.asfortranarray(make_input(ns, ns).T)
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        if ref is not None:
            util.reference_svd(ref, A)
        else:
            util.oracle_svd(A)
        ts.append(time.perf_counter() - t)
    return float(np.mean(ts)), ("reference" if ref is not None else "port")


def cpu_extrapolated(n_target, n_small, n_big, reps_small=1):
    """Measure two sizes, extrapolate with the MEASURED growth exponent (the reference's serial bidiag_seq grows
    faster than n^3, BASELINE.md 2a: 11-12x per doubling)."""
    ncores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(ncores))
    t_small, kind = _cpu_svd_seconds(n_small, reps_small)
    t_big, _ = _cpu_svd_seconds(n_big, 1)
    p = float(np.log(t_big / t_small) / np.log(n_big / n_small))
    val = t_big * (n_target / n_big) ** p
    sample = (f"full SVD at n={n_small}: {t_small:.2f} s (mean of {reps_small}), at n={n_big}: {t_big:.2f} s (once) "
              f"=> measured growth n^{p:.2f}; extrapolated from n={n_big} to n={n_target} with that exponent "
              f"(16384^2 is not runnable on CPU: BASELINE.md estimates 1.5 days)")
    return val, {"value": val, "unit": UNIT, "cores": ncores, "kind": kind, "sample": sample,
                 "measured_s": {str(n_small): t_small, str(n_big): t_big}, "exponent": p}, t_small


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded: W + K runs at n = 1024 (a few seconds each; K capped so the arm stays within minutes) and ONE at
    # n = 2048 (~90 s: the serial bidiag_seq), then the measured exponent carries the figure to the config's size
    reps = max(1, min(args.steps, 6))
    for _ in range(min(args.warmup, 1)):
        _cpu_svd_seconds(512, 1)
    val, cb, t_small = cpu_extrapolated(args.n, 1024, 2048, reps)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": val * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.n), "m": args.n, "n": args.n,
                       "cpu_samples_n": [1024, 2048], "timed_runs_at_1024": reps},
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(n):
    tag = {4096: "BASELINE.json configs[1]", 16384: "BASELINE.json configs[3], the north-star target"}.get(n, "BASELINE.json configs[1] family")
    return f"square {n}x{n} random uniform[1,4) FP64 (test-whole-svd.c recipe), full U/Sigma/V ({tag})"


# ------------------------------------------------------------------------------ own arm
class Env:
    pass


def measure(E, n, steps, warmup, e2e_steps, want_probe, want_pageable):
    """One workload size on the current group of ranks: value, phases, check, e2e, roofline."""
    import torch
    import torch.distributed as dist
    import ddc_svd_b200 as D
    L, dev, world, rank = E.L, E.dev, E.world, E.rank
    m = mn = n
    nb = int(os.environ.get("SVD_GPU_NB", "32"))
    stream = torch.cuda.current_stream().cuda_stream
    blk, i0, ns = D.shard_range(mn, world, rank)

    A_host = make_input(m, n) if rank == 0 else None
    A_master = torch.empty((n, m), dtype=torch.float64, device=dev)       # column-major m x n == row-major (n, m)
    if rank == 0:
        A_master.copy_(torch.from_numpy(A_host))
    A_work = torch.empty_like(A_master) if rank == 0 else None
    sigma = torch.zeros(mn, dtype=torch.float64, device=dev)
    U = torch.zeros((blk, m), dtype=torch.float64, device=dev)            # this rank's columns (all of them at N = 1)
    V = torch.zeros((blk, n), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    p1 = lambda t: (ctypes.c_void_p * 1)(t.data_ptr())
    st1 = (ctypes.c_void_p * 1)(stream)

    def step_device():
        flush.zero_()                                   # L2 flush between steps (126 MB L2 < 256 MiB)
        if rank == 0:
            A_work.copy_(A_master)                      # svd_gpu destroys A: restore the resident input
        if world == 1:
            L.svd_gpu_dev(m, n, A_work.data_ptr(), m, sigma.data_ptr(), U.data_ptr(), m, V.data_ptr(), n, stream)
        else:
            L.svd_gpu_sharded_dev(E.group.h, m, n, A_work.data_ptr() if rank == 0 else None, m, p1(sigma),
                                  p1(U), m, p1(V), n, st1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM
    for _ in range(warmup):
        step_device()
    barrier()
    sampler = ClockSampler(E.local)
    if rank == 0:
        sampler.start()
    launches0 = L.svdgpu_launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step_device()
    e1.record()
    barrier()
    launches = L.svdgpu_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps
    res = {"ms_per_step": ms_per_step, "gpu_launches": int(launches), "clocks": clocks}

    # ---- per-phase device times of one more (untimed) step
    step_device()
    barrier()
    if world == 1:
        ph = D.last_phase_ms()
        res["phases"] = {"bidiag_ms": ph[1], "ddc_ms": ph[2], "twisted_ms": ph[3], "backtransform_ms": ph[4]}
    else:
        mine = torch.tensor(E.group.phase_ms(0), dtype=torch.float64, device=dev)
        allp = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allp, mine)
        allp = [x.cpu().numpy() for x in allp]
        r0 = allp[0]
        vec_n = float(r0[6] - r0[1] - r0[2])             # rank 0: everything after the dDC values
        res["phases"] = {"bidiag_ms": float(r0[1]), "ddc_ms": float(r0[2]),
                         "comm_wait_ms_rank0": float(r0[7]),
                         "twisted_ms_max": float(max(x[3] for x in allp)),
                         "backtransform_ms_max": float(max(x[4] for x in allp)),
                         "twisted_ms_by_rank": [round(float(x[3]), 3) for x in allp],
                         "backtransform_ms_by_rank": [round(float(x[4]), 3) for x in allp],
                         "vector_phases_ms": vec_n,
                         "note": "vector_phases_ms = rank 0's step minus its bidiag and dDC phases: exposed panel wait "
                                 "+ twisted + back-transform + the sigma all-gather (the part that shards)"}
    # ---- the last step's result, checked on the device (every N)
    res["check"] = check_last(E, m, n, A_master, sigma, U, V, A_host)
    if world > 1:
        # the same box, one more step on rank 0 alone: what the vector phases cost without sharding
        if rank == 0:
            U1 = torch.zeros((mn, m), dtype=torch.float64, device=dev)
            V1 = torch.zeros((mn, n), dtype=torch.float64, device=dev)
            flush.zero_(); A_work.copy_(A_master)
            L.svd_gpu_dev(m, n, A_work.data_ptr(), m, sigma.data_ptr(), U1.data_ptr(), m, V1.data_ptr(), n, stream)
            torch.cuda.synchronize()
            p1g = D.last_phase_ms()
            vec_1 = float(p1g[6] - p1g[1] - p1g[2])
            res["phases"]["vector_phases_ms_1gpu_same_box"] = vec_1
            res["phases"]["vector_phase_speedup"] = vec_1 / res["phases"]["vector_phases_ms"]
            res["phases"]["step_ms_1gpu_same_box"] = float(p1g[6])
            del U1, V1
        barrier()

    # ---- roofline of the dominant kernels (at N > 1: rank 0's factorization, the slowest rank's back-transform)
    if True:
        ph = [0, res["phases"]["bidiag_ms"], 0, 0,
              res["phases"]["backtransform_ms"] if world == 1 else res["phases"]["backtransform_ms_max"]]
        pk, which = peaks()
        fused_on = os.environ.get("SVD_GPU_FUSED", "1") != "0"
        tail_on = os.environ.get("SVD_GPU_TAIL", "1") != "0"
        b_own, b_survey = bidiag_bytes(m, n, nb, fused_on, tail_on)
        probe = {}
        if want_probe and world == 1:
            wbytes = L.svdgpu_bidiag_workspace(m, n, m)
            work = torch.zeros(wbytes // 8 + 8, dtype=torch.float64, device=dev)
            scratchA = A_master.clone()            # the fused probe writes a reflector into column 0
            for wh, nm in ((0, "gemvT"), (1, "gemvN"), (2, "fused")):
                src = scratchA if wh == 2 else A_master
                for _ in range(3):
                    L.svdgpu_bidiag_pass_probe(m, n, src.data_ptr(), m, work.data_ptr(), wh, stream)
                q0 = torch.cuda.Event(enable_timing=True); q1 = torch.cuda.Event(enable_timing=True)
                reps, tot = 10, 0.0
                for _ in range(reps):
                    flush.zero_()
                    q0.record()
                    L.svdgpu_bidiag_pass_probe(m, n, src.data_ptr(), m, work.data_ptr(), wh, stream)
                    q1.record()
                    torch.cuda.synchronize()
                    tot += q0.elapsed_time(q1)
                probe[nm + "_full_pass_gbs"] = 8.0 * m * (n - 1 if wh == 2 else n) / (tot / reps * 1e-3) / 1e9
                if wh == 2:
                    probe["fused_full_pass_us"] = tot / reps * 1e3
                    probe["fused_full_pass_frac_of_peak"] = probe[nm + "_full_pass_gbs"] / pk["hbm_gbs"]
            del scratchA, work
        t_bd = ph[1] * 1e-3
        ach_survey = b_survey / t_bd / 1e9          # SURVEY.md 8(d) definition: B_alg = 8 * 1.5 * S
        ach_own = b_own / t_bd / 1e9                # bytes our scheme actually has to move
        bt = backxf_flops(m, n) / world / (ph[4] * 1e-3) / 1e12       # per GPU
        res["roofline"] = {
            "bound": "hbm",
            "kernel": ("fused_pass_kernel (single-read pass) + finish_xf + panel GEMM (dgemm_ws_kernel); bidiag_tail_kernel "
                       "once the trailing block fits on chip" if fused_on else "gemvT_kernel + gemvN_kernel + finish_y/x + panel GEMM"),
            "achieved": ach_survey, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach_survey / pk["hbm_gbs"],
            "own_scheme_achieved": ach_own, "own_scheme_frac": ach_own / pk["hbm_gbs"],
            "own_scheme_frac_of_8000_nominal": ach_own / 8000.0,
            "frac_of_8000_nominal": ach_survey / 8000.0,
            "traffic": NCU_TRAFFIC.get(n), "traffic_note": NCU_TRAFFIC_NOTE,
            "definition": "achieved = SURVEY 8(d) B_alg (12*S bytes: unblocked 3-transfer scheme) / bidiag phase time from CUDA "
                          "events (all bidiag launches); it exceeds the peak because the panel-deferred single-read scheme moves "
                          "~2.8x fewer bytes than B_alg assumes. own_scheme_* = the bytes this scheme really has to move "
                          "(one read of the trailing block per step + one read+write per panel) / the same time: the figure "
                          "that says how close the phase runs to the HBM roofline",
            "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs, copy); bench/calib.cu read stream: 7200 GB/s",
            "algorithmic_bytes": b_survey, "own_scheme_bytes": b_own, "phase_ms": ph[1],
            "fused_pass": fused_on, "on_chip_tail": tail_on, **probe,
            "backtransform_tflops": bt, "backtransform_frac_of_dmma_peak": bt / E.dmma[0],
            "dmma_peak_tflops": E.dmma[0], "dmma_peak_source": E.dmma[1]}

    # ---- e2e: the reference-facing call with HOST buffers, copies inside the timed region
    res["e2e"] = None
    if e2e_steps > 0:
        nbA = m * n * 8
        pinA = L.svdgpu_host_alloc(nbA) if rank == 0 else None
        pinS = L.svdgpu_host_alloc(mn * 8) if rank == 0 else None
        pinU = L.svdgpu_host_alloc(max(ns, 1) * m * 8)
        pinV = L.svdgpu_host_alloc(max(ns, 1) * n * 8)
        if rank == 0:
            hA = np.ctypeslib.as_array(ctypes.cast(pinA, c_dp), shape=(n * m,))
            src = A_host.reshape(-1)
        ub = (ctypes.c_void_p * 1)(pinU); vb = (ctypes.c_void_p * 1)(pinV)
        grp = E.group.h if world > 1 else None

        def step_e2e():
            if rank == 0:
                hA[:] = src                                   # refill the (destroyed) host input; not GPU work
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            if world == 1:      # the drop-in entry point itself
                L.svd_gpu(m, n, ctypes.cast(pinA, c_dp), ctypes.cast(pinS, c_dp), ctypes.cast(pinU, c_dp),
                          ctypes.cast(pinV, c_dp))
            else:               # what svd_gpu() runs on a group, one process per GPU
                L.svd_gpu_sharded(grp, m, n, ctypes.cast(pinA, c_dp) if rank == 0 else None,
                                  ctypes.cast(pinS, c_dp) if rank == 0 else None, ub, vb)
            return time.perf_counter() - t0
        for _ in range(max(1, min(2, warmup // 2))):
            step_e2e()
        ts = [step_e2e() for _ in range(e2e_steps)]
        tt = torch.tensor([float(np.mean(ts))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        res["e2e"] = {"value": float(tt.item()), "unit": UNIT, "steps": e2e_steps, "h2d_bytes_per_step": nbA,
                      "d2h_bytes_per_step": nbA + mn * 8 + m * mn * 8 + n * mn * 8,
                      "buffers": "page-locked (svdgpu_host_alloc); every rank downloads its own column block",
                      "phase_ms_last_rank0": dict(zip(["h2d", "bidiag", "ddc", "twisted", "backtransform", "d2h_tail", "total"],
                                                      [round(x, 3) for x in (E.group.phase_ms(0)[:7] if world > 1 else D.last_phase_ms())]))}
        for q in (pinA, pinS, pinU, pinV):
            if q:
                L.svdgpu_host_free(q)
        if want_pageable and world == 1:
            # malloc'd buffers, the way the reference's callers hand them over (test-whole-svd.c:44-66)
            out = {}
            Ah = np.empty((n, m)); sg = np.empty(mn); Uh = np.empty((mn, m)); Vh = np.empty((mn, n))
            for reg in (1, 0):
                D.set_option("host_register", reg)
                ts = []
                for _ in range(2):
                    Ah[:] = A_host
                    t0 = time.perf_counter()
                    L.svd_gpu(m, n, Ah.ctypes.data_as(c_dp), sg.ctypes.data_as(c_dp), Uh.ctypes.data_as(c_dp),
                              Vh.ctypes.data_as(c_dp))
                    ts.append(time.perf_counter() - t0)
                out["page_locked_for_the_call" if reg else "plain_pageable_copies"] = float(min(ts))
            D.set_option("host_register", 0)
            res["e2e_pageable"] = {"unit": UNIT, "value": out["plain_pageable_copies"], **out,
                                   "note": "malloc'd A/U/V (numpy), best of 2 calls each: the library's default (plain pageable "
                                           "copies) and with svd_gpu_set_option('host_register', 1) (cudaHostRegister + "
                                           "unregister inside the timed region)"}
            del Ah, Uh, Vh
    del A_master, A_work, U, V, flush
    torch.cuda.empty_cache()
    return res


def check_last(E, m, n, A_master, sigma, U, V, A_host):
    """The last step's result on the device: every rank checks its own block (needs the original matrix, which is
    broadcast for this purpose only), rank 0 checks the whole SVD after gathering the blocks; LAPACK for n <= 4096."""
    import torch
    import torch.distributed as dist
    import ddc_svd_b200 as D
    L, dev, world, rank = E.L, E.dev, E.world, E.rank
    mn = min(m, n)
    blk, i0, ns = D.shard_range(mn, world, rank)
    bound = 100 * EPS * max(m, n)
    out = np.zeros(6)
    stream = torch.cuda.current_stream().cuda_stream
    chk = {}
    if world > 1:
        dist.broadcast(A_master, 0)
        if ns > 0:
            L.svd_gpu_check_dev(m, n, A_master.data_ptr(), m, sigma[i0:].data_ptr(), U.data_ptr(), m, V.data_ptr(), n, ns,
                                out.ctypes.data_as(c_dp), stream)
        tb = torch.tensor(out[:3], dtype=torch.float64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        chk["blocks_max"] = {"orthU": float(tb[0]), "orthV": float(tb[1]), "resid_AV_minus_US": float(tb[2])}
        Uf = torch.zeros((world * blk, m), dtype=torch.float64, device=dev)
        Vf = torch.zeros((world * blk, n), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(Uf, U)
        dist.all_gather_into_tensor(Vf, V)
    else:
        Uf, Vf = U, V
    if rank == 0:
        L.svd_gpu_check_dev(m, n, A_master.data_ptr(), m, sigma.data_ptr(), Uf.data_ptr(), m, Vf.data_ptr(), n, mn,
                            out.ctypes.data_as(c_dp), stream)
        chk.update({"orthU": out[0], "orthV": out[1], "resid": out[2], "checksum": out[3], "ascending": bool(out[5] == 1.0),
                    "bound_100_eps_n": bound, "in_units_of_eps_n": [round(float(x / (EPS * max(m, n))), 2) for x in out[:3]],
                    "ok": bool(out[5] == 1.0 and max(out[0], out[1], out[2]) <= bound and out[3] <= bound),
                    "how": "svd_gpu_check_dev on the device (the reference driver's dormant check, test-whole-svd.c:81-96)"})
        if n <= 4096:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import util
            sg, Uh, Vh = sigma.cpu().numpy(), Uf[:mn].cpu().numpy().T, Vf[:mn].cpu().numpy().T
            lap = util.svd_metrics(np.asfortranarray(A_host.T), sg, Uh, Vh)
            chk["vs_lapack"] = lap
            chk["ok"] = bool(chk["ok"] and lap["sigma_abs_over_max"] <= 10 * EPS * n)
    if world > 1:
        del Uf, Vf
        dist.barrier()
    return chk if rank == 0 else None


def run_own(args):
    import torch
    import torch.distributed as dist
    import ddc_svd_b200 as D

    E = Env()
    E.world = int(os.environ.get("WORLD_SIZE", "1"))
    E.rank = int(os.environ.get("RANK", "0"))
    E.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(E.local)
    E.L = D.lib()
    E.L.svdgpu_set_device(E.local)
    E.dev = torch.device("cuda", E.local)
    E.group = None
    if E.world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner to stdout at the VERSION and WARN levels
        # and its log at INFO; send whatever is asked for to stderr, and ask for nothing by default
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN", ""):
            os.environ.pop("NCCL_DEBUG", None)
        else:
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=E.dev)
        # the library's own communicator (one rank per process): the 128-byte id travels over torch.distributed
        idt = torch.zeros(128, dtype=torch.uint8, device=E.dev)
        if E.rank == 0:
            idt.copy_(torch.frombuffer(bytearray(D.Group.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        E.group = D.Group.rank(E.world, E.rank, bytes(idt.cpu().numpy().tobytes()))
    E.dmma = dmma_peak() if E.rank == 0 else (37.1, "")

    n = args.n
    big = n >= 8192
    e2e_steps = 0 if args.no_e2e else (min(args.steps, 5) if big else args.steps)
    head = measure(E, n, args.steps, args.warmup, e2e_steps, want_probe=True, want_pageable=not args.no_e2e)
    secondary = None
    if E.world == 1 and n != 4096 and not args.no_secondary:
        s = measure(E, 4096, args.steps, args.warmup, 0 if args.no_e2e else args.steps, want_probe=True,
                    want_pageable=not args.no_e2e)
        secondary = {"config": {"workload": workload_name(4096), "m": 4096, "n": 4096},
                     "value": s["ms_per_step"] * 1e-3, "unit": UNIT, "ms_per_step": s["ms_per_step"], "steps": args.steps,
                     "warmup": args.warmup, "e2e": s["e2e"], "e2e_pageable": s.get("e2e_pageable"), "phases": s["phases"],
                     "roofline": s["roofline"], "check": s["check"], "gpu_launches": s["gpu_launches"]}
    if E.rank == 0:
        nb = int(os.environ.get("SVD_GPU_NB", "32"))
        line = {"metric": METRIC, "value": head["ms_per_step"] * 1e-3, "unit": UNIT, "n_gpus": E.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(n), "m": n, "n": n, "panel_nb": nb,
                           "l2": "flushed by a 256 MiB memset between steps (inside the timed region)",
                           "parallelism": "1 GPU" if E.world == 1 else
                           f"bidiag+dDC on rank 0 (WY panels broadcast over NCCL while it runs), twisted vectors + "
                           f"back-transform sharded by singular-value blocks over {E.world} ranks"},
                "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": head["clocks"],
                "phases": head["phases"], "check": head["check"]}
        if "e2e_pageable" in head:
            line["e2e_pageable"] = head["e2e_pageable"]
        if "roofline" in head:
            line["roofline"] = head["roofline"]
        if secondary:
            line["secondary"] = secondary
        if E.world == 1 and not args.no_cpu:
            # own arm: two short samples (about 5 s + 20 s of CPU work) and the measured exponent
            _, cb, _ = cpu_extrapolated(n, 1024, 1536, 1)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if E.world > 1:
        E.group.destroy()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # (--size, not --n: torchrun's own parser treats a bare --n as an ambiguous abbreviation)
    ap.add_argument("--size", "--n", dest="n", type=int, default=int(os.environ.get("SVD_BENCH_N", "16384")))
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the 4096^2 line that rides along at N = 1")
    ap.add_argument("--no-e2e", action="store_true",
                    help="skip the end-to-end legs (single-shot runs of the largest configs; e2e is then null)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
