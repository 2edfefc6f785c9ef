#!/usr/bin/env python
"""bench.py — svd_gpu() seconds on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W [--n 4096] [--impl reference]

A "step" is one full svd_gpu() of the workload matrix (n x n, FP64, uniform [1,4) synthetic,
the reference driver's input distribution; full U, Sigma, V).
  value : seconds per step with the input already resident in HBM (svd_gpu_dev through the C ABI,
          timed with CUDA events on the launching stream, K steps bracketed by barrier + synchronize).
  e2e   : seconds per step through the reference-facing call svd_gpu(m,n,A,sigma,U,V) with pinned HOST
          buffers; host->device and device->host copies are inside the timed region.
  roofline     : the bidiagonalization (fused single-read pass + finish + panel GEMM, on-chip tail), HBM bound;
                 the back-transform's DMMA rate and single full-size pass probes ride along.
  cpu_baseline : the reference's own CPU path (oracle/_ref, else the oracle port) on a bounded sample.
N > 1 (torchrun): bidiagonalization + dDC on rank 0, NCCL broadcast of reflectors / bidiagonal /
singular values, twisted vectors + back-transform sharded by singular-value blocks, NCCL all-gather of
the U / V column blocks ("scaling": "strong" — total work is fixed).
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
# (ncu --set full, profiles/r01_ncu_summary_v3.md): fused_pass_kernel at step i=1500 of n=16384
# (algorithmic single-read bytes of that launch: 8*(16384-1500)*(16384-1501) = 1.772e9)
NCU_TRAFFIC = {16384: 1.780507e9 + 5.105664e6, 4096: 77.380864e6 + 1.881856e6}
NCU_TRAFFIC_NOTE = ("dram__bytes_read.sum + dram__bytes_write.sum of ONE fused_pass_kernel launch (ncu --set full): "
                    "n=16384 step 1500 (single-read algorithmic bytes of that launch 1.772e9), "
                    "n=4096 step 1000 (76.7e6); profiles/r01_ncu_summary_v3.md")
METRIC = "svd_gpu seconds"
UNIT = "s"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def bidiag_bytes(m, n, nb, fused=True, tail=True):
    """Algorithmic HBM bytes of the panel bidiagonalization (DESIGN.md 3.1 'bytes per unit').
    Per step: the fused pass reads the trailing block (m-i) x (n-i-1) once; the split passes read
    it twice ((m-i)(n-i-1) for the column dots, (m-i-1)(n-i-1) for the row dots); plus one read+write
    of the trailing block per panel for the deferred rank-2nb update.  The fused pass is used while
    the trailing block has >= 16 rows and >= 8 columns (bidiag.cu:plan_fused).  Also returns the
    survey's figure 12*S for the unblocked 3-transfer scheme."""
    mn = min(m, n)
    i = np.arange(mn - 1, dtype=np.float64)
    t1 = (m - i) * (n - i - 1)
    t2 = np.where(i < n - 2, (m - i - 1) * (n - i - 1), 0.0)
    is_fused = (m - i >= 16) & (n - i - 1 >= 8) & (i < n - 2) if fused else np.zeros_like(i, dtype=bool)
    per_step = np.where(is_fused, t1, t1 + t2)
    # on-chip tail (bidiag_tail.cuh): from the first panel boundary where the trailing block fits the
    # SMs' shared memory, the rest is ONE read of that block (bidiag.cu:tail_fits)
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


def make_input(n, m=None):
    """Synthetic workload: uniform [1,4) FP64, column-major (test-whole-svd.c:18-24 distribution;
    numpy's generator is used above 2048 to keep set-up time out of the run)."""
    m = n if m is None else m
    rng = np.random.default_rng(1)
    return np.asfortranarray(rng.uniform(1.0, 4.0, size=(n, m)).T)


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    ncores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(ncores))
    ref = util.reference()
    kind = "reference" if ref is not None else "port"
    ns = args.cpu_n
    A = make_input(ns)

    def one():
        t = time.perf_counter()
        if ref is not None:
            util.reference_svd(ref, A)         # svd_gpu.c:100-121 sequence on the reference build
        else:
            util.oracle_svd(A)
        return time.perf_counter() - t

    for _ in range(args.warmup):
        one()
    ts = [one() for _ in range(args.steps)]
    t_sample = float(np.mean(ts))
    scale = (args.n / ns) ** 3
    val = t_sample * scale
    sample = (f"n={ns} full SVD measured {t_sample:.3f} s/step, scaled to n={args.n} by (n/{ns})^3 "
              f"(flop ratio; the reference's real growth is steeper, BASELINE.md 2a)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_sample * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"square {args.n}x{args.n} full U/Sigma/V (BASELINE.json configs[1] family)",
                       "cpu_sample_n": ns},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ own arm
def cpu_baseline(args):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    ncores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(ncores))
    ref = util.reference()
    ns = args.cpu_n
    A = make_input(ns)
    t = time.perf_counter()
    if ref is not None:
        util.reference_svd(ref, A)
        kind = "reference"
    else:
        util.oracle_svd(A)
        kind = "port"
    dt = time.perf_counter() - t
    scale = (args.n / ns) ** 3
    return {"value": dt * scale, "unit": UNIT, "cores": ncores, "kind": kind,
            "sample": f"n={ns} full SVD, {dt:.2f} s measured once, scaled by (n/{ns})^3 to n={args.n}"}


def run_own(args):
    import torch
    import torch.distributed as dist
    import ddc_svd_b200 as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    L = D.lib()
    L.svdgpu_set_device(local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at VERSION/INFO level
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = m = args.n
    mn = n
    nb = int(os.environ.get("SVD_GPU_NB", "32"))
    stream = torch.cuda.current_stream().cuda_stream

    A_host = make_input(n) if rank == 0 else None
    # column-major m x n == row-major (n, m) tensor
    A_master = torch.empty((n, m), dtype=torch.float64, device=dev)
    if rank == 0:
        A_master.copy_(torch.from_numpy(np.ascontiguousarray(A_host.T)))
    A_work = torch.empty_like(A_master)
    sigma = torch.empty(mn, dtype=torch.float64, device=dev)
    U = torch.empty((mn, m), dtype=torch.float64, device=dev)
    V = torch.empty((mn, n), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # shards for N > 1: contiguous singular-value blocks
    blk = (mn + world - 1) // world
    i0 = min(rank * blk, mn)
    ns = max(0, min(blk, mn - i0))
    if world > 1:
        alpha = torch.zeros(mn, dtype=torch.float64, device=dev)
        beta = torch.zeros(mn + 1, dtype=torch.float64, device=dev)
        sig_all = torch.zeros(mn, dtype=torch.float64, device=dev)
        Ublk = torch.zeros((blk, m), dtype=torch.float64, device=dev)
        Vblk = torch.zeros((blk, n), dtype=torch.float64, device=dev)
        Ufull = torch.empty((world * blk, m), dtype=torch.float64, device=dev)
        Vfull = torch.empty((world * blk, n), dtype=torch.float64, device=dev)
        sig_blk = torch.zeros(blk, dtype=torch.float64, device=dev)
        sig_full = torch.empty(world * blk, dtype=torch.float64, device=dev)

    def step_device():
        flush.zero_()                                   # L2 flush between steps (126 MB L2 < 256 MiB)
        A_work.copy_(A_master)                          # svd_gpu destroys A: restore the resident input
        if world == 1:
            L.svd_gpu_dev(m, n, A_work.data_ptr(), m, sigma.data_ptr(), U.data_ptr(), m, V.data_ptr(), n, stream)
        else:
            if rank == 0:
                L.svd_gpu_values_dev(m, n, A_work.data_ptr(), m, alpha.data_ptr(), beta.data_ptr(),
                                     sig_all.data_ptr(), stream)
            # "all-gather the bidiagonal": reflectors + alpha/beta/sigma from the bidiag GPU
            dist.broadcast(A_work, 0)
            dist.broadcast(alpha, 0)
            dist.broadcast(beta, 0)
            dist.broadcast(sig_all, 0)
            if ns > 0:
                L.svd_gpu_vectors_dev(m, n, A_work.data_ptr(), m, alpha.data_ptr(), beta.data_ptr(),
                                      sig_all.data_ptr(), i0, ns, Ublk.data_ptr(), m, Vblk.data_ptr(), n,
                                      sig_blk.data_ptr(), stream)
            dist.all_gather_into_tensor(Ufull, Ublk)
            dist.all_gather_into_tensor(Vfull, Vblk)
            dist.all_gather_into_tensor(sig_full, sig_blk)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.svdgpu_launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    launches = L.svdgpu_launch_count() - launches0
    ms_dev = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps

    # ---- per-phase device times (one extra, untimed step) and the roofline of the dominant kernels
    phase = None
    roof = None
    if world == 1:
        step_device()
        torch.cuda.synchronize()
        ph = D.last_phase_ms()
        phase = {"bidiag_ms": ph[1], "ddc_ms": ph[2], "twisted_ms": ph[3], "backtransform_ms": ph[4]}
        pk, which = peaks()
        fused_on = os.environ.get("SVD_GPU_FUSED", "1") != "0"
        tail_on = os.environ.get("SVD_GPU_TAIL", "1") != "0"
        b_alg, b_survey = bidiag_bytes(m, n, nb, fused_on, tail_on)
        ach = b_alg / (ph[1] * 1e-3) / 1e9
        # single full-size passes of the two streaming kernels, timed alone
        wbytes = L.svdgpu_bidiag_workspace(m, n, m)
        work = torch.zeros(wbytes // 8 + 8, dtype=torch.float64, device=dev)
        probe = {}
        scratchA = A_master.clone()            # the fused probe writes a reflector into column 0
        for wh, nm in ((0, "gemvT"), (1, "gemvN"), (2, "fused")):
            src = scratchA if wh == 2 else A_master
            for _ in range(3):
                L.svdgpu_bidiag_pass_probe(m, n, src.data_ptr(), m, work.data_ptr(), wh, stream)
            p0 = torch.cuda.Event(enable_timing=True); p1 = torch.cuda.Event(enable_timing=True)
            reps = 10
            tot = 0.0
            for _ in range(reps):
                flush.zero_()
                p0.record()
                L.svdgpu_bidiag_pass_probe(m, n, src.data_ptr(), m, work.data_ptr(), wh, stream)
                p1.record()
                torch.cuda.synchronize()
                tot += p0.elapsed_time(p1)
            probe[nm + "_full_pass_gbs"] = 8.0 * m * (n - 1 if wh == 2 else n) / (tot / reps * 1e-3) / 1e9
            if wh == 2:
                probe["fused_full_pass_us"] = tot / reps * 1e3
                probe["fused_full_pass_frac_of_peak"] = probe[nm + "_full_pass_gbs"] / pk["hbm_gbs"]
        del scratchA
        t_bd = ph[1] * 1e-3
        ach_survey = b_survey / t_bd / 1e9          # SURVEY.md 8(d) definition: B_alg = 8 * 1.5 * S
        ach_own = b_alg / t_bd / 1e9                # bytes our scheme actually has to move
        roof = {"bound": "hbm",
                "kernel": ("fused_pass_kernel (single-read pass) + finish_xf + panel GEMM (dgemm_ws_kernel); bidiag_tail_kernel "
                           "once the trailing block fits on chip"
                           if fused_on else "gemvT_kernel + gemvN_kernel + finish_y/x + panel GEMM"),
                "achieved": ach_survey, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach_survey / pk["hbm_gbs"],
                "frac_of_8000_nominal": ach_survey / 8000.0,
                "traffic": NCU_TRAFFIC.get(n), "traffic_note": NCU_TRAFFIC_NOTE,
                "definition": "achieved = SURVEY 8(d) B_alg (12*S bytes: unblocked 3-transfer scheme) / bidiag phase "
                              "time from CUDA events (all bidiag launches); can exceed 1.0 because the panel-deferred "
                              "fused scheme moves fewer bytes than B_alg assumes - see own_scheme_*",
                "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs, copy); bench/calib.cu read stream: 7200 GB/s",
                "algorithmic_bytes": b_survey, "phase_ms": ph[1], "fused_pass": fused_on, "on_chip_tail": tail_on,
                "own_scheme_bytes": b_alg, "own_scheme_achieved": ach_own, "own_scheme_frac": ach_own / pk["hbm_gbs"],
                **probe,
                "backtransform_tflops": backxf_flops(m, n) / (ph[4] * 1e-3) / 1e12,
                "backtransform_frac_of_dmma_peak": backxf_flops(m, n) / (ph[4] * 1e-3) / 1e12 / 37.1,
                "dmma_peak_tflops": 37.1, "dmma_peak_source": "bench/calib.cu on this pool's B200"}
        del work

    # ---- e2e: the reference-facing call with pinned host buffers, copies inside the timed region
    e2e = None
    if args.no_e2e:
        pass
    elif world == 1:
        nbytes = m * n * 8
        pin = [L.svdgpu_host_alloc(nbytes) for _ in range(3)]
        pin_sig = L.svdgpu_host_alloc(mn * 8)
        hA = np.ctypeslib.as_array(ctypes.cast(pin[0], ctypes.POINTER(ctypes.c_double)), shape=(n * m,))
        src = np.ascontiguousarray(A_host.T).reshape(-1)
        dp = ctypes.POINTER(ctypes.c_double)

        def step_e2e():
            hA[:] = src                                   # refill the (destroyed) host input; not GPU work
            t0 = time.perf_counter()
            L.svd_gpu(m, n, ctypes.cast(pin[0], dp), ctypes.cast(pin_sig, dp), ctypes.cast(pin[1], dp),
                      ctypes.cast(pin[2], dp))
            return time.perf_counter() - t0
        for _ in range(max(1, args.warmup // 2)):
            step_e2e()
        ts = [step_e2e() for _ in range(args.steps)]
        e2e = {"value": float(np.mean(ts)), "unit": UNIT, "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes + mn * 8 + m * mn * 8 + n * mn * 8,
               "phase_ms_last": dict(zip(["h2d", "bidiag", "ddc", "twisted", "backtransform", "d2h_tail", "total"],
                                         [round(x, 3) for x in D.last_phase_ms()]))}
        for q in pin + [pin_sig]:
            L.svdgpu_host_free(q)
    else:
        # N > 1: rank 0 uploads, every rank downloads its own contiguous column block in parallel
        hU = torch.empty((blk, m), dtype=torch.float64).pin_memory()
        hV = torch.empty((blk, n), dtype=torch.float64).pin_memory()
        hA_t = torch.from_numpy(np.ascontiguousarray(A_host.T)).pin_memory() if rank == 0 else None

        def step_e2e():
            if rank == 0:
                A_master.copy_(hA_t, non_blocking=True)
            step_device()
            hU.copy_(Ublk, non_blocking=True)
            hV.copy_(Vblk, non_blocking=True)
            torch.cuda.synchronize()
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": float(tt.item()) / args.steps, "unit": UNIT, "h2d_bytes_per_step": m * n * 8,
               "d2h_bytes_per_step": (m + n) * mn * 8}

    check = None
    if args.check and rank == 0 and n <= 8192:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import util
        step_device() if world == 1 else None
        torch.cuda.synchronize()
        if world == 1:
            sg, Uh, Vh = sigma.cpu().numpy(), U.cpu().numpy().T, V.cpu().numpy().T
        else:
            sg, Uh, Vh = sig_full[:mn].cpu().numpy(), Ufull[:mn].cpu().numpy().T, Vfull[:mn].cpu().numpy().T
        check = util.svd_metrics(A_host, sg, Uh, Vh)
        check["eps_n"] = float(np.finfo(np.float64).eps * n)
    if args.check and world > 1:
        if rank != 0:
            pass
        dist.barrier()
    if rank == 0:
        line = {"metric": METRIC, "value": ms_per_step * 1e-3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"square {n}x{n} random uniform[1,4) FP64, full U/Sigma/V "
                                       f"(BASELINE.json configs[1] family)", "m": m, "n": n, "panel_nb": nb,
                           "l2": "flushed by a 256 MiB memset between steps (inside the timed region)",
                           "parallelism": "1 GPU" if world == 1 else
                           f"bidiag+dDC on rank 0, vectors/back-transform sharded over {world} ranks"},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        if check:
            line["check_vs_lapack"] = check
        if phase:
            line["phases"] = phase
        if roof:
            line["roofline"] = roof
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # (--size, not --n: torchrun's own parser treats a bare --n as an ambiguous abbreviation)
    ap.add_argument("--size", "--n", dest="n", type=int, default=int(os.environ.get("SVD_BENCH_N", "4096")))
    ap.add_argument("--cpu-n", type=int, default=1024, help="size of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true",
                    help="skip the end-to-end leg (single-shot runs of the largest configs; e2e is then null)")
    ap.add_argument("--check", action="store_true", help="verify the last step's result against LAPACK (n <= 8192)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
