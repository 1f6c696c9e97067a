#!/usr/bin/env python
"""Benchmark of the fused D2Q9 MRT collide-and-stream step (BASELINE.json metric: MLUPS + % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Contract (one JSON line on stdout from rank 0):
  * N = 1  -> workload "cavity4096": BASELINE config 3, 4096 x 4096 cavity, Re 5000, uLB 0.08, fp64, MRT -- the
              "single-GPU roofline run" the metric's roofline fraction is quoted on.  (Config 2, 384^2, is L2-resident
              and launch-bound; it is covered by the parity tests and reported under "extra".)
  * N > 1  -> workload "cavity32768": BASELINE config 5, ONE 32768 x 32768 cavity, Re 10000, fp64, MRT, cut into N
              y-strips with a 3-population halo exchange per interface and step over NCCL (launched by torchrun).
  "step" = one lattice time step of the whole cavity; value = nodes * K / time / 1e6 with the populations resident
  in HBM; inputs are far larger than L2 (2.4 GB of A/B state vs 126 MB), so no explicit flush is needed.
  e2e    = the same metric through the public host API with HOST buffers: upload of the initial populations from
           pinned memory + K steps + download of rho, u (the returned fields), copies inside the timed region.
  roofline = algorithmic bytes (144 B / node fp64, 72 B fp32: 9 loads + 9 stores) / CUDA-event time per launch,
           against the measured copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline = the reference's own Cython/OpenMP step (functions.allfunc, compiled from the reference sources into
           oracle/_ref) on a bounded sample, timed on this box's host cores (N = 1, rank 0 only).
  --impl reference = the reference arm: only the reference's CPU implementation, all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nx, ny, Re, description)
    "cavity4096": (4096, 4096, 5000.0, "lid-driven cavity 4096x4096 Re=5000 uLB=0.08 D2Q9 MRT (BASELINE config 3)"),
    "cavity32768": (32768, 32768, 10000.0, "lid-driven cavity 32768x32768 Re=10000 uLB=0.08 D2Q9 MRT, y-strips (BASELINE config 5)"),
    "cavity384": (384, 384, 3200.0, "lid-driven cavity 384x384 Re=3200 uLB=0.08 D2Q9 MRT (BASELINE config 2)"),
    "datagen256": (384, 384, None, "batched sweep: 256 cavities 384x384, Re = linspace(100, 10000, 256), sharded "
                                   "cavity b -> rank b mod N, no communication (BASELINE config 4)"),
    "refdefault640": (640, 640, 1000.0, "the reference's own published shape and default configuration: 640x640, SRT, "
                                        "turb=1 (Smagorinsky), fp32, 3000 iterations (MRT_GPU.py:48-58, BASELINE.md section 1)"),
}
BYTES_PER_NODE = {"float64": 144, "float32": 72}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for _, l in self.lines]
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in rows:
            parts = [p.strip() for p in l.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own Cython step (oracle/_ref), else the NumPy oracle port
# ----------------------------------------------------------------------------------------------------------------
def cpu_step_timer(variant: str):
    """Returns (kind, cores, fn(n, Re, steps) -> seconds) driving the CPU step exactly like MRT_cython.py:210,232,453."""
    import numpy as np
    from oracle import ref_harness as R
    if R.ref_functions_built(variant):
        F = R.load_ref_functions(variant)
        cores = 4 if variant == "functions" else (os.cpu_count() or 1)

        def run(n, Re, steps):
            vel = np.zeros((2, n, n)); vel[0, :, 0] = 0.08
            fin = F.equ(np.ones((n, n)), vel[0], vel[1])                # MRT_cython.py:210
            F.set_omega(0.08, int(Re), n)                                # :232
            rho = np.sum(fin, axis=0); u = np.zeros((2, n, n)); feq = fin.copy()
            rho, u, fin, feq = F.allfunc(rho, u, fin, feq)               # warm-up call
            t = time.perf_counter()
            for _ in range(steps):
                rho, u, fin, feq = F.allfunc(rho, u, fin, feq)           # :453
            return time.perf_counter() - t
        return "reference", cores, run
    from oracle import lbm_oracle as O       # no compiled reference here: time the oracle's C restatement instead

    def run(n, Re, steps):
        p = O.Params(n, n, Re=Re, collision="MRT")
        O.run_fast(p, 1)
        t = time.perf_counter()
        O.run_fast(p, steps)
        return time.perf_counter() - t
    return "port", os.cpu_count() or 1, run


def cpu_baseline_sample(variant: str, Re: float, budget_s: float, n: int):
    """Time the CPU step on an n x n sample of the workload for about budget_s seconds -> cpu_baseline dict."""
    kind, cores, run = cpu_step_timer(variant)
    t1 = run(n, Re, 1)
    steps = max(2, min(400, int(budget_s / max(t1, 1e-4))))
    t = run(n, Re, steps)
    mlups = n * n * steps / t / 1e6
    what = ("functions.allfunc (functions.pyx:45-222, SRT, fp64) compiled from the reference sources"
            if kind == "reference" else "NumPy oracle port (C-MRT, fp64)")
    threads = ("4 OpenMP threads as shipped (functions.pyx:69)" if variant == "functions" else
               "all %d host threads (num_threads literal removed: modified thread count)" % cores)
    return {"value": round(mlups, 2), "unit": "MLUPS", "cores": cores, "kind": kind,
            "sample": "%d steps of a %dx%d cavity Re=%g, %s, %s, host has %d logical CPUs" % (
                steps, n, n, Re, what, threads if kind == "reference" else "1 thread", os.cpu_count() or 0)}, t, steps


# ----------------------------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    """The reference's own CPU implementation of the step (functions.allfunc of functions.pyx, compiled from the
    reference sources into oracle/_ref) on this box's host cores.  `value` is the module AS SHIPPED: its OpenMP loops
    carry a hard-coded num_threads=4 (functions.pyx:69), so four threads are all it can use; the same source with that
    literal removed (every host thread) is reported beside it as `all_cores_value`.  Under torchrun the launcher puts
    OMP_NUM_THREADS=1 into every rank's environment: the thread count is set explicitly here, before the OpenMP runtime
    is loaded with the module."""
    if rank != 0:
        return
    ncpu = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(ncpu)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    wl = args.workload or ("cavity4096" if args.gpus == 1 else "cavity32768")
    nx, ny, Re, desc = WORKLOADS[wl]
    Re = Re or 5000.0
    kind, cores, run = cpu_step_timer("functions")
    # the whole grid when it fits the host and the time budget (4096^2: ~1.3 s per step), else a bounded sub-cavity
    n = min(nx, 4096)
    run(n, Re, max(args.warmup - 1, 0))                       # warm-up calls (run() itself adds one)
    t = run(n, Re, args.steps)
    mlups = n * n * args.steps / t / 1e6
    extra = {}
    try:
        kind_a, cores_a, run_a = cpu_step_timer("functions_allcores")
        steps_a = max(3, args.steps // 2)
        ta = run_a(n, Re, steps_a)
        extra = {"all_cores_value": round(n * n * steps_a / ta / 1e6, 2), "all_cores": cores_a,
                 "all_cores_note": "same source with the num_threads=4 literal removed (modified thread count), %d steps" % steps_a}
    except Exception as exc:      # pragma: no cover
        extra = {"all_cores_error": str(exc)}
    sample = ("each step = one functions.allfunc call (reference Cython/OpenMP step as shipped: SRT, fp64, num_threads=4 "
              "hard-coded at functions.pyx:69) on %s; the CPU path has no MRT collision and no Smagorinsky closure, "
              "the wall rule differs (SURVEY.md 3.4)" % ("the full %dx%d grid of the workload" % (n, n) if n == nx else
                                                          "a %dx%d sub-cavity of the %dx%d workload" % (n, n, nx, ny)))
    cb = {"value": round(mlups, 2), "unit": "MLUPS", "cores": cores, "kind": kind, "sample": sample}
    cb.update(extra)
    out = {"impl": "reference", "metric": "MLUPS", "value": round(mlups, 2), "unit": "MLUPS", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t / args.steps * 1e3, 4),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": wl, "description": desc, "sample_grid": [n, n], "collision": "SRT (the only one the CPU path has)",
                      "host_logical_cpus": ncpu, "omp_num_threads_env": os.environ["OMP_NUM_THREADS"]},
           "cpu_baseline": cb,
           "e2e": {"value": round(mlups, 2), "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------------------------
def time_device_steps(step_fn, sync_fn, steps, torch):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_fn()
    e0.record()
    step_fn(steps)
    e1.record()
    sync_fn()
    return e0.elapsed_time(e1)       # ms


def run_single_gpu(args):
    import numpy as np
    import torch
    import latticeboltzmannsimulations_b200 as L
    wl = args.workload or "cavity4096"
    nx, ny, Re, desc = WORKLOADS[wl]
    torch.cuda.set_device(0)
    peak, peak_src = measured_peak()
    stream = torch.cuda.current_stream().cuda_stream
    results = {}
    clocks = None
    def timed_run(dtype, sample_clocks):
        nonlocal clocks
        with L.CavitySolver(nx, ny, 1, dtype, "MRT", engine=args.engine) as s:
            s.set_reynolds(Re, 0.08)
            s.init_equilibrium()
            s.step(args.warmup, write_macros=False, stream=stream)
            torch.cuda.synchronize()
            l0 = s.counters()[1]
            sampler = ClockSampler(0)
            if sample_clocks:
                sampler.start()
                time.sleep(0.25)
            t0 = time.time()
            ms = time_device_steps(lambda k: s.step(k, write_macros=False, stream=stream),
                                   torch.cuda.synchronize, args.steps, torch)
            t1 = time.time()
            if sample_clocks:
                clocks = sampler.stop(t0, t1)
            nl = s.counters()[1] - l0
            mlups = nx * ny * args.steps / ms / 1e3
            # roofline per LAUNCH: every step kernel moves 9 loads + 9 stores per node per launch (the fused kernel
            # advances two steps with them), so achieved = bytes/node x nodes / mean launch duration
            gbs = BYTES_PER_NODE[dtype] * nx * ny / (ms / nl) / 1e6
            return {"mlups": mlups, "ms_per_step": ms / args.steps, "gbs": gbs, "engine": s.engine, "launches": nl,
                    "steps_per_launch": args.steps / nl}

    for dtype in ("float64", "float32"):
        results[dtype] = timed_run(dtype, dtype == "float64")
    launches = results["float64"]["launches"]
    # the one-step kernels (temporal blocking off) for reference: these are the HBM-bound ones
    os.environ["LBM_B200_TUNING"] = "two_step=0"
    try:
        one_step = {dt: timed_run(dt, False) for dt in ("float64", "float32")}
    finally:
        del os.environ["LBM_B200_TUNING"]
    # ---- e2e through the host API, fp64: pinned f0 upload + K steps + rho,u download -----------------------------
    huge = nx * ny * 72 > 8e9            # a 77 GB initial state cannot sensibly come from the host: init on device
    rho_out = torch.empty((nx, ny), dtype=torch.float64, pin_memory=True).numpy()
    u_out = torch.empty((2, nx, ny), dtype=torch.float64, pin_memory=True).numpy()
    with L.CavitySolver(nx, ny, 1, "float64", "MRT", engine=args.engine) as s:
        s.set_reynolds(Re, 0.08)
        s.init_equilibrium()
        h2d = 0
        if not huge:
            f0 = torch.empty((9, nx, ny), dtype=torch.float64, pin_memory=True).numpy()
            s.download_f(out=f0)                   # synthetic initial populations, now in pinned host memory
            h2d = f0.nbytes
            for _ in range(2):                     # warm-up of the whole call path (staging buffers, page mapping)
                s.upload_f(f0); s.step(3); s.macros(rho_out=rho_out, u_out=u_out)
        torch.cuda.synchronize()
        t = time.perf_counter()
        if huge:
            s.init_equilibrium()
        else:
            s.upload_f(f0)
        s.step(args.steps, write_macros=True)
        s.macros(rho_out=rho_out, u_out=u_out)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t
        assert np.isfinite(rho_out[::7, ::7]).all() and abs(float(rho_out[::7, ::7].mean()) - 1.0) < 1e-2
    e2e = {"value": round(nx * ny * args.steps / e2e_s / 1e6, 1), "unit": "MLUPS",
           "h2d_bytes_per_step": int(h2d / args.steps), "d2h_bytes_per_step": int((rho_out.nbytes + u_out.nbytes) / args.steps),
           "call": ("CavitySolver.init_equilibrium() + step(K) + macros() -> pinned rho,u (state too large for a host upload)"
                    if huge else "CavitySolver.upload_f(pinned f0) + step(K) + macros() -> pinned rho,u"),
           "seconds": round(e2e_s, 4)}
    # ---- extra: the other configurations, each timed like the headline (device-resident, CUDA events) ----------------
    extra = {}

    def quick(nxx, nyy, dtype, coll, turb, steps, warm, key, tuning=None):
        try:
            with L.CavitySolver(nxx, nyy, 1, dtype, coll, turb, engine=args.engine, tuning=tuning) as s:
                s.set_reynolds(WORKLOADS.get(key, (0, 0, 1000.0))[2] or 1000.0, 0.08)
                s.init_equilibrium(); s.step(warm, write_macros=False, stream=stream)
                ms = time_device_steps(lambda k: s.step(k, write_macros=False, stream=stream), torch.cuda.synchronize, steps, torch)
                return round(nxx * nyy * steps / ms / 1e3, 1), round(ms / steps * 1e3, 3)
        except Exception as exc:      # pragma: no cover
            return "failed: %s" % exc, None

    extra["cavity384_f64_mlups"] = quick(384, 384, "float64", "MRT", False, 1000, 50, "cavity384")[0]
    # the reference's own published shape and default configuration (MRT_GPU.py:48-58: SRT, turb = 1, fp32, 640^2,
    # 3000 iterations; BASELINE.md section 1: Tesla P100 384 true MLUPS, Xeon E5-2680 28 threads 21.9 MLUPS)
    v, us = quick(640, 640, "float32", "SRT", True, 3000, 50, "refdefault640")
    extra["refdefault640"] = {"mlups": v, "us_per_step": us, "config": "640x640 SRT turb=1 fp32, 3000 steps, device-resident",
                              "published_P100_true_mlups": 384.0, "published_xeon28_mlups": 21.9,
                              "vs_published_P100": round(v / 384.0, 1) if isinstance(v, float) else None}
    extra["refdefault640_f64_mlups"] = quick(640, 640, "float64", "SRT", True, 3000, 50, "refdefault640")[0]
    # Smagorinsky closure at the roofline size: 176 / 88 B per node and launch (two extra state values read and written)
    for dt, key in (("float64", "turb4096_f64"), ("float32", "turb4096_f32")):
        extra[key + "_srt_mlups"] = quick(nx, ny, dt, "SRT", True, max(args.steps // 5, 20), 5, wl)[0]
        extra[key + "_srt_one_step_mlups"] = quick(nx, ny, dt, "SRT", True, max(args.steps // 5, 20), 5, wl, {"two_step": 0})[0]
    # same-size single-GPU anchor of the N > 1 strong-scaling runs (2 x 77.3 GB of populations on one B200)
    if wl == "cavity4096":
        try:
            free, _ = torch.cuda.mem_get_info()
            if free > 165e9:
                extra["cavity32768_f64_mlups"] = quick(32768, 32768, "float64", "MRT", False, 10, 3, "cavity32768")[0]
            else:
                extra["cavity32768_f64_mlups"] = "skipped: %.0f GB free" % (free / 1e9)
        except Exception as exc:      # pragma: no cover
            extra["cavity32768_f64_mlups"] = "failed: %s" % exc
    # host link rate (pinned H2D / D2H of 1 GiB), the bound of the e2e number
    try:
        hbuf = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
        dbuf = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
        rates = {}
        for name, (a_, b_) in (("h2d", (dbuf, hbuf)), ("d2h", (hbuf, dbuf))):
            a_.copy_(b_, non_blocking=True); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a_.copy_(b_, non_blocking=True); e1.record(); torch.cuda.synchronize()
            rates[name] = round((1 << 30) / e0.elapsed_time(e1) / 1e6, 1)
        extra["host_link_gbs"] = rates
        e2e["copy_bound_seconds"] = round(h2d / rates["h2d"] / 1e9 + (rho_out.nbytes + u_out.nbytes) / rates["d2h"] / 1e9, 4)
        del hbuf, dbuf
    except Exception as exc:      # pragma: no cover
        extra["host_link_error"] = str(exc)
    # ---- CPU baseline (reference Cython step as shipped, bounded sample) -------------------------------------------
    try:
        cb, _, _ = cpu_baseline_sample("functions", Re, 12.0, 640)
        try:
            cb_all, _, _ = cpu_baseline_sample("functions_allcores", Re, 8.0, 640)
            cb["all_cores_value"] = cb_all["value"]; cb["all_cores"] = cb_all["cores"]
        except Exception:
            pass
    except Exception as exc:      # pragma: no cover
        cb = {"value": None, "unit": "MLUPS", "cores": 0, "kind": "port", "sample": "failed: %s" % exc}
    r64, r32 = results["float64"], results["float32"]
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
            traffic = tj.get("%s_float64" % wl)
            traffic_src = tj.get("source")
    except Exception:
        pass
    out = {"metric": "MLUPS", "value": round(r64["mlups"], 1), "unit": "MLUPS", "n_gpus": 1, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(r64["ms_per_step"], 5), "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": wl, "description": desc, "collision": "MRT", "engine": r64["engine"],
                      "scaling_note": "N > 1 runs ONE 32768^2 cavity in y-strips (strong scaling); its same-size single-GPU "
                                      "anchor is measured in this run: extra.cavity32768_f64_mlups",
                      "l2": "state (2 x %.2f GB) far larger than the 126 MB L2: no flush needed" % (nx * ny * 72 / 1e9)},
           "roofline": {"bound": "hbm", "achieved": round(r64["gbs"], 1), "peak": peak, "unit": "GB/s",
                        "frac": round(r64["gbs"] / peak, 4), "traffic": traffic, "peak_source": peak_src,
                        "traffic_source": traffic_src or "not measured in this run: dram__bytes_read.sum + dram__bytes_write.sum "
                                                         "per launch from the ncu capture summarised under profiles/",
                        "kernel": ("lbm_step_aa (AA pattern, one buffer)" if r64["engine"] == "aa" else
                                   "lbm_step_slide2 (two lattice steps per launch)" if r64["steps_per_launch"] > 1.5 else "lbm_step_ldg"),
                        "algorithmic_bytes_per_node_per_launch": 144, "nodes_per_launch": nx * ny,
                        "steps_per_launch": round(r64["steps_per_launch"], 3),
                        "frac_of_nominal_8TBs": round(r64["gbs"] / 8000.0, 4),
                        "note": "temporal blocking: one launch advances two steps with one read and one write of the "
                                "populations, so MLUPS exceeds peak/144 B while the kernel stays below the HBM peak"},
           "fp32": {"value": round(r32["mlups"], 1), "ms_per_step": round(r32["ms_per_step"], 5),
                    "roofline_achieved": round(r32["gbs"], 1), "roofline_frac": round(r32["gbs"] / peak, 4),
                    "algorithmic_bytes_per_node_per_launch": 72, "steps_per_launch": round(r32["steps_per_launch"], 3)},
           "one_step_kernels": {"f64": {"value": round(one_step["float64"]["mlups"], 1),
                                        "roofline_achieved": round(one_step["float64"]["gbs"], 1),
                                        "roofline_frac": round(one_step["float64"]["gbs"] / peak, 4)},
                                "f32": {"value": round(one_step["float32"]["mlups"], 1),
                                        "roofline_achieved": round(one_step["float32"]["gbs"], 1),
                                        "roofline_frac": round(one_step["float32"]["gbs"] / peak, 4)},
                                "note": "tuning two_step=0: one step per launch, 144 / 72 B per node-step, HBM-bound"},
           "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "extra": extra}
    print(json.dumps(out), flush=True)


def _nccl_options():
    from latticeboltzmannsimulations_b200.distributed import nccl_options
    return nccl_options()


def run_multi_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from latticeboltzmannsimulations_b200.distributed import StripCavity
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=_nccl_options())
    wl = args.workload or "cavity32768"
    nx, ny, Re, desc = WORKLOADS[wl]
    peak, peak_src = measured_peak()
    # correctness of the real NCCL path before anything is timed: a 512 x 384 cavity in strips (one-step kernels at this
    # size, then the sliding two-step kernel forced on the strips) must equal the single-GPU run bit for bit
    import numpy as np
    import latticeboltzmannsimulations_b200 as L
    strips_ok = True
    for tuning in (None, {"slide_min_nodes": 0, "slide_h": 14}):
        chk = StripCavity(512, 384, 1000.0, 0.08, "float64", "MRT", overlap=not args.no_overlap, tuning=tuning)
        chk.step(41, write_macros=True)
        got = chk.gather_fields()
        chk.close()
        if rank == 0:
            want = L.run_cavity(512, 384, 1000.0, steps=41, dtype="float64", return_f=True)
            strips_ok = strips_ok and all(np.array_equal(a_, b_) for a_, b_ in zip(got, want))
    sc = StripCavity(nx, ny, Re, 0.08, "float64", "MRT", engine=args.engine, overlap=not args.no_overlap)
    sc.step(args.warmup)
    sc.sync()
    dist.barrier()
    torch.cuda.synchronize()
    l0 = sc.solver.counters()[1]
    p0 = sc.passes_done
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    time.sleep(0.25)                  # every rank waits alike, then all start together
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sc.fork_from_current_stream()
    sc.step(args.steps)
    sc.join_current_stream()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = sc.solver.counters()[1] - l0
    passes = sc.passes_done - p0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    # e2e: equilibrium start is generated on the device (77 GB of populations cannot sensibly come from the host);
    # the returned rho,u strips are downloaded to pinned host memory inside the timed region.
    rho_out = torch.empty((nx, sc.nyl), dtype=torch.float64, pin_memory=True).numpy()
    u_out = torch.empty((2, nx, sc.nyl), dtype=torch.float64, pin_memory=True).numpy()
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    sc.step(args.steps, write_macros=True)
    sc.sync()
    sc.solver.macros(rho_out=rho_out, u_out=u_out)
    torch.cuda.synchronize()
    e2e_t = torch.tensor([time.perf_counter() - t], device="cuda")
    dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())
    if rank == 0:
        mlups = nx * ny * args.steps / ms / 1e3
        # per pass over memory (a two-step pass advances two steps with 9 loads + 9 stores per node)
        gbs = 144.0 * nx * ny * passes / ms / 1e6
        out = {"metric": "MLUPS", "value": round(mlups, 1), "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 5), "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": wl, "description": desc, "collision": "MRT", "engine": sc.solver.engine,
                          "decomposition": "%d y-strips of %d rows; per interface, direction and double step nine rows of %d "
                                           "values (the 3 crossing populations of two rows + the 3 in-row ones of the edge "
                                           "row, for the two-step kernel) in one NCCL send/recv, %s" % (
                                               world, sc.nyl, nx, "overlapped with the interior update" if sc.overlap else "not overlapped"),
                          "note": "the N=1 line runs cavity4096 (config 3) with the same kernels and reports the same-size "
                                  "single-GPU anchor as extra.cavity32768_f64_mlups",
                          "l2": "per-GPU state far larger than L2: no flush needed"},
               "roofline": {"bound": "hbm", "achieved": round(gbs / world, 1), "peak": peak, "unit": "GB/s",
                            "frac": round(gbs / world / peak, 4), "traffic": None, "peak_source": peak_src,
                            "per_gpu": True, "algorithmic_bytes_per_node_per_pass": 144,
                            "steps_per_pass": round(args.steps / max(passes, 1), 3)},
               "e2e": {"value": round(nx * ny * args.steps / e2e_s / 1e6, 1), "unit": "MLUPS", "h2d_bytes_per_step": 0,
                       "d2h_bytes_per_step": int((rho_out.nbytes + u_out.nbytes) * world / args.steps),
                       "call": "StripCavity.step(K, write_macros) + macros() -> pinned rho,u strips (init generated on device)"},
               "gpu_launches": int(launches), "clocks": clocks, "strips_bitwise_ok": bool(strips_ok),
               "extra": {"halo": "nine rows per neighbour packed into one buffer: one NCCL send + one recv per interface "
                                 "and (double) step" if sc.packed else "row views, nine sends per interface",
                         "rows_per_strip": int(sc.nyl)}}
        print(json.dumps(out), flush=True)
    sc.close()
    dist.barrier()
    dist.destroy_process_group()


def run_datagen(args, rank, world, local_rank):
    """BASELINE config 4: independent cavities sharded over the ranks, no data-path collective."""
    import numpy as np
    import torch
    import latticeboltzmannsimulations_b200 as L
    from latticeboltzmannsimulations_b200.distributed import shard_indices
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=_nccl_options())
    nx, ny, _, desc = WORKLOADS["datagen256"]
    Re_all = np.linspace(100.0, 10000.0, 256)
    mine = shard_indices(256, rank, world)
    peak, peak_src = measured_peak()
    stream = torch.cuda.current_stream().cuda_stream
    res, nl = {}, {}
    for dtype in ("float64", "float32"):
        with L.CavitySolver(nx, ny, len(mine), dtype, "MRT", engine=args.engine) as s:
            s.set_reynolds(Re_all[mine], 0.08)
            s.init_equilibrium()
            s.step(args.warmup, write_macros=False, stream=stream)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            l0 = s.counters()[1]
            ms = time_device_steps(lambda k: s.step(k, write_macros=False, stream=stream), torch.cuda.synchronize, args.steps, torch)
            nl[dtype] = s.counters()[1] - l0
            if world > 1:
                t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
            res[dtype] = 256 * nx * ny * args.steps / ms / 1e3
    # e2e: the public call of the sweep, datagen(Re_list, ...) -> f_final, u_final, feq_initial in host memory.  Its
    # inputs are the Reynolds numbers (the equilibrium start is a function of them: MRT_GPU_datagen.py:259-267), its
    # outputs the population and velocity fields of every cavity, downloaded inside the timed region.
    e2e_s, e2e_err, d2h = float("inf"), None, 0
    try:
        L.datagen(Re_all[mine][:2], nx, ny, steps=3, collision="MRT", dtype="float64")          # warm the call path
        torch.cuda.synchronize()
    except Exception as exc:      # pragma: no cover
        e2e_err = str(exc)
    if world > 1:
        dist.barrier()
    if e2e_err is None:
        try:
            t = time.perf_counter()
            f_fin, u_fin, feq0, _ = L.datagen(Re_all[mine], nx, ny, steps=args.steps, collision="MRT", dtype="float64")
            e2e_s = time.perf_counter() - t
            assert np.isfinite(u_fin[:, :, ::17, ::17]).all()
            d2h = f_fin.nbytes + u_fin.nbytes + feq0.nbytes
            del f_fin, u_fin
        except Exception as exc:      # pragma: no cover
            e2e_err, e2e_s = str(exc), float("inf")
    if world > 1:                 # every rank takes part, whatever happened above: max over ranks, inf = a rank failed
        tt = torch.tensor([e2e_s], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); e2e_s = float(tt.item())
    if e2e_s != float("inf"):
        e2e = {"value": round(256 * nx * ny * args.steps / e2e_s / 1e6, 1), "unit": "MLUPS",
               "h2d_bytes_per_step": int(56 * 256 / args.steps),           # seven doubles of rates per cavity
               "d2h_bytes_per_step": int(d2h * world / args.steps),
               "call": "datagen(Re_list, 384, 384, steps=K) -> f_final, u_final, feq_initial in (pageable) host memory, per rank",
               "seconds": round(e2e_s, 4)}
    else:
        e2e = {"value": None, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "error": e2e_err or "failed on another rank"}
    if rank == 0:
        gbs = res["float64"] * 144 / 1e3 / world * nl["float64"] / args.steps      # per launch (two steps when fused)
        out = {"metric": "MLUPS", "value": round(res["float64"], 1), "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": round(256 * nx * ny / res["float64"] / 1e3, 5), "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": "datagen256", "description": desc, "collision": "MRT", "cavities_per_gpu": len(mine),
                          "l2": "per-GPU working set %.0f MB (A+B) vs 126 MB L2; one cavity (21 MB) is L2-resident, so the "
                                "HBM model can legitimately be exceeded" % (2 * len(mine) * nx * ny * 72 / 1e6)},
               "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4),
                            "traffic": None, "peak_source": peak_src, "per_gpu": True,
                            "algorithmic_bytes_per_node_per_launch": 144, "steps_per_launch": round(args.steps / nl["float64"], 3)},
               "fp32": {"value": round(res["float32"], 1)}, "e2e": e2e, "gpu_launches": int(nl["float64"])}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--engine", default="auto", choices=["auto", "ldg", "tma", "aa"])
    ap.add_argument("--no-overlap", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.steps is None:
        args.steps = 1000 if args.gpus == 1 else 50
    if args.warmup is None:
        args.warmup = 20 if args.gpus == 1 else 5
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.workload == "datagen256":
        if world == 1 and args.gpus > 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        run_datagen(args, rank, world, local_rank)
        return
    if world > 1:
        run_multi_gpu(args, rank, world, local_rank)
    elif args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run on this node
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    else:
        run_single_gpu(args)


if __name__ == "__main__":
    main()
