#!/usr/bin/env python
"""bench.py -- observations/s and ms per Gauss-Newton iteration of the B200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A step is one full Gauss-Newton iteration (the body of the reference's while loop,
main.m:412-494: BuildAwG, normal equations, Schur elimination of the points, (bordered) solve,
update, sumabs) over one synthetic network.  Workload at N=1 is BASELINE.json configs[3], the
configuration the metric is quoted on ("10M-observation synthetic network": 2,000 images,
1M points, ~10M observations, inner constraints, IOP + distortion estimated); it fits one GPU.
At N>1 the same network is sharded by object point (strong scaling) with one all-reduce of the
reduced camera system per iteration.

Output: ONE JSON line on rank 0 (see the keys below).  ``value`` is device-resident throughput
(inputs in HBM, CUDA events); ``e2e`` goes through the host-buffer C-ABI calls per step
(xhat H2D from pinned memory, iterate, xhat + deltasum D2H).  ``cpu_baseline`` / ``--impl
reference`` time the CPU restatement of the reference algorithm (oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (baseline_config index, description)
    "config4": (3, "BASELINE.json configs[3]: synthetic free network 2,000 images / 1M points / ~10M obs"),
    "config3": (2, "BASELINE.json configs[2]: synthetic free network 500 images / 200k points / ~5M obs"),
    "config2": (1, "BASELINE.json configs[1]: EOP-only, 50 images / 20k control points / ~500k obs"),
    "config5block": (4, "one block of BASELINE.json configs[4]: 200 images / 20k points / ~200k obs"),
    "config5": (4, "BASELINE.json configs[4]: BatchRun sweep, 100 independent blocks of 200 images / 20k points / "
                   "~200k obs, advanced concurrently; blocks sharded over ranks, no exchange (replicas only)"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- workload math

def algorithmic_counts(prob):
    """algorithmic bytes / flops of one iteration (per kernel)."""
    s = prob.settings
    n_obs, nImg, nCam, nPts, nTie = prob.n_obs, prob.numImg, prob.numCam, prob.numPts, prob.numtie
    NC = 3 + s.NK + 2
    u_c = prob.u_c
    b = s.u_perimage + s.u_percam + (3 if nTie else 0)
    B_asm = 24 * n_obs + 8 * (6 * nImg + NC * nCam + 3 * nPts) + 8 * (u_c * (u_c + 1) // 2 + u_c) + 8 * 12 * nTie
    F_asm = n_obs * (300 + 4 * (b * (b + 1) // 2 + b))
    m = np.bincount(prob.obs_pt, minlength=nPts)
    m_tie = m[prob.pt_tie >= 0] if nTie else np.zeros(0)
    F_schur = float(np.sum(3.0 * (s.u_perimage * m_tie + s.u_percam) ** 2))
    F_chol = u_c ** 3 / 3.0
    B_rsd = 64 * n_obs
    return dict(B_asm=float(B_asm), F_asm=float(F_asm), F_schur=F_schur, F_chol=float(F_chol), B_rsd=float(B_rsd))


def make_workload(name: str, scale: float):
    import feba_b200 as fb
    idx, desc = WORKLOADS[name]
    if name == "config5":
        name = "config5block"            # the reference arm times one block of the sweep
    # FEBA_BENCH_CACHE=<dir>: keep the generated network between runs on one box (sweeps; generation of
    # configs[3] takes ~45 s of host time and is not part of any timed region)
    cache = os.environ.get("FEBA_BENCH_CACHE")
    path = os.path.join(cache, f"workload_{name}_{scale}.pkl") if cache else None
    if path and os.path.exists(path):
        import pickle
        with open(path, "rb") as fh:
            return pickle.load(fh), desc
    prob = fb.synth.baseline_config(idx, scale=scale)
    if path:
        import pickle
        os.makedirs(cache, exist_ok=True)
        with open(path + f".{os.getpid()}", "wb") as fh:
            pickle.dump(prob, fh, protocol=4)
        os.replace(path + f".{os.getpid()}", path)
    return prob, desc


# --------------------------------------------------------------------------- CPU baseline

def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port(prob):
    """The C/OpenMP + LAPACK restatement of the reference algorithm (oracle/cport.py) on ALL host cores, also
    under a launcher that exported OMP_NUM_THREADS=1 (torchrun does)."""
    from oracle import cport
    threads = cport.set_threads(host_cores())
    return cport.CPort(prob), threads


def cpu_iteration_seconds(prob, fraction: float = 1.0):
    """Time of ONE Gauss-Newton iteration of the CPU restatement of the reference algorithm
    (oracle/cport.py: C + OpenMP block normal equations and Schur complement on all host cores,
    LAPACK dpotrf/dpotrs for the bordered reduced solve, C back-substitution).
    fraction = 1: the full workload.  fraction < 1 (opt-in bounded sample): assembly and back-substitution run
    on the first ``fraction`` of the object points with all their observations and are scaled by the observation
    ratio (their cost is linear in observations); the dense reduced solve is always timed at the FULL reduced
    size u_c.  Returns (seconds, parts)."""
    import copy
    import feba_b200 as fb
    nP = prob.numPts
    k = nP if fraction >= 1.0 else max(int(nP * fraction), min(nP, 1000))
    if k < nP:
        rows = np.nonzero(prob.obs_pt < k)[0]
        sub = copy.copy(prob)
        sub.obs_x, sub.obs_y = prob.obs_x[rows], prob.obs_y[rows]
        sub.obs_img, sub.obs_pt = prob.obs_img[rows], prob.obs_pt[rows]
        sub.xyz0 = prob.xyz0[:k]
        pt_tie = prob.pt_tie[:k].copy()
        sel = pt_tie >= 0
        pt_tie[sel] = np.arange(int(sel.sum()), dtype=np.int32)
        sub.pt_tie = pt_tie
        sub.tie_pt = np.nonzero(sel)[0].astype(np.int32)
        sub.point_ids = None
    else:
        sub = prob
    err, x0, _ = fb.Buildxhat(sub)
    cp, threads = cpu_port(sub)
    tm = {}
    cp.iterate(x0, tm, diag_shift=0.0 if k == nP else 1e-3)
    ratio = prob.n_obs / max(sub.n_obs, 1)
    total = (tm["assemble_s"] + tm["backsub_s"]) * ratio + tm["solve_s"]
    return total, dict(sample_obs=int(sub.n_obs), sample_points=int(k), assemble_schur_s=tm["assemble_s"],
                       solve_s=tm["solve_s"], backsub_s=tm["backsub_s"], scale=ratio, threads=threads)


def cpu_sample_text(prob, parts):
    if parts["sample_points"] == prob.numPts:
        return (f"full workload, one iteration: C/OpenMP assembly+Schur {parts['assemble_schur_s']:.2f} s, LAPACK "
                f"bordered solve at u_c={prob.u_c} {parts['solve_s']:.2f} s, back-substitution {parts['backsub_s']:.2f} s")
    return (f"assembly+Schur+back-substitution on {parts['sample_points']} of {prob.numPts} points "
            f"({parts['sample_obs']} obs, scaled x{parts['scale']:.2f}: {parts['assemble_schur_s']:.2f} s + "
            f"{parts['backsub_s']:.2f} s measured); LAPACK bordered solve at full u_c={prob.u_c} {parts['solve_s']:.2f} s")


def workload_config(args, prob, desc):
    """The keys BOTH arms print under ``config`` (the driver compares them)."""
    return {"workload": args.workload, "desc": desc, "n_obs": int(prob.n_obs), "n_img": int(prob.numImg),
            "n_pts": int(prob.numPts), "u": int(prob.u), "u_c": int(prob.u_c),
            "inner_constraints": int(prob.settings.Inner_Constraints), "type": prob.settings.type}


def run_reference(args, rank):
    """--impl reference: the CPU restatement of the reference algorithm on the host cores (MATLAB /
    Octave do not exist in this image, and literal main.m cannot hold these sizes: SURVEY.md 8d).  Every step is
    one FULL Gauss-Newton iteration of the workload (no sampling unless --cpu-fraction < 1 is given)."""
    if rank != 0:
        return
    prob, desc = make_workload(args.workload, args.scale)
    # every repetition is the same full iteration (10+ s on configs[3]): one untimed repetition warms the pages, and the
    # timed repetitions stop after --cpu-budget seconds (at least two) so that --steps 20 --warmup 5 still ends within
    # a few minutes; the line says how many were measured
    times, parts = [], None
    t, parts = cpu_iteration_seconds(prob, args.cpu_fraction) if args.warmup > 0 else (0.0, None)
    spent = 0.0
    for i in range(max(args.steps, 1)):
        if i >= 2 and spent + (spent / i) > args.cpu_budget:
            break
        t, parts = cpu_iteration_seconds(prob, args.cpu_fraction)
        times.append(t)
        spent += t
    sec = float(np.mean(times))
    val = prob.n_obs / sec
    line = {"impl": "reference", "metric": "observations/sec per Gauss-Newton iteration", "value": val,
            "unit": "obs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, prob, desc),
            "cpu_baseline": {"value": val, "unit": "obs/s", "cores": parts["threads"], "kind": "port",
                             "sample": cpu_sample_text(prob, parts)},
            "e2e": {"value": val, "unit": "obs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "measured_steps": len(times), "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm

def fp64_tensor_peak(torch):
    """cuBLAS DGEMM 8192^3 on this box (FP64 peaks are not in MEASURED_PEAKS.json)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    best = 0.0
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        if i:
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    torch.cuda.empty_cache()
    return best


def ncu_traffic(kernel_substr: str):
    """dram read+write bytes per launch of a kernel from the committed ncu summary of THIS round
    (profiles/r2_ncu_traffic.json, written by scripts/ncu_extract.py), or None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    try:
        d = json.load(open(p))
    except ValueError:
        return None, None
    for k, v in d.get("kernels", {}).items():
        if kernel_substr in k:
            return float(v["dram_bytes_per_launch"]), d.get("source")
    return None, None


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import feba_b200 as fb
    from feba_b200 import shard as sh

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    prob, desc = make_workload(args.workload, args.scale)
    counts = algorithmic_counts(prob)
    err, x0 = fb.Buildxhat(prob)[:2]
    assert err == 0
    # one non-default torch stream carries the library's kernels (graph capture is not allowed on the
    # legacy default stream), its NCCL collectives and the timing events
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    mode, shard, t_create = "single GPU", None, time.perf_counter()
    if world == 1:
        h = fb.Handle(prob, plan=args.plan)
        h.set_stream(stream.cuda_stream)
        adj = None
    else:
        adj = None
        if not args.replicated:
            try:
                adj = sh.GroupAdjustment(prob, plan=args.plan)
                h = adj.h
                mode = "group"
            except fb.FebaError as exc:
                if exc.code != fb.lib.FEBA_ERR_INVALID:
                    raise
                if rank == 0:
                    print(f"[bench] {exc.text}; falling back to the replicated form", file=sys.stderr)
        if adj is None:
            shard = sh.shard_problem(prob, rank, world)
            h = fb.Handle(shard.prob, plan=-1)
            h.set_stream(stream.cuda_stream)
            adj = sh.ShardedAdjustment(h, shard)
            mode = "replicated"
    t_create = time.perf_counter() - t_create
    x_start = x0 if shard is None else fb.Buildxhat(shard.prob)[1]
    h.set_xhat(x_start)
    step_async = h.iterate_async if adj is None else adj.iterate_async
    step_sync = h.iterate if adj is None else adj.iterate

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: inputs already in HBM, nothing read back inside the region
    for _ in range(max(args.warmup, 3)):
        step_async()
    h.sync()
    sampler = ClockSampler(local_rank)
    launches0 = h.launch_count()
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_async()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = h.launch_count() - launches0
    h.sync()
    t = h.last_timing_ex()
    phase_ms = np.array([t["prep_ms"], t["assemble_ms"], t["factor_ms"], t["solve_ms"], t["update_ms"], t["total_ms"],
                         t["exchange_ms"]])
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        pm = torch.tensor(phase_ms, device="cuda")
        dist.all_reduce(pm, op=dist.ReduceOp.MAX)
        phase_ms = pm.cpu().numpy()
    ms_step = float(ms_total.item()) / args.steps
    value = prob.n_obs / (ms_step * 1e-3)

    # ---- end to end through the host-buffer ABI: per step xhat H2D (pinned), iterate, xhat + deltasum D2H
    u_loc = h.u
    pins = [torch.empty(u_loc, dtype=torch.float64).pin_memory() for _ in range(2)]
    bufs = [p.numpy() for p in pins]
    bufs[0][:] = h.get_xhat()
    cur = [0]

    owned = mode == "group"            # a group keeps xhat split over the ranks: each rank moves its own part
    bufs[1][:] = bufs[0]
    n_moved = prob.u_c + 3 * h.num_owned_ties() if owned else u_loc

    def e2e_step():
        # the caller's xhat goes in from pinned host memory, the updated xhat comes back into the
        # other pinned buffer (main.m:484 keeps xhat on the host between iterations)
        if owned:
            h.set_xhat_owned(bufs[cur[0]])
            ds = step_sync()
            h.get_xhat_owned(bufs[1 - cur[0]])
        else:
            h.set_xhat(bufs[cur[0]])
            ds = step_sync()
            h.get_xhat(bufs[1 - cur[0]])
        cur[0] = 1 - cur[0]
        return ds

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_ms = float(t_e2e.item()) * 1e3 / args.steps
    e2e_val = prob.n_obs / (e2e_ms * 1e-3)

    # ---- the whole adjustment from the initial values (main.m:407-494) + residual stage (main.m:569-602): the
    # numbers the parity object compares, and the residual-stage timing.  v / RSD go into page-locked arrays of the
    # caller (one DMA each); the same call into pageable arrays is timed next to it.
    n_rows = h.n_obs
    v_pin = torch.empty(2 * n_rows, dtype=torch.float64).pin_memory()
    r_pin = torch.empty((n_rows, 5), dtype=torch.float64).pin_memory()
    h.set_xhat(x_start)
    if adj is None or mode == "group":
        its, trace = h.solve()
    else:
        its, trace, ds = 0, [], 100.0
        while ds > prob.settings.threshold and its < prob.settings.Iteration_Cap:
            ds = adj.iterate()
            trace.append(ds)
            its += 1
    barrier()
    t0 = time.perf_counter()
    res = h.residuals(v_out=v_pin.numpy(), rsd_out=r_pin.numpy())
    rsd_ms = (time.perf_counter() - t0) * 1e3
    rsd_kernel_ms = h.last_timing_ex()["residual_kernel_ms"]
    rsd_pageable_ms = None
    if world == 1:
        t0 = time.perf_counter()
        h.residuals()
        rsd_pageable_ms = (time.perf_counter() - t0) * 1e3
    xhat_end = h.get_xhat()
    st = prob.settings
    if mode == "replicated":
        # variance factor over ALL ranks (main.m:601): sum the squared residuals of the shards
        ss = torch.tensor([res["sxx"], res["syy"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(ss, op=dist.ReduceOp.SUM)
        sigma02 = float((ss[0] / st.sigma_x ** 2 + ss[1] / st.sigma_y ** 2).item()) / (2 * prob.n_obs - prob.u)
    else:
        sigma02 = float(res["sigma02"])

    # ---- group vs one GPU (driver-visible multi-GPU parity): rank 0 repeats the FIRST step alone on the complete
    # problem and compares the increment of the group with it
    group_check = None
    if world > 1 and mode == "group" and not args.no_group_check:
        h.set_xhat(x_start)
        h.iterate()
        d_group = h.get_delta()
        if rank == 0:
            with fb.Handle(prob, plan=args.plan) as h1:
                h1.set_xhat(x_start)
                h1.iterate()
                d_one = h1.get_delta()
            from oracle.compare import group_rel
            group_check = {"first_step_delta_rel_vs_one_gpu": float(group_rel(prob, d_group, d_one)),
                           "tolerance": 1e-9}
        barrier()

    sp_info = h.sparse_info()
    plan_info = h.plan_info()
    h.close()                                   # every rank: the handle may hold a communicator
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, peak_src = measured_peaks()
    dgemm_peak = fp64_tensor_peak(torch)
    fact_flop = plan_info["flop"] if plan_info["nested_dissection"] else counts["F_chol"]
    kernels = {
        "assemble_schur": {"ms": float(phase_ms[1]), "GBps": counts["B_asm"] / world / (phase_ms[1] * 1e-3) / 1e9,
                           "TFLOPs": (counts["F_asm"] + counts["F_schur"]) / world / (phase_ms[1] * 1e-3) / 1e12},
        "cholesky": {"ms": float(phase_ms[2]), "TFLOPs": fact_flop / (phase_ms[2] * 1e-3) / 1e12,
                     "flop": fact_flop, "flop_dense": counts["F_chol"], "chain_blocks": plan_info["chain_blocks"],
                     "exchange_ms": float(phase_ms[6])},
        "prep_clear": {"ms": float(phase_ms[0])}, "triangular_solve": {"ms": float(phase_ms[3])},
        "update_backsub": {"ms": float(phase_ms[4])},
        "residuals": {"ms": float(rsd_kernel_ms), "GBps": counts["B_rsd"] / world / (max(rsd_kernel_ms, 1e-6) * 1e-3) / 1e9,
                      "frac_of_hbm_peak": counts["B_rsd"] / world / (max(rsd_kernel_ms, 1e-6) * 1e-3) / 1e9 / hbm_peak},
    }
    if phase_ms[2] >= phase_ms[1]:
        tr, tr_src = ncu_traffic("k_gemm_nt")
        roof = {"kernel": "k_gemm_nt (FP64 DMMA products of the blocked Cholesky; achieved = flop EXECUTED by the plan's "
                          "task graph over the WHOLE factorisation phase incl. its latency-bound chain of diagonal blocks)",
                "bound": "tensor", "achieved": kernels["cholesky"]["TFLOPs"], "peak": dgemm_peak, "unit": "TFLOP/s",
                "frac": kernels["cholesky"]["TFLOPs"] / dgemm_peak if dgemm_peak else None,
                "traffic": tr, "traffic_source": tr_src,
                "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (FP64 is not in MEASURED_PEAKS.json); "
                               "DMMA issue peak by microbenchmark 37.1 TFLOP/s (profiles/r1_dmma_rate_microbench.txt)"}
    else:
        tr, tr_src = ncu_traffic("assembly (")
        roof = {"kernel": "k_point_pass + k_image_pass + k_pair_pass (fused BuildAwG + normal blocks + Schur); achieved = "
                          "algorithmic bytes of SURVEY 8(d) over the assembly phase",
                "bound": "hbm", "achieved": kernels["assemble_schur"]["GBps"], "peak": hbm_peak, "unit": "GB/s",
                "frac": kernels["assemble_schur"]["GBps"] / hbm_peak, "traffic": tr, "traffic_source": tr_src,
                "peak_source": peak_src}
    # CPU baseline + parity at the benchmark's own size (rank 0, N=1 only): the CPU port runs the WHOLE adjustment
    # from the same initial values; its mean iteration time is the baseline, its results are the oracle
    cpu, parity = None, None
    if world == 1 and not args.no_cpu:
        from oracle.compare import group_rel
        cp, threads = cpu_port(prob)
        xh, ds_c, n_it, tms, state = np.array(x0, dtype=np.float64), 100.0, 0, [], None
        trace_c = []
        while ds_c > st.threshold and n_it < st.Iteration_Cap and n_it < args.cpu_iterations:
            tm = {}
            xh, ds_c, state = cp.iterate(xh, tm)
            tms.append(tm)
            trace_c.append(ds_c)
            n_it += 1
        sec = float(np.mean([m["total_s"] for m in tms]))
        cpu = {"value": prob.n_obs / sec, "unit": "obs/s", "cores": threads, "kind": "port", "ms_per_step": sec * 1e3,
               "sample": (f"full workload, mean of {n_it} measured iterations of the whole adjustment: assembly+Schur "
                          f"{np.mean([m['assemble_s'] for m in tms]):.2f} s, LAPACK bordered solve at u_c={prob.u_c} "
                          f"{np.mean([m['solve_s'] for m in tms]):.2f} s, back-substitution "
                          f"{np.mean([m['backsub_s'] for m in tms]):.2f} s")}
        converged = ds_c <= st.threshold
        parity = {"against": "oracle/cport.py (C/OpenMP + LAPACK restatement), whole adjustment from the same xhat0",
                  "iterations": [int(its), int(n_it)], "cpu_converged": bool(converged),
                  "deltasum_first_rel": float(abs(trace[0] - trace_c[0]) / trace_c[0])}
        if converged and n_it == its:
            ref = cp.residuals(state, xh)
            vmax = float(np.max(np.abs(ref["v"])))
            parity.update({"xhat_group_rel": float(group_rel(prob, xhat_end, xh)), "xhat_tol": 1e-9,
                           "v_max_abs_over_max_v": float(np.max(np.abs(v_pin.numpy() - ref["v"])) / vmax), "v_tol": 1e-8,
                           "sigma02_rel": float(abs(sigma02 - ref["sigma02"]) / ref["sigma02"]), "sigma02_tol": 1e-8,
                           "sigma02_cpu": float(ref["sigma02"])})
            parity["ok"] = bool(parity["xhat_group_rel"] < 1e-9 and parity["v_max_abs_over_max_v"] < 1e-8
                                and parity["sigma02_rel"] < 1e-8)
        else:
            parity["ok"] = False
    cfg = workload_config(args, prob, desc)
    cfg.update({
        "parallelism": ("single GPU" if world == 1 else
                        (f"nested-dissection group x{world}: one subtree of the image block per rank, local elimination, "
                         f"shared top part ({plan_info['rows'] - plan_info['top_row0']} of {plan_info['rows']} rows) summed "
                         "over NVLink inside feba_iterate (3 small + 1 large ncclAllReduce)" if mode == "group" else
                         f"point-sharded assembly x{world}, all-reduce of the whole reduced system, "
                         + ("column-cyclic shared factorisation (panel broadcasts)"
                            if (adj.shared_factorisation and prob.u_c > 95 * 64) else "replicated factorisation"))),
        "reduced_system": (f"nested dissection: {plan_info['nodes']} tree nodes, {plan_info['supertiles']} supertiles, "
                           f"{sp_info['nonzero_supertiles']} of {sp_info['lower_supertiles']} lower supertiles factorised, "
                           f"chain {plan_info['chain_blocks']} of {plan_info['rows'] // 64} blocks, "
                           f"{sp_info['datum_images']} datum images" if sp_info["active"] else "dense (Buildxhat order)"),
        "l2": "inputs larger than L2 (observations + records + reduced system > 126 MB); no flush needed"})
    line = {
        "metric": "observations/sec per Gauss-Newton iteration", "value": value, "unit": "obs/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e_val, "unit": "obs/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(8 * n_moved),
                "d2h_bytes_per_step": int(8 * n_moved + 16),
                "note": ("per rank: the EOP/IOP part + the tie points the rank owns (feba_set_xhat_owned / "
                         "feba_get_xhat_owned); the complete vector is gathered once after the loop") if owned else
                        "the complete xhat each way (feba_set_xhat / feba_get_xhat)"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": kernels,
        "fp64_dgemm_peak_tflops": dgemm_peak,
        "residual_stage_ms": rsd_ms, "residual_stage": {
            "e2e_ms_pinned_outputs": rsd_ms, "e2e_ms_pageable_outputs": rsd_pageable_ms, "kernel_ms": float(rsd_kernel_ms),
            "d2h_bytes": int(56 * n_rows)},
        "adjustment": {"iterations": int(its), "deltasum": [float(x) for x in trace], "sigma02": sigma02,
                       "xhat_l2": float(np.linalg.norm(xhat_end)), "xhat_sum": float(np.sum(xhat_end)),
                       "xhat_cam_l2": float(np.linalg.norm(xhat_end[:prob.u_c]))},
        "sigma02": sigma02, "create_s": t_create,
        "parity": parity, "group_check": group_check,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_batch(args, rank, world, local_rank):
    """--workload config5: the BatchRun.m sweep.  A step = one Gauss-Newton iteration of EVERY block;
    blocks b = rank, rank + world, ... live on this rank (no collective on the data path)."""
    import torch
    import torch.distributed as dist
    import feba_b200 as fb

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_blocks = max(world, int(round(100 * args.scale))) if args.scale < 1.0 else 100
    mine = list(range(rank, n_blocks, world))
    probs = [fb.synth.baseline_config(4, block=b) for b in mine]
    x0s = [fb.Buildxhat(p)[1] for p in probs]
    handles = [fb.Handle(p) for p in probs]
    for h, x0 in zip(handles, x0s):
        h.set_xhat(x0)
    total_obs = torch.tensor([float(sum(p.n_obs for p in probs))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_obs)
    n_obs_all = float(total_obs.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch = fb.lib.Batch(handles)          # one CUDA graph per step of ALL blocks of this rank (feba_batch)

    def step():
        batch.iterate_async()

    for _ in range(max(args.warmup, 3)):
        step()
    for h in handles:
        h.sync()
    launches0 = sum(h.launch_count() for h in handles)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = sum(h.launch_count() for h in handles) - launches0
    for h in handles:
        h.sync()
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    ms_step = float(dt.item()) * 1e3 / args.steps
    # the round-1 form for comparison: one feba_iterate_async per block (two graph launches each)
    for h in handles:
        h.iterate_async()
    for h in handles:
        h.sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for h in handles:
            h.iterate_async()
    torch.cuda.synchronize()
    per_handle_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    for h in handles:
        h.sync()
    # one block at a time (how BatchRun.m:57-65 runs them), same handles
    t0 = time.perf_counter()
    for h in handles:
        for _ in range(args.steps):
            h.iterate_async()
        h.sync()
    seq_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    # end to end through the host-buffer calls
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for h, x0 in zip(handles, x0s):
            h.set_xhat(x0)
        batch.iterate_async()
        for h in handles:
            h.sync()
            h.get_xhat()
    e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e.item()) * 1e3 / args.steps
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            k = min(5, len(probs))
            t = [cpu_iteration_seconds(p, 1.0) for p in probs[:k]]
            sec = float(np.mean([a for a, _ in t])) * n_blocks
            cpu = {"value": n_obs_all / sec, "unit": "obs/s", "cores": t[0][1]["threads"], "kind": "port",
                   "ms_per_step": sec * 1e3,
                   "sample": f"CPU port on {k} of the {n_blocks} blocks, one after the other (BatchRun.m:57-65), scaled x{n_blocks / k:.0f}"}
        line = {"metric": "observations/sec per Gauss-Newton iteration", "value": n_obs_all / (ms_step * 1e-3),
                "unit": "obs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "config5", "desc": WORKLOADS["config5"][1], "blocks": n_blocks,
                           "blocks_per_rank": len(mine), "n_obs": int(n_obs_all), "u_c_per_block": probs[0].u_c,
                           "timing": "host clock around a device-synchronised region; one CUDA graph launch per step holds "
                                     "the step of every block of the rank (feba_batch)",
                           "l2": "per-rank working set (observations + records + reduced systems) > 126 MB"},
                "e2e": {"value": n_obs_all / (e2e_ms * 1e-3), "unit": "obs/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(8 * sum(h.u for h in handles)),
                        "d2h_bytes_per_step": int(8 * sum(h.u + 1 for h in handles))},
                "gpu_launches": int(launches), "clocks": clocks,
                "one_block_at_a_time_ms_per_step": seq_ms, "one_launch_pair_per_block_ms_per_step": per_handle_ms,
                "cpu_baseline": cpu,
                "roofline": None}
        print(json.dumps(line), flush=True)
    batch.close()
    for h in handles:
        h.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config4", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (tests only)")
    ap.add_argument("--cpu-fraction", type=float, default=1.0,
                    help="--impl reference: fraction of the points the linear stages are timed on per step (1 = the "
                         "full workload, measured)")
    ap.add_argument("--cpu-budget", type=float, default=150.0,
                    help="--impl reference: stop timing further full iterations once this many seconds are spent (>= 2 are timed)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-iterations", type=int, default=12,
                    help="cap on the iterations of the CPU port's whole adjustment (cpu_baseline + parity at N=1)")
    ap.add_argument("--plan", type=int, default=0, help="feba_settings.plan: 0 automatic, -1 dense order, 1 nested dissection")
    ap.add_argument("--replicated", action="store_true",
                    help="N > 1: the replicated form (all-reduce of the whole reduced system) instead of the group form")
    ap.add_argument("--no-group-check", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: libraries that print there on their own (NCCL's version
    # banner at communicator creation) are sent to stderr; print() keeps the real stdout
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload == "config5":
        run_batch(args, rank, world, local_rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
