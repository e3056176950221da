#!/usr/bin/env python
"""bench.py — the GP likelihood hot path on B200, one JSON line per run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload NAME] [--patterns P]

One *step* = what BASELINE.json's north_star targets: one full-DAG likelihood pass
(GPDAG::PopulatePLVs + GPDAG::ComputeLikelihoods, gp_instance.cpp:231-235) plus ONE branch-length
optimisation sweep (every edge once, Brent, all edges of the DAG batched into one dependency level)
through GPEngine::ProcessOperations, over one synthetic alignment, from branch lengths 0.1.
metric = GP pattern x edge PLV updates / s = 2 (E - R) P_total / step time (the unit of
work is one IncrementWithWeightedEvolvedPLV on one pattern; the sweep adds time, not units). The
pass-only number (round 1's `value`) is kept under `pass_only`.

Every N runs BASELINE.json configs[4] (1000 taxa, DAG from 5000 trees) weak-scaled at 125k site
patterns per GPU (N = 8 is the full 1M-pattern alignment; one shard fills 127 GB of HBM), patterns
sharded, per-edge scalars all-reduced inside the engine. At N = 1 the line also carries configs[3]
(200 taxa x 100k patterns) under `config3_single_b200` and configs[0..2] (DS1 pass, DS1
EstimateBranchLengths x 10 + marginal, fluA pass + SBN update) under `small_configs`, each beside the
reference CPU engine timed in the same run.

  value   : device-resident throughput (alignment already in HBM), CUDA events, max over ranks
  e2e     : same step through the C-ABI with HOST buffers: every step uploads the alignment, weights
            and branch lengths from pinned memory and reads back per-edge log-likelihoods, the
            marginal and the optimised branch lengths
  roofline: dominant kernel, algorithmic bytes / event-timed kernel time
  parity  : N = 1: the CUDA engine against the reference CPU engine on the pattern subsample the
            cpu_baseline leg times (1e-9 / counts equal / 1e-6); N > 1: shard additivity and
            rank agreement. A failed check makes the run exit non-zero.
  cpu_baseline: the UNMODIFIED reference CPU GPEngine (oracle/_ref; falls back to the plain-C port)
            on a bounded pattern sample of the same workload, 1 core
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

METRIC = "gp_pattern_edge_plv_updates_per_s"
UNIT = "updates/s"
# Every N runs the SAME per-GPU work (weak scaling): BASELINE.json configs[4], 1000 taxa, DAG from 5000
# trees, 125k site patterns per GPU (N = 8 is the full 1M-pattern alignment; one shard is 127 GB of HBM).
BENCH_WORKLOAD = "synthetic-1000taxa-1Mpat-5000trees"
# configs[3] (200 taxa x 100k patterns, DAG from 1000 trees) is reported beside it at N = 1.
SINGLE_B200_WORKLOAD = "synthetic-200taxa-100kpat-1000trees"
STEP = ("PopulatePLVs + ComputeLikelihoods (full-DAG likelihood pass) + one batched Brent branch-length sweep "
        "(every edge once, from t = 0.1)")
LL_RTOL, BL_ATOL = 1e-9, 1e-6  # BASELINE.json north_star tolerances


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--patterns", type=int, default=None, help="patterns per GPU (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the Gauss-Seidel (reference schedule) sweep")
    ap.add_argument("--no-config3", action="store_true", help="N = 1: skip the extra configs[3] measurement")
    ap.add_argument("--no-small", action="store_true", help="N = 1: skip configs[0..2]")
    ap.add_argument("--ref-procs", type=int, default=0, help="--impl reference: processes (default: all cores, <= 32)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def rel_err(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    both_ninf = np.isneginf(got) & np.isneginf(want)
    with np.errstate(invalid="ignore"):
        err = np.where(both_ninf, 0.0, np.abs(got - want) / np.maximum(np.abs(want), 1.0))
    return float(np.max(err)) if err.size else 0.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- the reference CPU engine (oracle/_ref, else the plain-C port): timing AND the parity checker --------
def make_cpu_engine(symbols, weights, site_count, node_count, edge_count, q, unconditional, inverted, thr=1e-40):
    from oracle import ref_engine
    try:
        if not ref_engine.available():
            raise RuntimeError("oracle/_ref not built")
        return "reference", ref_engine.RefEngine.from_arrays(symbols, weights, site_count, node_count, edge_count, q,
                                                             unconditional, inverted, rescaling_threshold=thr)
    except Exception:  # the prebuilt reference is missing: fall back to the plain-C port
        from oracle.port_engine import PortEngine
        return "port", PortEngine(symbols, weights, site_count, node_count, edge_count, q, unconditional, inverted, thr)


def cpu_step(workload, sample_patterns, repeats):
    """Times the bench step (pass, then one batched Brent sweep from t = 0.1) on the reference CPU GPEngine
    (1 thread: the reference engine is single-threaded by construction) over the first `sample_patterns`
    site patterns. Returns the engine kind, per-repeat (pass_s, sweep_s) and the outputs the parity check
    compares (taken from the last repeat)."""
    sub = workload.subsample(sample_patterns)
    pop, lik = workload.ops("populate_plvs"), workload.ops("compute_likelihoods")
    blo = workload.ops("batched_branch_length_optimization")
    kind, eng = make_cpu_engine(sub.symbols, sub.weights, sub.site_count, workload.dag.node_count,
                                workload.dag.edge_count, workload.sbn_prior, workload.unconditional, workload.inverted)
    eng.process_operations(*pop)  # warm-up (page-faults the reference's mmap'd PLV file in)
    eng.process_operations(*lik)
    times, out = [], {}
    for _ in range(repeats):
        eng.set_branch_lengths_to_constant(0.1)
        eng.reset_optimization_count()
        t0 = time.perf_counter()
        eng.process_operations(*pop)
        eng.process_operations(*lik)
        t1 = time.perf_counter()
        out = {"per_gpcsp_ll": eng.per_gpcsp_log_likelihoods().copy(), "log_marginal": eng.log_marginal_likelihood(),
               "counts": eng.rescaling_counts().copy()}
        t2 = time.perf_counter()
        eng.process_operations(*blo)
        t3 = time.perf_counter()
        times.append((t1 - t0, t3 - t2))
    out["branch_lengths"] = eng.branch_lengths().copy()
    out["engine"] = eng  # kept open: the parity check asks it for objective values of disputed edges
    return kind, np.asarray(times), out, sub


def cpu_sample_size(workload):
    # ~1.5e8 pattern x edge updates per timed pass (a few seconds at the reference's ~3e7/s; the sweep is
    # ~4x a pass), and a reference mmap file (6 * (2 (N + 16) + 16) PLVs of 32 P bytes) under ~6 GB.
    per_pattern = workload.updates_per_pass()
    by_time = int(1.0e8 // per_pattern)
    by_memory = int(6e9 // (6 * (2 * (workload.dag.node_count + 16) + 16) * 32))
    return max(64, min(workload.pattern_count, by_time, by_memory))


def _reference_worker(job):
    """One process of the reference arm: the unmodified single-threaded reference GPEngine on its own
    slice of the workload's site patterns (the path is pattern-parallel)."""
    name, index, patterns, warmup, steps, gate = job
    from bito_b200.synthetic import make_named_workload
    wl = make_named_workload(name, rank=1000 + index, pattern_count=patterns)
    if gate is not None:
        gate.wait()  # all workers enter the timed steps together
    kind, times, out, _ = cpu_step(wl, wl.pattern_count, warmup + steps)
    out["engine"].close()
    return dict(kind=kind, times=times.tolist(), marginal=out["log_marginal"], updates=wl.updates_per_pass(),
                summary=wl.dag.summary())


def run_reference(args, rank, world):
    """bench.py --impl reference: the reference's own CPU implementation of the step on the box's host
    cores. The reference GPEngine is single-threaded, so "all the host threads it can use" is one
    process per core, each running the unmodified engine over its own slice of the site patterns."""
    if rank != 0:
        return
    import multiprocessing as mp
    name = args.workload or BENCH_WORKLOAD
    procs = args.ref_procs or min(32, os.cpu_count() or 1)
    # per process: a few seconds of CPU work per step and ~1.3 GB of touched PLV pages (the reference's PLVs
    # are an mmap'd file); all processes together stay under ~21 GB
    patterns = args.patterns or max(256, min(768 if "1000taxa" in name else 1536, 24576 // procs))
    if procs == 1:
        results = [_reference_worker((name, 0, patterns, args.warmup, args.steps, None))]
    else:
        ctx = mp.get_context("spawn")
        with ctx.Manager() as manager, ctx.Pool(procs) as pool:
            gate = manager.Barrier(procs)
            results = pool.map(_reference_worker,
                               [(name, i, patterns, args.warmup, args.steps, gate) for i in range(procs)], chunksize=1)
    # every process repeats `steps` timed steps; aggregate throughput = all updates / slowest process
    step_s = [float(np.sum(np.asarray(r["times"])[args.warmup:])) for r in results]
    pass_s = [float(np.sum(np.asarray(r["times"])[args.warmup:, 0])) for r in results]
    sec = max(step_s) / args.steps
    updates = results[0]["updates"] * patterns * procs
    value = updates / sec
    kind = "reference" if all(r["kind"] == "reference" for r in results) else "port"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, **results[0]["summary"], "patterns_per_step": patterns * procs, "step": STEP,
                   "note": "per-pattern throughput on a bounded sample: same DAG, same op lists, the path is "
                           "linear in the pattern count (SURVEY.md 8d allows timing the reference on a subsample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": f"{procs} processes x {patterns} site patterns of the workload per step (same "
                                   "DAG and op lists); the reference GPEngine is single-threaded, so each core "
                                   "runs its own unmodified engine on a disjoint pattern slice; slowest process "
                                   f"timed; 1-process rate {results[0]['updates'] * patterns / (step_s[0] / args.steps):.3e}"},
        "pass_only": {"value": updates / (max(pass_s) / args.steps), "unit": UNIT,
                      "ms_per_step": max(pass_s) / args.steps * 1e3},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "log_marginal": results[0]["marginal"],
    }
    print(json.dumps(line))


def check_jacobi_sweep(got, want, blo, cpu):
    """Branch lengths after one batched sweep, CUDA vs reference. Brent's first parabolic step sits on an
    acceptance boundary for ~0.1 % of the edges, where 1e-11 relative noise in the objective flips the
    decision (two builds of the UNMODIFIED reference disagree there too, oracle/ref_jacobi_sweep_sensitivity.py):
    every edge outside 1e-6 must lie inside Brent's own tolerance AND have the same objective value at both
    lengths to 1e-9 (a flat objective). Returns the report and whether it passes."""
    off = np.nonzero(np.abs(got - want) > BL_ATOL)[0]
    tol = 2.0 ** -9
    ok = off.size <= 0.005 * want.size
    inside = bool(np.all(np.abs(np.log(got[off]) - np.log(want[off])) <= 4 * (tol * np.abs(np.log(want[off])) + tol / 4)))
    by_edge = {int(r[3]): (int(r[1]), int(r[2])) for r in blo[0]}
    flat_worst = 0.0
    for g in off[:64]:
        leafward, rootward = by_edge[int(g)]
        values = []
        for t in (got[g], want[g]):
            bl = want.copy()
            bl[g] = t
            cpu.set_branch_lengths(bl)
            values.append(cpu.log_likelihood_and_derivatives(int(g), rootward, leafward)[0])
        flat_worst = max(flat_worst, abs(values[0] - values[1]) / max(1.0, abs(values[1])))
    ok = ok and inside and flat_worst <= 1e-9
    within = np.abs(got - want) <= BL_ATOL
    return {"edges": int(want.size), "edges_outside_1e-6": int(off.size),
            "max_abs_err_within": float(np.max(np.abs(got - want)[within])) if within.any() else 0.0,
            "outside_all_within_brent_tolerance": inside, "outside_objective_rel_diff_max": flat_worst}, ok


def parity_against_cpu(wl, sub, cpu_out, local_rank):
    """N = 1: the CUDA engine on the pattern subsample the cpu_baseline leg just ran, same op lists. The sweep is
    checked twice: with the engine's own choice for a level of this size (one block per edge at a few thousand
    patterns) and with the streamed Taylor-model scheme the timed full-size run uses (BITO_GP_OPT_SCHEME=0)."""
    from bito_b200.gp_engine import GPEngine
    dag = wl.dag
    pop, lik = wl.ops("populate_plvs"), wl.ops("compute_likelihoods")
    blo = wl.ops("batched_branch_length_optimization")
    out = {}
    ok = True
    for label, scheme_env in (("engine_choice", None), ("streamed_as_timed", "0")):
        if scheme_env is not None:
            os.environ["BITO_GP_OPT_SCHEME"] = scheme_env
        try:
            with GPEngine(sub.symbols, sub.weights, sub.site_count, dag.node_count, dag.edge_count,
                          sbn_prior=wl.sbn_prior, unconditional_node_probabilities=wl.unconditional,
                          inverted_sbn_prior=wl.inverted, device=local_rank) as gpu:
                gpu.process_operations(*pop)
                gpu.process_operations(*lik)
                if label == "engine_choice":
                    ll_err = rel_err(gpu.get_per_gpcsp_log_likelihoods(), cpu_out["per_gpcsp_ll"])
                    marg_err = rel_err(gpu.get_log_marginal_likelihood(), cpu_out["log_marginal"])
                    counts = gpu.get_rescaling_counts()
                    want_counts = cpu_out["counts"]
                    n = min(counts.size, want_counts.size)
                    counts_equal = bool(np.array_equal(counts[:n], want_counts[:n]))
                s0 = gpu.stats()
                gpu.process_operations(*blo)
                sweep, sweep_ok = check_jacobi_sweep(gpu.get_branch_lengths(), cpu_out["branch_lengths"], blo,
                                                     cpu_out["engine"])
                s1 = gpu.stats()
                sweep["optimizer_scheme"] = int(s1["optimizer_scheme"])
                sweep["objective_evaluations"] = int(s1["objective_evaluations"] - s0["objective_evaluations"])
                sweep["objective_passes"] = int(s1["objective_passes"] - s0["objective_passes"])
                out[label] = sweep
                ok = ok and sweep_ok
        finally:
            os.environ.pop("BITO_GP_OPT_SCHEME", None)
    ok = ok and ll_err <= LL_RTOL and marg_err <= LL_RTOL and counts_equal
    return {"ok": bool(ok), "patterns": int(sub.pattern_count), "against": "reference CPU GPEngine, same DAG and op lists",
            "max_rel_err": max(ll_err, marg_err), "per_edge_ll_max_rel_err": ll_err, "log_marginal_rel_err": marg_err,
            "counts_equal": counts_equal, "nonzero_counts": int(np.count_nonzero(want_counts)),
            "sweep_branch_lengths": out["streamed_as_timed"], "sweep_branch_lengths_engine_choice": out["engine_choice"],
            "sweep_optimizer_scheme": out["streamed_as_timed"]["optimizer_scheme"],
            "tolerances": {"log_likelihoods_rel": LL_RTOL, "branch_lengths_abs": BL_ATOL}}


def small_configs(local_rank):
    """BASELINE.json configs[0..2] (the reference's own CPU-runnable cases, fixtures generated from the
    reference under tests/golden/): latency-bound at 934 / 238 patterns. Each is timed on the GPU (CUDA
    events, graph replays) and on the reference CPU engine in this same run, and checked against it."""
    import torch
    from bito_b200.gp_engine import GPEngine
    out = {}
    stream = torch.cuda.Stream()  # a real stream: the handle of torch's default stream is NULL, which bito_gp_set_stream reads as
    # "the engine's own stream", and events recorded on the default stream do not see work on that one
    torch.cuda.set_stream(stream)

    def load(name):
        # NpzFile re-reads (and unzips) an array on EVERY z[...]: read everything once, so that the timed loops
        # below replay op lists that are already in memory (parsing a fixture is not part of the path)
        with np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")) as f:
            z = {k: f[k] for k in f.files}
        return z, (lambda which: (z["ops_" + which], z["vec_" + which]))

    def engines(z):
        args = (z["symbols"], z["weights"], int(z["site_count"]), int(z["node_count"]), int(z["edge_count"]))
        kind, cpu = make_cpu_engine(*args, z["sbn_prior"], z["unconditional_node_probabilities"], z["inverted_sbn_prior"])
        gpu = GPEngine(*args, sbn_prior=z["sbn_prior"], unconditional_node_probabilities=z["unconditional_node_probabilities"],
                       inverted_sbn_prior=z["inverted_sbn_prior"], device=local_rank)
        gpu.set_stream(stream.cuda_stream)
        for e in (cpu, gpu):
            e.set_branch_lengths(z["initial_branch_lengths"])
        return kind, cpu, gpu

    def gpu_ms(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def cpu_ms(fn, reps):
        fn()
        best = float("inf")
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        return best * 1e3

    def shape(z):
        return {"taxa": int(z["symbols"].shape[0]), "patterns": int(z["symbols"].shape[1]),
                "nodes": int(z["node_count"]), "edges": int(z["edge_count"])}

    # configs[0]: DS1 PopulatePLVs + ComputeLikelihoods
    z, ops = load("ds1_config1")
    kind, cpu, gpu = engines(z)

    def pass_on(e):
        return lambda: (e.process_operations(*ops("populate_plvs")), e.process_operations(*ops("compute_likelihoods")))
    g, c = gpu_ms(pass_on(gpu), 50), cpu_ms(pass_on(cpu), 5)
    err = max(rel_err(gpu.get_per_gpcsp_log_likelihoods(), cpu.per_gpcsp_log_likelihoods()),
              rel_err(gpu.get_log_marginal_likelihood(), cpu.log_marginal_likelihood()),
              rel_err(gpu.get_per_gpcsp_log_likelihoods(), z["t0_pass_per_gpcsp_ll"]))
    out["ds1_pass"] = {**shape(z), "what": "configs[0]: DS1 PopulatePLVs + ComputeLikelihoods, JC69", "gpu_ms": g,
                       "cpu_ms": c, "cpu_kind": kind, "max_rel_err": err,
                       "counts_equal": bool(np.array_equal(gpu.get_rescaling_counts()[:cpu.rescaling_counts().size],
                                                           cpu.rescaling_counts())),
                       "ok": bool(err <= LL_RTOL)}

    # configs[1]: EstimateBranchLengths (gp_instance.cpp:241-308): 10 iterations of (Gauss-Seidel Brent sweep,
    # PopulatePLVs, marginal), then ComputeMarginalLikelihood
    def estimate(e):
        def run():
            e.set_branch_lengths(z["initial_branch_lengths"])
            e.set_optimization_method("brent")
            e.reset_optimization_count()
            e.process_operations(*ops("populate_plvs"))
            e.process_operations(*ops("marginal_likelihood"))
            for _ in range(10):
                e.process_operations(*ops("branch_length_optimization"))
                e.process_operations(*ops("populate_plvs"))
                e.process_operations(*ops("marginal_likelihood"))
                e.increment_optimization_count()
        return run
    g, c = gpu_ms(estimate(gpu), 5), cpu_ms(estimate(cpu), 2)
    bl_err = float(np.max(np.abs(gpu.get_branch_lengths() - cpu.branch_lengths())))
    marg_err = rel_err(gpu.get_log_marginal_likelihood(), cpu.log_marginal_likelihood())
    out["ds1_estimate_branch_lengths"] = {
        **shape(z), "what": "configs[1]: DS1 EstimateBranchLengths, 10 Brent iterations (reference Gauss-Seidel "
                            "schedule) + ComputeMarginalLikelihood", "gpu_ms": g, "cpu_ms": c, "cpu_kind": kind,
        "branch_length_max_abs_err": bl_err, "log_marginal_rel_err": marg_err,
        # ten coupled sweeps amplify Brent's 2^-9 stopping tolerance: the marginal is the stable check
        "ok": bool(marg_err <= 1e-7)}
    cpu.close()
    gpu.close()

    # configs[2]: fluA full pass + UpdateSBNProbabilities (GPInstance::EstimateSBNParameters)
    z, ops = load("fluA")
    kind, cpu, gpu = engines(z)

    def sbn_on(e):
        def run():
            e.set_sbn_parameters(z["sbn_prior"])
            e.process_operations(*ops("populate_plvs"))
            e.process_operations(*ops("compute_likelihoods"))
            e.process_operations(*ops("optimize_sbn_parameters"))
        return run
    g, c = gpu_ms(sbn_on(gpu), 50), cpu_ms(sbn_on(cpu), 5)
    q_err = float(np.max(np.abs(gpu.get_sbn_parameters() - cpu.sbn_parameters())))
    err = rel_err(gpu.get_per_gpcsp_log_likelihoods(), cpu.per_gpcsp_log_likelihoods())
    out["fluA_pass_sbn_update"] = {**shape(z), "what": "configs[2]: fluA PopulatePLVs + ComputeLikelihoods + "
                                   "UpdateSBNProbabilities", "gpu_ms": g, "cpu_ms": c, "cpu_kind": kind,
                                   "max_rel_err": err, "sbn_parameters_max_abs_err": q_err,
                                   "ok": bool(err <= LL_RTOL and q_err <= 1e-6)}
    cpu.close()
    gpu.close()
    return out


def measure(args, name, rank, world, local_rank, extras=True):
    """One workload on this rank's GPU: device-resident step, e2e step, per-kernel roofline, sweeps,
    parity, CPU baseline. Returns (JSON line or None, all checks passed)."""
    import torch
    from bito_b200 import _lib
    from bito_b200 import distributed as D
    from bito_b200.gp_engine import GPEngine
    from bito_b200.synthetic import make_named_workload
    t_setup = time.time()
    wl = make_named_workload(name, rank=rank, pattern_count=args.patterns)
    P_local = wl.pattern_count
    P_total = P_local * world
    dag = wl.dag
    pop, lik = wl.ops("populate_plvs"), wl.ops("compute_likelihoods")
    blo = wl.ops("batched_branch_length_optimization")
    n_edges_opt = int((blo[0][:, 0] == 5).sum())  # gp_operation.OPTIMIZE_BRANCH_LENGTH

    # Pinned host staging for the e2e leg (and the initial upload).
    sym_pinned = torch.from_numpy(wl.symbols).pin_memory()
    w_pinned = torch.from_numpy(wl.weights).pin_memory()
    bl_pinned = torch.full((dag.edge_count,), 0.1, dtype=torch.float64).pin_memory()
    flags = _lib.FLAG_NO_LOGLIK_MATRIX if dag.edge_count * P_local * 8 > 16e9 else 0
    engine = GPEngine(sym_pinned.numpy(), w_pinned.numpy(), wl.site_count * world, dag.node_count, dag.edge_count,
                      sbn_prior=wl.sbn_prior, unconditional_node_probabilities=wl.unconditional,
                      inverted_sbn_prior=wl.inverted, device=local_rank, flags=flags)
    stream = torch.cuda.Stream()  # a real stream: the handle of torch's default stream is NULL, which bito_gp_set_stream reads as
    # "the engine's own stream", and events recorded on the default stream do not see work on that one
    torch.cuda.set_stream(stream)
    engine.set_stream(stream.cuda_stream)

    barrier, max_over_ranks = D.barrier, D.max_over_ranks
    checks_ok = True

    def pass_step():
        engine.process_operations(*pop)
        engine.process_operations(*lik)

    def device_step():
        engine.set_branch_lengths_to_constant(0.1)  # device fills: every step does the same work
        engine.reset_optimization_count()
        engine.process_operations(*pop)
        engine.process_operations(*lik)
        engine.process_operations(*blo)

    def upload():
        engine.set_site_patterns(sym_pinned.numpy(), w_pinned.numpy())
        engine.set_branch_lengths(bl_pinned.numpy())

    def e2e_step():
        upload()
        engine.reset_optimization_count()
        engine.process_operations(*pop)
        engine.process_operations(*lik)
        engine.process_operations(*blo)
        return engine.get_per_gpcsp_log_likelihoods(), engine.get_log_marginal_likelihood(), engine.get_branch_lengths()

    def timed(step_fn, steps, warmup):
        for _ in range(warmup):
            step_fn()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = engine.stats()["kernel_launches"]
        ev0.record(stream)
        for _ in range(steps):
            step_fn()
        ev1.record(stream)
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        return ms / steps, engine.stats()["kernel_launches"] - launches0

    # ---- N > 1: the same shard on one GPU with no communicator (weak-scaling reference + additivity) ----
    standalone = None
    parity = None
    if world > 1:
        ms_alone, _ = timed(device_step, max(3, args.steps // 2), 3)
        engine.set_branch_lengths_to_constant(0.1)  # device_step ends with optimised lengths
        pass_step()
        local_ll, local_marg = engine.get_per_gpcsp_log_likelihoods(), engine.get_log_marginal_likelihood()
        standalone = {"value_per_gpu": wl.updates_per_pass() * P_local / (ms_alone * 1e-3), "unit": UNIT,
                      "ms_per_step": ms_alone,
                      "what": "this rank's shard (same DAG, same patterns per GPU, same step) before joining the "
                              "communicator; max over ranks"}
        D.connect_engine(engine)
        # shard additivity: the all-reduced per-edge sums are the sum over ranks of the stand-alone sums, and
        # every rank holds the same rescaling counts
        engine.set_branch_lengths_to_constant(0.1)
        pass_step()
        global_ll, global_marg = engine.get_per_gpcsp_log_likelihoods(), engine.get_log_marginal_likelihood()
        summed = D.sum_over_ranks(np.concatenate([local_ll, [local_marg]]))
        add_err = rel_err(np.concatenate([global_ll, [global_marg]]), summed)
        counts = engine.get_rescaling_counts().astype(np.float64)
        counts_same = bool(np.array_equal(D.max_over_ranks_array(counts), -D.max_over_ranks_array(-counts)))
        parity = {"shard_additivity_max_rel_err": add_err, "counts_identical_on_all_ranks": counts_same,
                  "tolerance": 1e-12}

    # ---- device-resident throughput (`value`): pass + one batched sweep ---------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_per_step, launches = timed(device_step, args.steps, max(3, args.warmup))
    clocks = sampler.stop()
    updates = wl.updates_per_pass() * P_total
    value = updates / (ms_per_step * 1e-3)
    st = engine.stats()
    sweep_scheme = int(st["optimizer_scheme"])
    bl_after_batched = engine.get_branch_lengths()
    pass_ms, pass_launches = timed(pass_step, args.steps, 2)
    log_marginal = engine.get_log_marginal_likelihood()
    pass_levels = 0
    for lst in (pop, lik):
        engine.process_operations(*lst)
        pass_levels += int(engine.stats()["levels_last"])
    st = engine.stats()

    # ---- e2e through the C-ABI with host buffers --------------------------------------------------
    e2e_ms, _ = timed(e2e_step, max(2, args.steps // 2), 1)
    e2e_value = updates / (e2e_ms * 1e-3)
    h2d = wl.symbols.nbytes + wl.weights.nbytes + dag.edge_count * 8
    d2h = 2 * dag.edge_count * 8 + 8
    upload_ms, _ = timed(upload, 5, 1)  # the host->device copies alone, on their own events

    # ---- per-kernel roofline: CUDA events around every launch, graphs bypassed ------------------
    peak_gbs, peak_src = peaks()
    engine.set_profiling(True)
    engine.reset_kernel_profile()
    prof_steps = min(3, args.steps)
    for _ in range(prof_steps):
        device_step()
    prof = engine.kernel_profile()
    engine.set_profiling(False)
    prof.sort(key=lambda k: -k["total_ms"])
    total_prof_ms = sum(k["total_ms"] for k in prof) or 1.0
    top = prof[0]
    per_launch_ms = top["total_ms"] / top["launches"]
    per_launch_bytes = top["algorithmic_bytes"] / top["launches"]
    achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
    traffic = None
    for tp in ("r02_traffic.json", "r01_traffic.json"):
        traffic_path = os.path.join(ROOT, "profiles", tp)
        if traffic is None and os.path.exists(traffic_path) and args.patterns is None:
            # ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel on this workload's
            # per-GPU shard (every rank moves the same bytes)
            with open(traffic_path) as f:
                traffic = json.load(f).get(name, {}).get(top["name"], {}).get("dram_bytes_per_launch")
    step_alg_bytes = sum(k["algorithmic_bytes"] for k in prof) / prof_steps
    sweep_alg_bytes = 64.0 * n_edges_opt * P_local
    roofline = {
        "bound": "hbm", "kernel": top["name"], "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
        "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
        # bytes the kernel actually moved (ncu, per launch) over the live launch time: the fused node
        # kernel moves fewer bytes than the op list's algorithmic bytes, so `frac` can exceed 1
        "dram_achieved": (traffic / (per_launch_ms * 1e-3) / 1e9) if traffic else None,
        "dram_frac": (traffic / (per_launch_ms * 1e-3) / 1e9 / peak_gbs) if traffic else None,
        # what bare kernels of k_node's own shape reach on this part (tools/mix_bw.cu, profiles/r02T_mix_bw.log:
        # 1 read : 3 writes per pattern over hundreds of 4 MB PLVs = 6.1-6.2 TB/s, 2:3 = 6.4, 4:3 = 6.6): the
        # copy peak above is a 1:1 mix over two linear streams
        "mix_ceiling": {"GBps": 6150.0, "what": "bare 1 read : 3 writes kernel of k_node's shape, measured "
                        "(profiles/r02_k_node_ceiling.md)",
                        "dram_frac_of_ceiling": (traffic / (per_launch_ms * 1e-3) / 1e9 / 6150.0) if traffic else None},
        "launches_per_step": top["launches"] / prof_steps, "kernel_share_of_step": top["total_ms"] / total_prof_ms,
        "algorithmic_bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_ms,
        # north_star's target: pass + one sweep against the HBM roofline (SURVEY 8d bytes: the sweep counts the
        # two PLVs of every edge once)
        "whole_step": {"algorithmic_bytes": step_alg_bytes, "achieved": step_alg_bytes / (ms_per_step * 1e-3) / 1e9,
                       "frac": step_alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak_gbs,
                       "frac_of_8TBs": step_alg_bytes / (ms_per_step * 1e-3) / 8e12},
        "whole_pass": {"algorithmic_bytes": step_alg_bytes - sweep_alg_bytes,
                       "achieved": (step_alg_bytes - sweep_alg_bytes) / (pass_ms * 1e-3) / 1e9,
                       "frac": (step_alg_bytes - sweep_alg_bytes) / (pass_ms * 1e-3) / 1e9 / peak_gbs},
        "kernels": [{"name": k["name"], "share": k["total_ms"] / total_prof_ms,
                     "GBps": (k["algorithmic_bytes"] / (k["total_ms"] * 1e-3) / 1e9) if k["total_ms"] > 0 else None}
                    for k in prof[:8]],
        "kernels_note": "shares are from the profiling pass (graphs bypassed, the pipelined optimiser's two "
                        "streams serialised), not from the overlapped timed run",
    }

    # ---- the sweeps on their own -----------------------------------------------------------------
    scheme_names = {0: "streamed Taylor-model Brent: all edges of a level in lockstep; a pass over rho (8 B/pattern, HBM) yields the "
                       "objective at up to 4 points + 12 power sums, later requests inside the model's radius cost no pass",
                    1: "on chip: one thread block per edge, coefficients in shared memory",
                    2: "on chip: one thread-block cluster per edge, rho in distributed shared memory",
                    3: "pipelined: a streaming producer turns PLVs into rho (HBM-bound) while one thread-block cluster "
                       "per edge runs the search from shared memory"}

    def time_sweep(op_list, schedule):
        n_opt = int((op_list[0][:, 0] == 5).sum())
        runs = []
        for _ in range(2):  # same starting point twice: the first run also compiles the list and allocates
            engine.set_branch_lengths_to_constant(0.1)
            engine.reset_optimization_count()
            pass_step()
            barrier()
            f0 = engine.stats()["objective_evaluations"]
            p0 = engine.stats()["objective_passes"]
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            engine.process_operations(*op_list)
            ev1.record(stream)
            barrier()
            runs.append((max_over_ranks(ev0.elapsed_time(ev1)), engine.stats()["objective_evaluations"] - f0,
                         engine.stats()["objective_passes"] - p0))
        ms, fevals, passes = min(runs)
        s = engine.stats()
        bl = engine.get_branch_lengths()
        pass_step()
        # HBM traffic as executed by the optimiser: the two PLVs of every edge once; the round scheme also
        # writes rho once and re-reads it for every objective evaluation; the pipelined scheme writes and reads it once
        executed = 64.0 * n_opt * P_local
        if s["optimizer_scheme"] == 0:  # rho written once, re-read once per streamed pass
            executed += 8.0 * ((passes if passes > 0 else fevals) + n_opt) * P_local
        elif s["optimizer_scheme"] == 3:
            executed += 16.0 * n_opt * P_local
        return {"schedule": schedule, "optimizer": scheme_names[s["optimizer_scheme"]],
                "optimizer_scheme": int(s["optimizer_scheme"]),
                "cluster_size": s["optimizer_cluster_size"], "cluster_threads": s["optimizer_cluster_threads"],
                "edges_in_flight": s["optimizer_edges_in_flight"], "levels": s["levels_last"],
                "ms": ms, "ms_first_call": runs[0][0], "edges": n_opt, "objective_evaluations": fevals,
                "objective_passes": passes,
                "algorithmic_bytes": 64.0 * n_opt * P_local,
                "frac_of_hbm_peak": 64.0 * n_opt * P_local / (ms * 1e-3) / 1e9 / peak_gbs,
                "as_executed_bytes": executed,
                "as_executed_frac_of_hbm_peak": executed / (ms * 1e-3) / 1e9 / peak_gbs,
                "log_marginal_after": engine.get_log_marginal_likelihood()}, bl

    sweep, bl_batched = time_sweep(blo, "batched (all edges in one level, Brent)")
    sweep_reference_schedule = None
    if world > 1:
        # every rank took the same optimiser decisions: branch lengths bit-identical across ranks
        same = bool(np.array_equal(D.max_over_ranks_array(bl_batched), -D.max_over_ranks_array(-bl_batched)))
        parity["batched_sweep_branch_lengths_identical_on_all_ranks"] = same
    if not args.no_sweep:
        # the reference's own schedule (GPDAG::BranchLengthOptimization, gp_dag.cpp:52-176): a depth-first
        # Gauss-Seidel walk, one or two edges per dependency level, PLV updates between them - bound by
        # latency per edge, not by HBM (its `frac_of_hbm_peak` counts the optimiser's PLV reads only). On several
        # GPUs every objective evaluation ends in an exchange over NVLink inside the optimiser kernel.
        # Measured last and never allowed to take the other numbers down with it.
        try:
            sweep_reference_schedule, bl_gs = time_sweep(
                wl.ops("branch_length_optimization"),
                "GPDAG::BranchLengthOptimization (Gauss-Seidel; optimise + PLV updates interleaved, Brent)")
            if world > 1:
                same = bool(np.array_equal(D.max_over_ranks_array(bl_gs), -D.max_over_ranks_array(-bl_gs)))
                parity["gauss_seidel_sweep_branch_lengths_identical_on_all_ranks"] = same
        except Exception as exc:  # noqa: BLE001
            sweep_reference_schedule = {"error": str(exc)[:300]}
    if world > 1:
        parity["ok"] = bool(parity["shard_additivity_max_rel_err"] <= 1e-12 and parity["counts_identical_on_all_ranks"]
                            and all(v for k, v in parity.items() if k.endswith("identical_on_all_ranks")))
        parity["ok"] = bool(D.max_over_ranks(0.0 if parity["ok"] else 1.0) == 0.0)
        checks_ok = checks_ok and parity["ok"]

    try:
        st_end = engine.stats()
        collectives = {"all_reduces_so_far": int(st_end["collective_calls"]),
                       "over_nvlink_peer_memory": int(st_end["peer_collective_calls"]),
                       "note": "scalar all-reduces; peer memory = one k_peer_allreduce launch each, the rest ncclAllReduce"}
    except Exception as exc:  # noqa: BLE001
        collectives = {"error": str(exc)[:300]}
    hbm_bytes = int(st["device_bytes_in_use"])
    plvs_resident = int(st["plvs_resident"])
    engine.close()

    # ---- CPU baseline + parity: the reference's own engine on this box's host cores (rank 0, N = 1) -------
    cpu_baseline = None
    if rank == 0 and world == 1 and extras and not args.no_cpu_baseline:
        sample = cpu_sample_size(wl)
        kind, times, cpu_out, sub = cpu_step(wl, sample, 2)
        best = times[np.argmin(times.sum(axis=1))]
        sec = float(best.sum())
        cpu_baseline = {"value": wl.updates_per_pass() * sample / sec, "unit": UNIT, "cores": 1, "kind": kind,
                        "pass_only_value": wl.updates_per_pass() * sample / float(best[0]),
                        "sample": f"first {sample} of {P_local} site patterns, same DAG and op lists, best of 2 "
                                  f"steps (pass {best[0]:.2f} s + batched sweep {best[1]:.2f} s); host has "
                                  f"{os.cpu_count()} cores, the reference GPEngine uses 1"}
        parity = parity_against_cpu(wl, sub, cpu_out, local_rank)
        cpu_out["engine"].close()
        checks_ok = checks_ok and parity["ok"]

    line = None
    if rank == 0:
        shard = f"{name}/shard-{P_local}-of-{P_local * 8 if name == BENCH_WORKLOAD and args.patterns is None else P_total}-patterns"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": shard, **dag.summary(), "patterns_per_gpu": P_local, "patterns_total": P_total,
                       "step": STEP, "levels": pass_levels + 1, "pass_levels": pass_levels,
                       "plvs_resident": plvs_resident, "hbm_bytes_in_use": hbm_bytes,
                       "patterns": "distinct columns (duplicates merged into weights, SURVEY 8d)",
                       "l2": "inputs larger than L2: the working set (resident PLVs) is ~1000x the 126 MB L2",
                       "parallelism": f"patterns sharded over {world} GPU(s)", "setup_s": time.time() - t_setup},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "h2d_ms_per_step": upload_ms,
                    "h2d_GBps": h2d / (upload_ms * 1e-3) / 1e9,
                    "h2d_note": "upload timed on its own events (includes validating the symbols on the device and "
                                "the read-back of the total weight)"},
            "gpu_launches": int(launches),
            "pass_only": {"value": updates / (pass_ms * 1e-3), "unit": UNIT, "ms_per_step": pass_ms,
                          "gpu_launches": int(pass_launches), "log_marginal": log_marginal},
            "batched_sweep_ms": sweep["ms"], "batched_sweep_scheme": sweep_scheme,
            "parity": parity,
            "collectives": collectives,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "sweep": sweep,
            "sweep_reference_schedule": sweep_reference_schedule,
            "single_gpu_same_shard": standalone,
            "log_marginal": log_marginal,
            "full_pass_ms": pass_ms,
        }
    return line, checks_ok


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from bito_b200 import distributed as D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GP engine has no CPU fallback")
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun")
    torch.cuda.set_device(local_rank)
    D.init("nccl")

    name = args.workload or BENCH_WORKLOAD
    line, ok = measure(args, name, rank, world, local_rank)
    if world == 1 and args.workload is None and not args.no_config3:
        # BASELINE.json configs[3] (the "single B200" shape) beside the weak-scaling shard, same run
        other, ok3 = measure(args, SINGLE_B200_WORKLOAD, rank, world, local_rank, extras=True)
        ok = ok and ok3
        if line is not None and other is not None:
            line["config3_single_b200"] = {k: other[k] for k in ("value", "unit", "ms_per_step", "e2e", "config", "pass_only",
                                                                 "parity", "roofline", "sweep", "sweep_reference_schedule",
                                                                 "log_marginal")}
    if world == 1 and rank == 0 and args.workload is None and not args.no_small:
        try:
            line["small_configs"] = small_configs(local_rank)
            ok = ok and all(v["ok"] for v in line["small_configs"].values())
        except Exception as exc:  # noqa: BLE001
            line["small_configs"] = {"error": str(exc)[:300]}
            ok = False
    if rank == 0:
        line["checks_ok"] = bool(ok)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("bench.py: a parity check failed (see the `parity` / `small_configs` fields of the line above)")


if __name__ == "__main__":
    main()
