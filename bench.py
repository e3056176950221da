#!/usr/bin/env python
"""bench.py — the GP likelihood hot path on B200, one JSON line per run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload NAME] [--patterns P]

One *step* = one full-DAG likelihood pass: GPDAG::PopulatePLVs + GPDAG::ComputeLikelihoods
(gp_instance.cpp:231-235) through GPEngine::ProcessOperations, over one synthetic alignment.
metric = GP pattern x edge PLV updates / s  =  2 (E - R) P_total / pass time  (SURVEY.md 8d).

Every N runs BASELINE.json configs[4] (1000 taxa, DAG from 5000 trees) weak-scaled at 125k
site patterns per GPU (N = 8 is the full 1M-pattern alignment; one shard fills 127 GB of
HBM), patterns sharded, per-edge scalars all-reduced with NCCL inside the engine. At N = 1
the line also carries configs[3] (200 taxa x 100k patterns, DAG from 1000 trees, "single
B200") under `config3_single_b200`, measured in the same run.

  value  : device-resident throughput (alignment already in HBM), CUDA events, max over ranks
  e2e    : same metric through the C-ABI with HOST buffers: every step uploads the alignment,
           weights and branch lengths from pinned memory and reads back per-edge
           log-likelihoods + the marginal
  roofline: dominant kernel, algorithmic bytes / event-timed kernel time
  cpu_baseline: the UNMODIFIED reference CPU GPEngine (oracle/_ref; falls back to the
           plain-C port) on a bounded pattern sample of the same workload, 1 core
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--patterns", type=int, default=None, help="patterns per GPU (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-config3", action="store_true", help="N = 1: skip the extra configs[3] measurement")
    ap.add_argument("--ref-procs", type=int, default=0, help="--impl reference: processes (default: all cores, <= 32)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


def cpu_pass_seconds(workload, sample_patterns, repeats):
    """Times PopulatePLVs + ComputeLikelihoods on the reference CPU GPEngine (1 thread: the
    reference engine is single-threaded by construction) over the first `sample_patterns`."""
    from oracle import ref_engine
    sub = workload.subsample(sample_patterns)
    pop, lik = workload.ops("populate_plvs"), workload.ops("compute_likelihoods")
    kind = "reference"
    try:
        if not ref_engine.available():
            raise RuntimeError("oracle/_ref not built")
        eng = ref_engine.RefEngine.from_arrays(sub.symbols, sub.weights, sub.site_count, workload.dag.node_count,
                                               workload.dag.edge_count, workload.sbn_prior, workload.unconditional,
                                               workload.inverted)
        eng.process_operations(*pop)  # warm-up (page-faults the mmap'd PLV file in)
        eng.process_operations(*lik)
        times = eng.time_operations(*pop, repeats=repeats) + eng.time_operations(*lik, repeats=repeats)  # per pass
    except Exception:  # the prebuilt reference is missing: fall back to the plain-C port
        from oracle.port_engine import PortEngine
        kind = "port"
        eng = PortEngine(sub.symbols, sub.weights, sub.site_count, workload.dag.node_count,
                         workload.dag.edge_count, workload.sbn_prior, workload.unconditional, workload.inverted)
        eng.process_operations(*pop)
        eng.process_operations(*lik)
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            eng.process_operations(*pop)
            eng.process_operations(*lik)
            times.append(time.perf_counter() - t0)
        times = np.array(times)
    marginal = eng.log_marginal_likelihood()
    eng.close()
    return kind, np.asarray(times), marginal


def cpu_sample_size(workload):
    # ~1.5e8 pattern x edge updates per timed pass (a few seconds at the reference's ~5e7/s),
    # and a reference mmap file (6 * (2 (N + 16) + 16) PLVs of 32 P bytes) under ~6 GB.
    per_pattern = workload.updates_per_pass()
    by_time = int(1.5e8 // per_pattern)
    by_memory = int(6e9 // (6 * (2 * (workload.dag.node_count + 16) + 16) * 32))
    return max(64, min(workload.pattern_count, by_time, by_memory))


def _reference_worker(job):
    """One process of the reference arm: the unmodified single-threaded reference GPEngine on its own
    slice of the workload's site patterns (the path is pattern-parallel)."""
    name, index, patterns, warmup, steps, gate = job
    from bito_b200.synthetic import make_named_workload
    wl = make_named_workload(name, rank=1000 + index, pattern_count=patterns)
    if gate is not None:
        gate.wait()  # all workers enter the timed passes together
    t0 = time.time()
    kind, times, marginal = cpu_pass_seconds(wl, wl.pattern_count, warmup + steps)
    return dict(kind=kind, times=times.tolist(), marginal=marginal, began=t0, updates=wl.updates_per_pass(),
                summary=wl.dag.summary())


def run_reference(args, rank, world):
    """bench.py --impl reference: the reference's own CPU implementation of the path on the box's host
    cores. The reference GPEngine is single-threaded, so "all the host threads it can use" is one
    process per core, each running the unmodified engine over its own slice of the site patterns."""
    if rank != 0:
        return
    import multiprocessing as mp
    name = args.workload or BENCH_WORKLOAD
    procs = args.ref_procs or min(32, os.cpu_count() or 1)
    # per process: ~1 s of CPU work per step and ~1.3 GB of touched PLV pages (the reference's PLVs are
    # an mmap'd file); all processes together stay under ~21 GB
    patterns = args.patterns or max(512, min(1024 if "1000taxa" in name else 2048, 32768 // procs))
    if procs == 1:
        results = [_reference_worker((name, 0, patterns, args.warmup, args.steps, None))]
    else:
        ctx = mp.get_context("spawn")
        with ctx.Manager() as manager, ctx.Pool(procs) as pool:
            gate = manager.Barrier(procs)
            results = pool.map(_reference_worker,
                               [(name, i, patterns, args.warmup, args.steps, gate) for i in range(procs)], chunksize=1)
    # every process repeats `steps` timed passes; aggregate throughput = all updates / slowest process
    per_proc = [float(np.sum(r["times"][args.warmup:])) for r in results]
    sec = max(per_proc) / args.steps
    value = results[0]["updates"] * patterns * procs / sec
    kind = "reference" if all(r["kind"] == "reference" for r in results) else "port"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, **results[0]["summary"], "patterns_per_step": patterns * procs,
                   "step": "PopulatePLVs+ComputeLikelihoods on the reference CPU GPEngine"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": f"{procs} processes x {patterns} site patterns of the workload per step (same "
                                   "DAG and op lists); the reference GPEngine is single-threaded, so each core "
                                   "runs its own unmodified engine on a disjoint pattern slice; slowest process "
                                   f"timed; 1-process rate {results[0]['updates'] * patterns / (per_proc[0] / args.steps):.3e}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "log_marginal": results[0]["marginal"],
    }
    print(json.dumps(line))


def measure(args, name, rank, world, local_rank, extras=True):
    """One workload on this rank's GPU: device-resident pass, e2e pass, per-kernel roofline, sweep,
    CPU baseline. Returns the JSON line (rank 0) or None."""
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

    # Pinned host staging for the e2e leg (and the initial upload).
    sym_pinned = torch.from_numpy(wl.symbols).pin_memory()
    w_pinned = torch.from_numpy(wl.weights).pin_memory()
    bl_pinned = torch.full((dag.edge_count,), 0.1, dtype=torch.float64).pin_memory()
    flags = _lib.FLAG_NO_LOGLIK_MATRIX if dag.edge_count * P_local * 8 > 16e9 else 0
    engine = GPEngine(sym_pinned.numpy(), w_pinned.numpy(), wl.site_count * world, dag.node_count, dag.edge_count,
                      sbn_prior=wl.sbn_prior, unconditional_node_probabilities=wl.unconditional,
                      inverted_sbn_prior=wl.inverted, device=local_rank, flags=flags)
    stream = torch.cuda.current_stream()
    engine.set_stream(stream.cuda_stream)

    barrier, max_over_ranks = D.barrier, D.max_over_ranks

    def device_step():
        engine.process_operations(*pop)
        engine.process_operations(*lik)

    def e2e_step():
        engine.set_site_patterns(sym_pinned.numpy(), w_pinned.numpy())
        engine.set_branch_lengths(bl_pinned.numpy())
        engine.process_operations(*pop)
        engine.process_operations(*lik)
        ll = engine.get_per_gpcsp_log_likelihoods()
        return ll, engine.get_log_marginal_likelihood()

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

    # ---- N > 1: the same shard on one GPU with no communicator (weak-scaling reference) ----------
    standalone = None
    if world > 1:
        ms_alone, _ = timed(device_step, max(3, args.steps // 2), 3)
        standalone = {"value_per_gpu": wl.updates_per_pass() * P_local / (ms_alone * 1e-3), "unit": UNIT,
                      "ms_per_step": ms_alone,
                      "what": "this rank's shard (same DAG, same patterns per GPU) before joining the NCCL "
                              "communicator; max over ranks"}
        D.connect_engine(engine)

    # ---- device-resident throughput (`value`) ---------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_per_step, launches = timed(device_step, args.steps, max(3, args.warmup))
    clocks = sampler.stop()
    updates = wl.updates_per_pass() * P_total
    value = updates / (ms_per_step * 1e-3)
    log_marginal = engine.get_log_marginal_likelihood()
    st = engine.stats()
    alg_bytes_pass = None

    # ---- e2e through the C-ABI with host buffers --------------------------------------------------
    e2e_ms, _ = timed(e2e_step, max(2, args.steps // 2), 1)
    e2e_value = updates / (e2e_ms * 1e-3)
    h2d = wl.symbols.nbytes + wl.weights.nbytes + dag.edge_count * 8
    d2h = dag.edge_count * 8 + 8

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
    traffic_path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(traffic_path) and args.patterns is None:
        # ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel on this workload's
        # per-GPU shard (every rank moves the same bytes)
        with open(traffic_path) as f:
            traffic = json.load(f).get(name, {}).get(top["name"], {}).get("dram_bytes_per_launch")
    pass_alg_bytes = sum(k["algorithmic_bytes"] for k in prof) / prof_steps
    roofline = {
        "bound": "hbm", "kernel": top["name"], "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
        "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
        # bytes the kernel actually moved (ncu, per launch) over the live launch time: the fused node
        # kernel moves fewer bytes than the op list's algorithmic bytes, so `frac` can exceed 1
        "dram_achieved": (traffic / (per_launch_ms * 1e-3) / 1e9) if traffic else None,
        "dram_frac": (traffic / (per_launch_ms * 1e-3) / 1e9 / peak_gbs) if traffic else None,
        "launches_per_step": top["launches"] / prof_steps, "kernel_share_of_step": top["total_ms"] / total_prof_ms,
        "algorithmic_bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_ms,
        "whole_pass": {"algorithmic_bytes": pass_alg_bytes,
                       "achieved": pass_alg_bytes / (ms_per_step * 1e-3) / 1e9,
                       "frac": pass_alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak_gbs},
        "kernels": [{"name": k["name"], "share": k["total_ms"] / total_prof_ms,
                     "GBps": (k["algorithmic_bytes"] / (k["total_ms"] * 1e-3) / 1e9) if k["total_ms"] > 0 else None}
                    for k in prof[:6]],
    }

    # ---- one branch-length optimisation sweep (reported beside the pass; not part of `value`) ----
    sweep = None
    sweep_reference_schedule = None
    scheme_names = {0: "rounds: one launch per objective round over all edges of a level, rho (8 B/pattern) streamed from HBM",
                    1: "on chip: one thread block per edge, coefficients in shared memory",
                    2: "on chip: one thread-block cluster per edge, rho in distributed shared memory"}

    def time_sweep(op_list, schedule):
        n_edges_opt = int((op_list[0][:, 0] == 5).sum())  # gp_operation.OPTIMIZE_BRANCH_LENGTH
        runs = []
        for _ in range(2):  # same starting point twice: the first run also compiles the list and allocates
            engine.set_branch_lengths(bl_pinned.numpy())
            engine.reset_optimization_count()
            device_step()
            barrier()
            f0 = engine.stats()["objective_evaluations"]
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            engine.process_operations(*op_list)
            ev1.record(stream)
            barrier()
            runs.append((max_over_ranks(ev0.elapsed_time(ev1)), engine.stats()["objective_evaluations"] - f0))
        ms, fevals = min(runs)
        st = engine.stats()
        device_step()
        # HBM traffic as executed by the optimiser: the two PLVs of every edge once; the round scheme also
        # writes rho once and re-reads it for every objective evaluation
        executed = 64.0 * n_edges_opt * P_local + (8.0 * (fevals + n_edges_opt) * P_local if st["optimizer_scheme"] == 0 else 0.0)
        return {"schedule": schedule, "optimizer": scheme_names[st["optimizer_scheme"]],
                "cluster_size": st["optimizer_cluster_size"], "cluster_threads": st["optimizer_cluster_threads"],
                "edges_in_flight": st["optimizer_edges_in_flight"], "levels": st["levels_last"],
                "ms": ms, "ms_first_call": runs[0][0], "edges": n_edges_opt, "objective_evaluations": fevals,
                "algorithmic_bytes": 64.0 * n_edges_opt * P_local,
                "frac_of_hbm_peak": 64.0 * n_edges_opt * P_local / (ms * 1e-3) / 1e9 / peak_gbs,
                "as_executed_bytes": executed,
                "as_executed_frac_of_hbm_peak": executed / (ms * 1e-3) / 1e9 / peak_gbs,
                "log_marginal_after": engine.get_log_marginal_likelihood()}

    if not args.no_sweep:
        sweep = time_sweep(wl.ops("batched_branch_length_optimization"), "batched (all edges in one level, Brent)")
        # the reference's own schedule (GPDAG::BranchLengthOptimization, gp_dag.cpp:52-176): a depth-first
        # Gauss-Seidel walk, one or two edges per dependency level, PLV updates between them - bound by
        # latency per edge, not by HBM (its `frac_of_hbm_peak` counts the optimiser's PLV reads only). On several
        # GPUs every objective evaluation ends in an exchange over NVLink inside the optimiser kernel.
        # Measured last and never allowed to take the pass numbers down with it.
        try:
            sweep_reference_schedule = time_sweep(
                wl.ops("branch_length_optimization"),
                "GPDAG::BranchLengthOptimization (Gauss-Seidel; optimise + PLV updates interleaved, Brent)")
        except Exception as exc:  # noqa: BLE001
            sweep_reference_schedule = {"error": str(exc)[:300]}

    # ---- CPU baseline: the reference's own engine on this box's host cores (rank 0, N = 1) -------
    cpu_baseline = None
    if rank == 0 and world == 1 and extras and not args.no_cpu_baseline:
        sample = cpu_sample_size(wl)
        kind, times, cpu_marginal = cpu_pass_seconds(wl, sample, 3)
        sec = float(np.min(times))
        cpu_baseline = {"value": wl.updates_per_pass() * sample / sec, "unit": UNIT, "cores": 1, "kind": kind,
                        "sample": f"first {sample} of {P_local} site patterns, same DAG and op lists, best of 3 "
                                  f"passes ({sec:.2f} s each); host has {os.cpu_count()} cores, the reference "
                                  "GPEngine uses 1"}

    try:
        st_end = engine.stats()
        collectives = {"all_reduces_so_far": int(st_end["collective_calls"]),
                       "over_nvlink_peer_memory": int(st_end["peer_collective_calls"]),
                       "note": "scalar all-reduces; peer memory = one k_peer_allreduce launch each, the rest ncclAllReduce"}
    except Exception as exc:  # noqa: BLE001
        collectives = {"error": str(exc)[:300]}
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, **dag.summary(), "patterns_per_gpu": P_local, "patterns_total": P_total,
                       "step": "PopulatePLVs + ComputeLikelihoods (full-DAG likelihood pass)",
                       "levels": int(st["levels_last"]), "plvs_resident": int(st["plvs_resident"]),
                       "hbm_bytes_in_use": int(st["device_bytes_in_use"]),
                       "l2": "working set (resident PLVs) is far larger than the 126 MB L2",
                       "parallelism": f"patterns sharded over {world} GPU(s)", "setup_s": time.time() - t_setup},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "collectives": collectives,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "sweep": sweep,
            "sweep_reference_schedule": sweep_reference_schedule,
            "single_gpu_same_shard": standalone,
            "log_marginal": log_marginal,
            "full_pass_ms": ms_per_step,
        }
    engine.close()
    return line


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
    from bito_b200 import _lib
    from bito_b200 import distributed as D
    from bito_b200.gp_engine import GPEngine
    from bito_b200.synthetic import make_named_workload

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GP engine has no CPU fallback")
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torchrun")
    torch.cuda.set_device(local_rank)
    D.init("nccl")

    name = args.workload or BENCH_WORKLOAD
    line = measure(args, name, rank, world, local_rank)
    if world == 1 and args.workload is None and not args.no_config3:
        # BASELINE.json configs[3] (the "single B200" shape) beside the weak-scaling shard, same run
        other = measure(args, SINGLE_B200_WORKLOAD, rank, world, local_rank, extras=False)
        if line is not None and other is not None:
            line["config3_single_b200"] = {k: other[k] for k in ("value", "unit", "ms_per_step", "e2e", "config",
                                                                 "roofline", "sweep", "sweep_reference_schedule", "log_marginal")}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
