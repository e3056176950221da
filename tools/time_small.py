"""Latency of the real-data shapes (BASELINE.json configs[0..2]: DS1, fluA) on the CUDA engine next
to the unmodified reference CPU engine (oracle/_ref), same op lists (the reference GPDAG's own, from
tests/golden/*.npz): one full pass (PopulatePLVs + ComputeLikelihoods) and one Gauss-Seidel
branch-length sweep (BranchLengthOptimization + PopulatePLVs + MarginalLikelihood, the body of
GPInstance::EstimateBranchLengths, gp_instance.cpp:241-308). These alignments have 238-934 site
patterns: the GPU path is bound by launch latency here, not by HBM.

    python tools/time_small.py [case ...]          (needs a GPU; the reference leg needs oracle/_ref)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gp_cases as C  # noqa: E402


def best_of(fn, sync, repeats):
    times = []
    for _ in range(repeats):
        sync()
        t0 = time.perf_counter()
        fn()
        sync()
        times.append(time.perf_counter() - t0)
    return min(times) * 1e3


def main():
    cases = sys.argv[1:] or ["ds1", "ds1_config1", "fluA"]
    from oracle import ref_engine
    print("| case | P | edges | engine | pass ms | sweep ms | evals/sweep | log marginal after |")
    print("|---|---|---|---|---|---|---|---|")
    for name in cases:
        fx = C.Fixture(name)
        a = fx.engine_args(0)
        pop, lik = fx.ops("populate_plvs"), fx.ops("compute_likelihoods")
        blo, marg = fx.ops("branch_length_optimization"), fx.ops("marginal_likelihood")
        engines = []
        try:
            engines.append(("cuda", C.make_cuda(fx)))
        except Exception as exc:  # no GPU here: reference leg only
            print(f"(cuda engine unavailable: {str(exc)[:80]})", file=sys.stderr)
        if ref_engine.available():
            ref = ref_engine.RefEngine.from_arrays(a["symbols"], a["weights"], a["site_count"], a["node_count"],
                                                   a["edge_count"], a["q"], a["unconditional"], a["inverted"],
                                                   a["rescaling_threshold"])
            ref.set_branch_lengths(fx["initial_branch_lengths"])
            engines.append(("reference-cpu", ref))
        for label, e in engines:
            sync = e.synchronize if hasattr(e, "synchronize") else (lambda: None)

            def one_pass():
                e.process_operations(*pop)
                e.process_operations(*lik)

            def one_sweep():
                e.process_operations(*blo)
                e.process_operations(*pop)
                e.process_operations(*marg)

            one_pass()
            pass_ms = best_of(one_pass, sync, 5)
            e.set_optimization_method("brent")
            e.reset_optimization_count()
            e.process_operations(*pop)
            evals0 = e.stats()["objective_evaluations"] if hasattr(e, "stats") else 0
            sweeps = []
            for s in range(3):
                e.set_branch_lengths(fx["initial_branch_lengths"])
                e.reset_optimization_count()
                e.process_operations(*pop)
                sync()
                t0 = time.perf_counter()
                one_sweep()
                sync()
                sweeps.append((time.perf_counter() - t0) * 1e3)
            evals = (e.stats()["objective_evaluations"] - evals0) // 3 if hasattr(e, "stats") else ""
            lm = e.get_log_marginal_likelihood() if hasattr(e, "get_log_marginal_likelihood") else e.log_marginal_likelihood()
            print(f"| {name} | {a['symbols'].shape[1]} | {a['edge_count']} | {label} | {pass_ms:.3f} | "
                  f"{min(sweeps):.3f} | {evals} | {lm:.6f} |")
            e.close()


if __name__ == "__main__":
    main()
