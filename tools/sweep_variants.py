"""Times one branch-length sweep of a bench workload under each optimiser scheme and checks that they agree:
    python tools/sweep_variants.py [workload] [patterns] [gauss_seidel]
Schemes (environment read when an engine is created): 0 = rounds (rho streamed from HBM every objective
evaluation), C or CxT = one thread-block cluster of C blocks of T (256 | 512 | 1024) threads per edge reading the PLVs,
p / pC / pCxT / pCxT/R = the pipelined cluster scheme (streaming producer + clusters reading rho; R edges per ring
half), auto = the engine's own per-level choice.
The workload is built once; each engine is destroyed before the next is created (they fill the HBM)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bito_b200 import _lib  # noqa: E402
from bito_b200.gp_engine import GPEngine  # noqa: E402
from bito_b200.synthetic import make_named_workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "synthetic-200taxa-100kpat-1000trees"
patterns = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "-" else None
gauss_seidel = "gauss_seidel" in sys.argv
variants = [v for v in os.environ.get("SWEEP_VARIANTS", "0,8,16,auto").split(",") if v]
wl = make_named_workload(name, pattern_count=patterns)
dag = wl.dag
flags = _lib.FLAG_NO_LOGLIK_MATRIX if dag.edge_count * wl.pattern_count * 8 > 16e9 else 0
pop = wl.ops("populate_plvs")
sweep_ops = wl.ops("branch_length_optimization" if gauss_seidel else "batched_branch_length_optimization")
stream = torch.cuda.Stream()  # not torch's default stream: its handle is NULL = "the engine's own stream" to set_stream
torch.cuda.set_stream(stream)
print(f"# {name}: P={wl.pattern_count} nodes={dag.node_count} edges={dag.edge_count} "
      f"sweep={'gauss-seidel' if gauss_seidel else 'batched'} ({sweep_ops[0].shape[0]} ops)", flush=True)
print("| variant | scheme | cluster | threads | edges in flight | sweep ms (best of 3) | evals | passes | max abs dBL vs first | edges > 1e-9 / > 1e-6 | log marginal after |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
first_bl = None
for v in variants:
    for k in ("BITO_GP_OPT_CLUSTER", "BITO_GP_OPT_CLUSTER_THREADS", "BITO_GP_OPT_SCHEME", "BITO_GP_OPT_RING_EDGES"):
        os.environ.pop(k, None)
    for k in [k for k in os.environ if k.startswith("BITO_GP_")]:
        os.environ.pop(k)
    v, *extra = v.split("@")  # "...@KEY=VAL": also set BITO_GP_KEY=VAL for this variant
    for kv in extra:
        k, _, val = kv.partition("=")
        os.environ["BITO_GP_" + k] = val
    label = "@".join([v] + extra)
    shape = v
    if v.startswith("p"):  # pipelined cluster scheme: "p", "pC", "pCxT", "pCxT/R" (R = edges per ring half)
        os.environ["BITO_GP_OPT_SCHEME"] = "3"
        shape, _, ring = v[1:].partition("/")
        if ring:
            os.environ["BITO_GP_OPT_RING_EDGES"] = ring
    if shape == "0":
        os.environ["BITO_GP_OPT_SCHEME"] = "0"
    elif shape not in ("auto", ""):  # "C" or "CxT": cluster size, threads per block
        c, _, t = shape.partition("x")
        os.environ["BITO_GP_OPT_CLUSTER"] = c
        if t:
            os.environ["BITO_GP_OPT_CLUSTER_THREADS"] = t
    with GPEngine(wl.symbols, wl.weights, wl.site_count, dag.node_count, dag.edge_count, sbn_prior=wl.sbn_prior,
                  unconditional_node_probabilities=wl.unconditional, inverted_sbn_prior=wl.inverted, flags=flags) as eng:
        eng.set_stream(stream.cuda_stream)
        best, evals = None, 0
        for _ in range(3):
            eng.set_branch_lengths_to_constant(0.1)
            eng.reset_optimization_count()
            eng.process_operations(*pop)
            f0 = eng.stats()["objective_evaluations"]
            p0 = eng.stats()["objective_passes"]
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.process_operations(*sweep_ops)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
            evals = eng.stats()["objective_evaluations"] - f0
            passes = eng.stats()["objective_passes"] - p0
        bl = eng.get_branch_lengths()
        eng.process_operations(*pop)
        eng.process_operations(*wl.ops("marginal_likelihood"))
        st = eng.stats()
        if first_bl is None:
            first_bl = bl
        print(f"| {label} | {st['optimizer_scheme']} | {st['optimizer_cluster_size']} | {st['optimizer_cluster_threads']} | {st['optimizer_edges_in_flight']} | "
              f"{best:.3f} | {evals} | {passes} | {np.max(np.abs(bl - first_bl)):.3e} | "
              f"{int(np.sum(np.abs(bl - first_bl) > 1e-9))} / {int(np.sum(np.abs(bl - first_bl) > 1e-6))} | "
              f"{eng.get_log_marginal_likelihood():.6f} |",
              flush=True)
