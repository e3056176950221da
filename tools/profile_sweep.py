"""Per-kernel CUDA-event profile of one branch-length sweep on the bench workload (batched list, or the reference's
Gauss-Seidel list with `gauss_seidel` as the last argument)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bito_b200.gp_engine import GPEngine  # noqa: E402
from bito_b200.synthetic import make_named_workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] != "gauss_seidel" else "synthetic-1000taxa-1Mpat-5000trees"
patterns = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] not in ("-", "gauss_seidel") else None
wl = make_named_workload(name, pattern_count=patterns)
dag = wl.dag
eng = GPEngine(wl.symbols, wl.weights, wl.site_count, dag.node_count, dag.edge_count, sbn_prior=wl.sbn_prior,
               unconditional_node_probabilities=wl.unconditional, inverted_sbn_prior=wl.inverted)
eng.process_operations(*wl.ops("populate_plvs"))
blo = wl.ops("branch_length_optimization" if "gauss_seidel" in sys.argv else "batched_branch_length_optimization")
eng.reset_optimization_count()
eng.process_operations(*blo)  # warm-up (allocations)
eng.set_branch_lengths_to_constant(0.1)
eng.process_operations(*wl.ops("populate_plvs"))
eng.reset_optimization_count()
eng.set_profiling(True)
eng.reset_kernel_profile()
eng.process_operations(*blo)
prof = eng.kernel_profile()
eng.set_profiling(False)
tot = sum(k["total_ms"] for k in prof)
for k in sorted(prof, key=lambda k: -k["total_ms"]):
    print(f"{k['name']:20s} launches {k['launches']:6d}  total {k['total_ms']:8.3f} ms  share {k['total_ms'] / tot:.3f}  "
          f"avg {1e3 * k['total_ms'] / k['launches']:8.1f} us")
print("sum of kernel times", tot, "ms; last_process_ms", eng.stats()["last_process_ms"])
