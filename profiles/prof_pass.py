"""Runs the bench workload's likelihood pass a few times WITHOUT CUDA graphs so ncu sees plain
kernel launches:  ncu ... python profiles/prof_pass.py [workload] [patterns] [passes]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bito_b200 import _lib
from bito_b200.gp_engine import GPEngine
from bito_b200.synthetic import make_named_workload

name = sys.argv[1] if len(sys.argv) > 1 else "synthetic-1000taxa-1Mpat-5000trees"
patterns = int(sys.argv[2]) if len(sys.argv) > 2 else None
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
with_sweep = len(sys.argv) > 4 and sys.argv[4] == "sweep"
with_gs = len(sys.argv) > 4 and sys.argv[4] == "gs"  # the reference's Gauss-Seidel sweep (on-chip optimiser kernels)
wl = make_named_workload(name, pattern_count=patterns)
dag = wl.dag
eng = GPEngine(wl.symbols, wl.weights, wl.site_count, dag.node_count, dag.edge_count, sbn_prior=wl.sbn_prior,
               unconditional_node_probabilities=wl.unconditional, inverted_sbn_prior=wl.inverted,
               flags=_lib.FLAG_NO_CUDA_GRAPHS)
for _ in range(passes):
    eng.process_operations(*wl.ops("populate_plvs"))
    eng.process_operations(*wl.ops("compute_likelihoods"))
if with_sweep:
    eng.process_operations(*wl.ops("batched_branch_length_optimization"))
if with_gs:
    eng.process_operations(*wl.ops("branch_length_optimization"))
print("log marginal", eng.get_log_marginal_likelihood(), eng.stats())
