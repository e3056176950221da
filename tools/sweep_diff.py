"""Where does a batched Brent sweep on the bench DAG differ between the CUDA engine and the C oracle?"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bito_b200.gp_engine import GPEngine  # noqa: E402
from bito_b200.synthetic import make_named_workload  # noqa: E402
from oracle.port_engine import PortEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
wl = make_named_workload("synthetic-200taxa-100kpat-1000trees").subsample(n)
dag = wl.dag
pop, lik, blo = wl.ops("populate_plvs"), wl.ops("compute_likelihoods"), wl.ops("batched_branch_length_optimization")
cpu = PortEngine(wl.symbols, wl.weights, wl.site_count, dag.node_count, dag.edge_count, wl.sbn_prior,
                 wl.unconditional, wl.inverted)
gpu = GPEngine(wl.symbols, wl.weights, wl.site_count, dag.node_count, dag.edge_count, sbn_prior=wl.sbn_prior,
               unconditional_node_probabilities=wl.unconditional, inverted_sbn_prior=wl.inverted)
for e in (cpu, gpu):
    e.process_operations(*pop)
    e.process_operations(*lik)
    e.process_operations(*blo)
a, b = gpu.get_branch_lengths(), cpu.branch_lengths()
d = np.abs(a - b)
bad = np.nonzero(d > 1e-6)[0]
print(f"edges {d.size}, |dBL| > 1e-6: {bad.size}, max {d.max():.3e}, median {np.median(d):.3e}")
ops = blo[0]
by_edge = {int(r[3]): (int(r[1]), int(r[2])) for r in ops}  # gpcsp -> (leafward, rootward)
for e in bad[:12]:
    lw, rw = by_edge[int(e)]
    vals = []
    for t in (a[e], b[e]):
        bl = cpu.branch_lengths().copy()
        bl[e] = t
        cpu.set_branch_lengths(bl)
        vals.append(cpu.log_likelihood_and_derivatives(int(e), rw, lw)[0])
    print(f"edge {e}: gpu t={a[e]:.12g} cpu t={b[e]:.12g}  ll(gpu t)={vals[0]:.15g} ll(cpu t)={vals[1]:.15g} "
          f"diff {vals[0] - vals[1]:.3e}")
