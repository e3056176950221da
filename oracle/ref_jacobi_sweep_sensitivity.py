"""TEST INFRASTRUCTURE ONLY. How reproducible is a cold-start batched (Jacobi) Brent sweep?

Runs one OptimizeBranchLength per edge of the bench DAG (200 taxa, 8369 edges, first N patterns of the
bench alignment, every branch length 0.1) through two builds of the UNMODIFIED reference engine
(oracle/_ref: -O3; oracle/_ref/variant: -O2 -march=native -ffp-contract=fast, `make -C oracle refvar`)
and through the plain-C oracle, and counts the edges whose optimised length differs by more than 1e-6.
All edges start at log 0.1, so Brent's first three evaluations coincide and its first parabolic step
sits on an acceptance boundary for ~0.1 % of the edges: 1e-11 relative noise in the objective flips the
decision and the search ends at another point inside Brent's own 2^-9 tolerance.

    make -C oracle -j8 ref refvar port && python oracle/ref_jacobi_sweep_sensitivity.py [N]
"""
import sys, numpy as np, importlib
sys.path.insert(0, "/root/repo")
from bito_b200.synthetic import make_named_workload
from oracle import ref_engine
from oracle.port_engine import PortEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
wl = make_named_workload("synthetic-200taxa-100kpat-1000trees").subsample(n)
dag = wl.dag
pop, lik, blo = wl.ops("populate_plvs"), wl.ops("compute_likelihoods"), wl.ops("batched_branch_length_optimization")
res = {}
for tag, path in (("O3", "/root/repo/oracle/_ref/libbito_gp_ref.so"), ("native", "/root/repo/oracle/_ref/variant/libbito_gp_ref.so")):
    ref_engine.LIB_PATH = path
    ref_engine._lib = None
    e = ref_engine.RefEngine.from_arrays(wl.symbols, wl.weights, wl.site_count, dag.node_count, dag.edge_count, wl.sbn_prior, wl.unconditional, wl.inverted)
    e.process_operations(*pop); e.process_operations(*lik); e.process_operations(*blo)
    res[tag] = e.branch_lengths().copy(); e.close()
p = PortEngine(wl.symbols, wl.weights, wl.site_count, dag.node_count, dag.edge_count, wl.sbn_prior, wl.unconditional, wl.inverted)
p.process_operations(*pop); p.process_operations(*lik); p.process_operations(*blo)
res["port"] = p.branch_lengths()
for a, b in (("O3","native"),("O3","port"),("native","port")):
    d = np.abs(res[a]-res[b]); print(a, b, "n>1e-6:", int((d>1e-6).sum()), "max", d.max(), "edges", np.nonzero(d>1e-6)[0][:10])
