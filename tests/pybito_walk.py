"""Drives the `bito` Python module (whichever build is first on sys.path) through the GP surface a bito user
touches - the reference's own test (/root/reference/test/test_bito.py:166-184) plus the calls test/nni_search.py
makes on a gp_instance - and prints one JSON object. Run as a subprocess by tests/test_pybito_gpu.py, once per
build (CUDA engine / reference CPU engine), so the two builds never share a process."""
import json
import os
import sys

import numpy as np

import bito

fasta, newick, workdir = sys.argv[1], sys.argv[2], sys.argv[3]
threshold = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-40
light = len(sys.argv) > 5 and sys.argv[5] == "light"  # large DAGs: no run to convergence
import time
t_start = time.time()
out = {"backend": bito.gp_engine_backend}
inst = bito.gp_instance(os.path.join(workdir, "mmapped_plv_pybito.data"))
inst.read_fasta_file(fasta)
inst.read_newick_file(newick)
t_read = time.time()
inst.make_dag()
t_dag = time.time()
inst.make_gp_engine(rescaling_threshold=threshold)
dag, engine = inst.get_dag(), inst.get_gp_engine()
out["seconds"] = {"read": t_read - t_start, "make_dag": t_dag - t_read, "make_gp_engine": time.time() - t_dag}
out["dag"] = [dag.node_count(), dag.edge_count(), dag.taxon_count(), dag.topology_count()]
out["engine"] = [engine.node_count(), engine.plv_count(), engine.edge_count()]
out["init_branch_lengths"] = inst.get_branch_lengths().tolist()
t0 = time.time()
inst.populate_plvs()
inst.compute_likelihoods()
out["seconds"]["first_pass_planning_and_running"] = time.time() - t0
out["pass_per_pcsp_llh"] = inst.get_per_pcsp_log_likelihoods().tolist()
inst.hot_start_branch_lengths()
out["hot_start_branch_lengths"] = inst.get_branch_lengths().tolist()
t0 = time.time()
inst.estimate_branch_lengths(1e-3, 1, True)  # tol, max_iter, quiet: one sweep (the north-star tolerances)
out["seconds"]["estimate_branch_lengths_one_iteration"] = time.time() - t0
out["estimated_branch_lengths"] = inst.get_branch_lengths().tolist()
inst.populate_plvs()
inst.compute_likelihoods()
out["estimated_per_pcsp_llh"] = inst.get_per_pcsp_log_likelihoods().tolist()
out["log_marginal"] = inst.get_log_marginal_likelihood()
inst.estimate_sbn_parameters()
out["sbn_parameters"] = inst.get_sbn_parameters().tolist()
pcsp_of_edge = inst.build_edge_idx_to_pcsp_map()
out["edge_pcsps"] = {str(k): v.pcsp_to_string() for k, v in sorted(pcsp_of_edge.items())}
out["node_bitsets"] = sorted(b.subsplit_to_string() for b in dag.build_set_of_node_bitsets())
trees = inst.currently_loaded_trees_with_gp_branch_lengths()
out["first_tree"] = trees.trees[0].to_newick_topology()
out["first_tree_branch_lengths"] = np.asarray(trees.trees[0].branch_lengths).tolist()
# bitset / nni value types
some_edge = next(iter(pcsp_of_edge.values()))
parent, child = some_edge.pcsp_get_parent_subsplit(), some_edge.pcsp_get_child_subsplit()
out["pcsp_roundtrip"] = bito.pcsp(parent, child) == some_edge and dag.contains_edge(some_edge)
out["edge_id_roundtrip"] = all(dag.get_edge_id(v).value() == k for k, v in pcsp_of_edge.items())
try:
    inst.get_likelihood_tree_engine()
    out["beagle"] = "no error"
except RuntimeError as exc:
    out["beagle"] = "RuntimeError" if "BEAGLE" in str(exc) else str(exc)
# a longer optimisation, as test_bito.py runs it
if not light:
    inst.estimate_branch_lengths(1e-3, 100, True)
    inst.populate_plvs()
    inst.compute_likelihoods()
out["converged_log_marginal"] = inst.get_log_marginal_likelihood()
print("PYBITO_WALK " + json.dumps(out))
