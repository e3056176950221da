"""Generates tests/golden/*.npz from the UNMODIFIED reference GP path (oracle/_ref).

Run in the build container (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

Every fixture holds the reference's own inputs (SitePattern symbols/weights, priors, DAG
edges, GPDAG op lists — so PLV/edge/pattern indexing is the reference's by construction)
and the reference GPEngine's outputs on them. /root/reference does not exist on the GPU
box, so the GPU parity tests read these files instead.

Per-case protocol (mirrors GPInstance::PopulatePLVs/ComputeLikelihoods and
GPInstance::EstimateBranchLengths, /root/reference/src/gp_instance.cpp:231-308, 401-406):
  1. PopulatePLVs + ComputeLikelihoods                      -> pass_* outputs
  2. ResetOptimizationCount; `sweeps` x {BranchLengthOptimization, PopulatePLVs,
     MarginalLikelihood, IncrementOptimizationCount}        -> sweep_* outputs per method
  3. ComputeLikelihoods + OptimizeSBNParameters             -> sbn_q
"""
from __future__ import annotations

import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_engine import RefEngine  # noqa: E402

DATA = "/root/reference/data/"
LISTS = ["populate_plvs", "compute_likelihoods", "marginal_likelihood", "branch_length_optimization",
         "optimize_sbn_parameters", "rootward_pass", "leafward_pass"]

# name -> dict(fasta, newick, thresholds, branch lengths, methods to sweep with, sweeps)
CASES = {
    "hello": dict(fasta="hello.fasta", newick="hello_rooted.nwk", bl=[0, 0.22, 0.113, 0.15, 0.1],
                  thresholds=[1e-40],
                  # logspace_gradient_ascent is left out here: with step 1.0005 in log space its
                  # iteration is chaotic on this data (rounding noise grows to O(1) within the 1000
                  # iterations), so even two builds of the reference disagree. five_taxon covers it.
                  methods=["brent", "brent_with_gradients", "gradient_ascent", "newton"], sweeps=4),
    "hello_single_nucleotide": dict(fasta="hello_single_nucleotide.fasta", newick="hello_rooted.nwk",
                                    bl=[0, 0.22, 0.113, 0.15, 0.1], thresholds=[1e-40], methods=["brent"],
                                    sweeps=1),
    "hello_two_trees": dict(fasta="hello.fasta", newick="hello_rooted_two_trees.nwk", thresholds=[1e-40],
                            methods=["brent", "newton"], sweeps=3),
    "five_taxon": dict(fasta="five_taxon.fasta", newick="five_taxon_rooted.nwk", thresholds=[1e-40, 0.5],
                       methods=["brent", "brent_with_gradients", "newton", "logspace_gradient_ascent"],
                       sweeps=3),
    "ds1_reduced_5": dict(fasta="ds1-reduced-5.fasta", newick="ds1-reduced-5.nwk", thresholds=[1e-40],
                          methods=["brent", "newton"], sweeps=3),
    "seven_taxon": dict(fasta="7-taxon-slice-of-ds1.fasta", newick="simplest-hybrid-marginal-all-trees.nwk",
                        thresholds=[1e-40, 0.9], methods=["brent"], sweeps=3),
    "six_taxon": dict(fasta="six_taxon.fasta", newick="six_taxon_rooted_simple.nwk", thresholds=[1e-40],
                      methods=["brent"], sweeps=2),
    "fluA": dict(fasta="fluA.fa", newick="fluA.tree", const_bl=0.01,
                 thresholds=[1e-40, 1e-4, 0.1, 0.5, 0.9], methods=["brent"], sweeps=2),
    "ds1": dict(fasta="ds1/ds1.fasta", newick="ds1/ds1.credible.with-branches.rerooted.nwk",
                thresholds=[1e-40, 0.1, 0.9], methods=["brent"], sweeps=3),
    # BASELINE.json configs[0]: DS1.subsampled_10.t holds UNROOTED trees (trifurcating root), which
    # the rooted GP path rejects (rooted_tree.cpp:151-152). They are re-rooted deterministically,
    # "(A,B,C);" -> "(A,(B,C):0.0);", into ds1_subsampled_10_rerooted.nwk next to this script.
    "ds1_config1": dict(fasta="DS1.fasta", newick="@ds1_subsampled_10_rerooted.nwk", thresholds=[1e-40],
                        methods=["brent"], sweeps=1),
}


def reroot_trifurcation(newick_line: str) -> str:
    """(A,B,C); -> (A,(B,C):0.0); on the top-level trifurcation."""
    s = newick_line.strip()
    assert s.endswith(";")
    body = s[:-1]
    m = re.match(r"^(\[&U\]\s*)?(.*)$", body)
    body = m.group(2)
    assert body[0] == "(" and body.rfind(")") > 0
    close = body.rfind(")")
    inner, tail = body[1:close], body[close + 1:]
    depth, parts, start = 0, [], 0
    for i, ch in enumerate(inner):
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "," and depth == 0:
            parts.append(inner[start:i])
            start = i + 1
    parts.append(inner[start:])
    assert len(parts) == 3, f"expected a trifurcating root, found {len(parts)} children"
    return f"({parts[0]},({parts[1]},{parts[2]}):0.0){tail};"


def make_rerooted_ds1():
    out = os.path.join(HERE, "ds1_subsampled_10_rerooted.nwk")
    with open(DATA + "DS1.subsampled_10.t.nwk") as f:
        lines = [ln for ln in f.read().splitlines() if ln.strip()]
    with open(out, "w") as f:
        for ln in lines:
            f.write(reroot_trifurcation(ln) + "\n")
    return out


def open_case(spec, thr, use_gradients=False):
    newick = spec["newick"]
    newick = os.path.join(HERE, newick[1:]) if newick.startswith("@") else DATA + newick
    e = RefEngine.from_files(DATA + spec["fasta"], newick, thr, use_gradients)
    if "bl" in spec:
        e.set_branch_lengths(spec["bl"])
    if "const_bl" in spec:
        e.set_branch_lengths_to_constant(spec["const_bl"])
    return e


def generate(name, spec):
    out = {}
    e = open_case(spec, spec["thresholds"][0])
    sym, w = e.patterns()
    q, un, inv = e.priors()
    parent, child, on_left = e.edges()
    out.update(symbols=sym, weights=w, site_count=e.site_count, node_count=e.node_count,
               edge_count=e.edge_count, rootsplit_count=e.rootsplit_count, taxon_count=e.taxon_count,
               topology_count=e.topology_count, sbn_prior=q, unconditional_node_probabilities=un,
               inverted_sbn_prior=inv, edge_parent=parent, edge_child=child, edge_on_left=on_left,
               node_bitsets=np.array(e.node_bitsets()), taxon_names=np.array(e.taxon_names()),
               initial_branch_lengths=e.branch_lengths(), thresholds=np.array(spec["thresholds"]),
               methods=np.array(spec["methods"]), sweeps=spec["sweeps"])
    lists = {}
    for ln in LISTS:
        ops, vec = e.oplist(ln)
        lists[ln] = (ops, vec)
        out[f"ops_{ln}"] = ops
        out[f"vec_{ln}"] = vec
    e.close()

    rng = np.random.default_rng(7)
    for ti, thr in enumerate(spec["thresholds"]):
        e = open_case(spec, thr)
        e.process_operations(*lists["populate_plvs"])
        e.process_operations(*lists["compute_likelihoods"])
        key = f"t{ti}"
        out[f"{key}_pass_per_gpcsp_ll"] = e.per_gpcsp_log_likelihoods()
        out[f"{key}_pass_log_marginal"] = e.log_marginal_likelihood()
        out[f"{key}_pass_per_pattern_marginal"] = e.per_pattern_log_marginal()
        out[f"{key}_pass_counts"] = e.rescaling_counts()
        mat = e.log_likelihood_matrix()
        if mat.size <= 40000:
            rows = np.arange(e.edge_count)
        else:
            rows = np.sort(rng.choice(e.edge_count, size=max(8, 40000 // e.pattern_count), replace=False))
        out[f"{key}_pass_ll_rows"] = rows
        out[f"{key}_pass_ll_matrix"] = mat[rows]
        plv_ids = np.sort(rng.choice(e.plv_count, size=min(6, e.plv_count), replace=False))
        out[f"{key}_pass_plv_ids"] = plv_ids
        out[f"{key}_pass_plvs"] = np.stack([e.get_plv(i) for i in plv_ids])
        out[f"{key}_pass_components"] = e.per_gpcsp_components_of_full_log_marginal()
        # derivatives on every 5th non-rootsplit Likelihood op's (edge, parent r-PLV, child p-PLV)
        lik_ops = [r for r in lists["compute_likelihoods"][0] if r[0] == 4][::5][:12]
        trip = np.array([[r[1], r[3], r[2]] for r in lik_ops], dtype=np.int64)  # gpcsp, rootward, leafward
        out[f"{key}_deriv_triples"] = trip
        out[f"{key}_deriv_values"] = np.array(
            [e.log_likelihood_and_derivatives(g, rw, lw, two=True) for g, rw, lw in trip])
        # SBN update straight after the pass (GPInstance::EstimateSBNParameters)
        e.process_operations(*lists["optimize_sbn_parameters"])
        out[f"{key}_sbn_q"] = e.sbn_parameters()
        e.close()

        for method in spec["methods"]:
            e = open_case(spec, thr)
            e.set_optimization_method(method)
            e.reset_optimization_count()
            e.process_operations(*lists["populate_plvs"])
            e.process_operations(*lists["marginal_likelihood"])
            bls, margs, diffs = [], [], []
            for _ in range(spec["sweeps"]):
                e.process_operations(*lists["branch_length_optimization"])
                e.process_operations(*lists["populate_plvs"])
                e.process_operations(*lists["marginal_likelihood"])
                bls.append(e.branch_lengths())
                diffs.append(e.branch_length_differences())
                margs.append(e.log_marginal_likelihood())
                e.increment_optimization_count()
            out[f"{key}_sweep_{method}_bl"] = np.stack(bls)
            out[f"{key}_sweep_{method}_diff"] = np.stack(diffs)
            out[f"{key}_sweep_{method}_log_marginal"] = np.array(margs)
            out[f"{key}_sweep_{method}_counts"] = e.rescaling_counts()
            e.close()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: P={out['symbols'].shape[1]} N={out['node_count']} E={out['edge_count']} "
          f"R={out['rootsplit_count']} -> {os.path.getsize(path) / 1024:.0f} KiB; "
          f"log marginal {out['t0_pass_log_marginal']:.10f}")


def main():
    make_rerooted_ds1()
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        generate(name, spec)


if __name__ == "__main__":
    main()
