"""Adds quartet hybrid-marginal goldens (tests/golden/quartet_<case>.npz) from the UNMODIFIED
reference (oracle/_ref): the requests GPInstance::CalculateHybridMarginals issues
(/root/reference/src/gp_instance.cpp:408-417, gp_dag.cpp:413-458), every summand of
GPEngine::CalculateQuartetHybridLikelihoods (gp_engine.cpp:748-808), the stored hybrid marginals
(:810-816) and the SBN parameters UpdateSBNProbabilities then derives from them (:304-321).

    python tests/golden/make_golden_quartet.py

Requests whose rootward tip is the DAG root (the parent of the central edge is a rootsplit) index
unconditional_node_probabilities_ and the PLV table one past the last node
(gp_dag.cpp:419-425 with grandparent_id == DAG root; the engine is sized WITHOUT the DAG root,
gp_instance.cpp:158-160) — undefined behaviour in the reference — so they are recorded in
`well_defined` = 0 and left out of the stored outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from make_golden import CASES, open_case  # noqa: E402

QUARTET_CASES = ["hello_two_trees", "five_taxon", "six_taxon", "seven_taxon", "ds1_reduced_5", "ds1"]


def split_requests(central, counts, tips):
    off = 0
    for r in range(central.size):
        n = int(counts[r].sum())
        yield int(central[r]), counts[r], tips[off:off + n]
        off += n


def generate(name):
    spec = CASES[name]
    e = open_case(spec, 1e-40)
    central, counts, tips = e.quartet_requests()
    e.process_operations(*e.oplist("populate_plvs"))
    e.process_operations(*e.oplist("compute_likelihoods"))
    well = np.zeros(central.size, dtype=np.int32)
    liks = []
    for r, (c, n, t) in enumerate(split_requests(central, counts, tips)):
        rootward = t[:n[0]]
        well[r] = int(np.all(rootward[:, 0] < e.node_count))
        if well[r]:
            liks.append(e.calculate_quartet_hybrid_likelihoods(c, n, t))
    # ProcessQuartetHybridRequest on the well-defined requests only
    keep_tips = np.concatenate([t for r, (_, _, t) in enumerate(split_requests(central, counts, tips))
                                if well[r]] or [np.zeros((0, 3), dtype=np.int64)])
    e.process_quartet_hybrid_requests(central[well == 1], counts[well == 1], keep_tips)
    hybrid = e.hybrid_marginals()
    e.process_operations(*e.oplist("optimize_sbn_parameters"))
    q_after = e.sbn_parameters()
    out = dict(central=central, tip_counts=counts, tips=tips, well_defined=well,
               likelihoods=np.concatenate(liks) if liks else np.zeros(0),
               hybrid_marginals=hybrid, sbn_q_after=q_after)
    e.close()
    path = os.path.join(HERE, f"quartet_{name}.npz")
    np.savez_compressed(path, **out)
    formed = int(np.sum(np.prod(counts, axis=1) > 0))
    print(f"{name}: {central.size} requests ({formed} fully formed, {int(well.sum())} well defined), "
          f"{out['likelihoods'].size} summands, finite hybrid marginals {int(np.isfinite(hybrid).sum())}")


if __name__ == "__main__":
    for name in (sys.argv[1:] or QUARTET_CASES):
        generate(name)
