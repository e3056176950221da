"""Prints, for every golden fixture / threshold / optimisation method, the largest deviation of the
CUDA engine from the reference's outputs per sweep (branch lengths, log marginal) without asserting:
the table DESIGN.md quotes.   python tools/parity_report.py [case ...]   (needs a GPU)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gp_cases import ALL_CASES, Fixture, make_cuda, rel_err  # noqa: E402

for case in (sys.argv[1:] or ALL_CASES):
    fx = Fixture(case)
    for ti, thr in enumerate(fx.thresholds):
        for method in fx.methods:
            with make_cuda(fx, ti) as e:
                e.set_optimization_method(method)
                e.reset_optimization_count()
                e.process_operations(*fx.ops("populate_plvs"))
                e.process_operations(*fx.ops("marginal_likelihood"))
                bl_err, ml_err = [], []
                for s in range(int(fx["sweeps"])):
                    e.process_operations(*fx.ops("branch_length_optimization"))
                    e.process_operations(*fx.ops("populate_plvs"))
                    e.process_operations(*fx.ops("marginal_likelihood"))
                    key = f"t{ti}_sweep_{method}"
                    bl_err.append(float(np.max(np.abs(e.branch_lengths() - fx[key + "_bl"][s]))))
                    ml_err.append(rel_err(e.log_marginal_likelihood(), fx[key + "_log_marginal"][s]))
                    e.increment_optimization_count()
                counts_ok = np.array_equal(e.rescaling_counts()[:fx[key + "_counts"].size], fx[key + "_counts"])
            print(f"{case:24s} thr={thr:<7g} {method:26s} |dBL| " + " ".join(f"{x:.1e}" for x in bl_err) +
                  "  rel dlogmarg " + " ".join(f"{x:.1e}" for x in ml_err) + f"  counts_equal={counts_ok}",
                  flush=True)
