"""TEST INFRASTRUCTURE ONLY. How rounding-stable are the reference's own optimised branch lengths?

Builds the UNMODIFIED reference GP path a second time with different floating-point code
generation (`make -C oracle refvar`: -O2 -march=native -ffp-contract=fast instead of -O3) and
replays the protocol of tests/golden/make_golden.py against the committed goldens. Prints the
largest |branch length - golden| per sweep. Needs /root/reference (build container only).

    make -C oracle -j8 refvar && python oracle/ref_rounding_sensitivity.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_engine  # noqa: E402

ref_engine.LIB_PATH = os.path.join(HERE, "_ref", "variant", "libbito_gp_ref.so")
from golden.make_golden import CASES, open_case  # noqa: E402
from gp_cases import Fixture  # noqa: E402

for name, spec in CASES.items():
    fx = Fixture(name)
    for ti, thr in enumerate(spec["thresholds"]):
        for method in spec["methods"]:
            e = open_case(spec, thr)
            e.set_optimization_method(method)
            e.reset_optimization_count()
            e.process_operations(*fx.ops("populate_plvs"))
            e.process_operations(*fx.ops("marginal_likelihood"))
            errs = []
            for s in range(spec["sweeps"]):
                e.process_operations(*fx.ops("branch_length_optimization"))
                e.process_operations(*fx.ops("populate_plvs"))
                e.process_operations(*fx.ops("marginal_likelihood"))
                errs.append(float(np.max(np.abs(e.branch_lengths() - fx[f"t{ti}_sweep_{method}_bl"][s]))))
                e.increment_optimization_count()
            e.close()
            print(f"{name:24s} thr={thr:<7g} {method:26s} " + " ".join(f"{x:.2e}" for x in errs), flush=True)
