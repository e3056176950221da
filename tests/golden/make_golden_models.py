"""Generates tests/golden/model_<model>_<case>.npz: the reference GPEngine under GTR and HKY.

    make -C oracle refmodel && python tests/golden/make_golden_models.py

The reference engine hard-wires JC69 (gp_engine.hpp:366) but only talks to its model through the generic
SubstitutionModel interface; oracle/_ref/libbito_gp_ref_model.so is the same unmodified reference with the member's
TYPE swapped (oracle/model_patch.hpp, force-included - no reference file is edited) for a wrapper that builds the
reference's own GTRModel / HKYModel (substitution_model.cpp:79-186) from $BITO_REF_MODEL. Each fixture holds what
make_golden.py's fixtures hold (inputs, op lists, pass and sweep outputs) plus the model's eigensystem exactly as the
reference computed it (GTR: Eigen's SelfAdjointEigenSolver ordering, zero eigenvalue ~1e-16; HKY: the analytic one),
which is what the CUDA engine is given through bito_gp_set_substitution_model.
"""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# name -> BITO_REF_MODEL. gtr: 4 distinct eigenvalues; hky: pi_R = pi_Y makes two of them equal (3 distinct);
# hky4: 4 distinct; gtr_jc: GTR with equal rates and frequencies = JC69 up to rounding (numerically triple eigenvalue)
MODELS = {
    "gtr": "GTR 0.1 0.3 0.1 0.15 0.25 0.1 0.3 0.2 0.2 0.3",
    "hky": "HKY 2.5 0.3 0.2 0.2 0.3",
    "hky4": "HKY 4.0 0.35 0.15 0.2 0.3",
}
CASES = {
    "five_taxon": dict(fasta="five_taxon.fasta", newick="five_taxon_rooted.nwk", thresholds=[1e-40, 0.5],
                       methods=["brent", "brent_with_gradients", "newton"], sweeps=3),
    "ds1_reduced_5": dict(fasta="ds1-reduced-5.fasta", newick="ds1-reduced-5.nwk", thresholds=[1e-40],
                          methods=["brent", "newton"], sweeps=2),
}


def child(model_name):
    import make_golden  # the same protocol as the JC69 fixtures
    from oracle.ref_engine import RefEngine
    for case, spec in CASES.items():
        name = f"model_{model_name}_{case}"
        make_golden.generate(name, spec)
        e = make_golden.open_case(spec, spec["thresholds"][0])
        v, vinv, lam, pi = e.model_eigensystem()
        m = np.stack([e.transition_matrix(t) for t in (0.01, 0.1, 0.75)])
        e.close()
        path = os.path.join(HERE, name + ".npz")
        z = dict(np.load(path))
        z.update(model_spec=np.array(os.environ["BITO_REF_MODEL"]), eigenvectors=v, inverse_eigenvectors=vinv,
                 eigenvalues=lam, frequencies=pi, transition_matrix_times=np.array([0.01, 0.1, 0.75]),
                 transition_matrices=m)
        np.savez_compressed(path, **z)
        print(f"  {name}: eigenvalues {lam}")


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--child":
        sys.path.insert(0, HERE)
        child(sys.argv[2])
    else:
        lib = os.path.join(ROOT, "oracle", "_ref", "libbito_gp_ref_model.so")
        assert os.path.exists(lib), "run `make -C oracle refmodel` first"
        for model_name, spec in MODELS.items():
            env = dict(os.environ, BITO_REF_LIB=lib, BITO_REF_MODEL=spec)
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", model_name], env=env)
