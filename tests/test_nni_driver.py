"""CPU check of the NNI parity driver (tests/cpp/nni_parity.cpp) in its all-reference build: the unmodified
reference NNIEngine + CPU GPEngine score and accept NNIs on a generated case, so the GPU test
(test_nni_parity_gpu.py) compares the swapped build against a run that is known to do real work."""
import os
import subprocess

import pytest

from test_host_shim_gpu import _write_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "nni_parity_ref")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/nni_parity_ref not built (make -C oracle nniparity)")
def test_reference_build_of_the_nni_driver_scores_and_accepts_nnis(tmp_path):
    fasta, newick = _write_case(tmp_path, 7, 500, 3, 1, seed=7 * 977 + 3)
    run = subprocess.run([REF, fasta, newick, "3", "1", "2"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr[-2000:]
    keys = [line.split()[0] for line in run.stdout.splitlines()]
    assert keys[0] == "dag" and keys[-1] == "final_log_marginal"
    assert keys.count("iteration") == 3 and keys.count("accepted") == 3      # top-1 filter: one NNI per iteration
    assert keys.count("scored") >= 3 * 4
    sizes = [tuple(map(int, (l.split()[2], l.split()[4]))) for l in run.stdout.splitlines() if l.startswith("dag_after")]
    assert sizes == sorted(sizes) and sizes[0] < sizes[-1]                   # the DAG (and the engine) grew
