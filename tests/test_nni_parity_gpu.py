"""GPU test of the NNI-search caller on top of the engine.

oracle/_ref/nni_parity_ref and oracle/_ref/nni_parity_b200 are ONE program (tests/cpp/nni_parity.cpp) built
twice in the build container by `make -C oracle nniparity`: once against the unmodified reference, once with
bito_b200/host/gp_engine_b200.hpp installed as gp_engine.hpp and the reference's own nni_engine.cpp /
nni_evaluation_engine.cpp recompiled against it (the swap INTEGRATION.md describes). Both run the reference's
NNIEngine with the GP evaluation engine (grow + reindex, spare PLVs/edges, Copy*Data, graft-DAG scoring,
branch lengths written through the DAGBranchHandler reference; nni_evaluation_engine.cpp:51-843) on inputs
written here, and print every scored and accepted NNI per iteration. Same NNIs in the same order, scores to
1e-9 relative where branch lengths are copied, 1e-6 branch lengths / 1e-7 scores where new edges are optimised."""
import os
import subprocess

import numpy as np
import pytest

from test_host_shim_gpu import _write_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "nni_parity_ref")
B200 = os.path.join(ROOT, "oracle", "_ref", "nni_parity_b200")

pytestmark = pytest.mark.gpu


def _run(binary, *args):
    run = subprocess.run([binary, *map(str, args)], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, (binary, run.stdout[-2000:], run.stderr[-2000:])
    return [line.split() for line in run.stdout.splitlines()]


@pytest.mark.parametrize("taxa,sites,trees,moves,iterations,optimize_new_edges", [
    (5, 300, 2, 1, 3, 0),
    (6, 400, 3, 1, 3, 1),
    (8, 800, 4, 2, 4, 0),
    (8, 800, 4, 2, 4, 1),
    (12, 1500, 6, 2, 3, 1),
])
def test_nni_search_over_the_host_class_matches_reference(cuda_engine_lib, tmp_path, taxa, sites, trees, moves,
                                                          iterations, optimize_new_edges):
    for b in (REF, B200):
        if not os.path.exists(b):
            pytest.fail(f"{b} is missing: run `make -C oracle nniparity` in the build container "
                        "(needs /root/reference); the binaries travel with the snapshot")
    fasta, newick = _write_case(tmp_path, taxa, sites, trees, moves, seed=taxa * 977 + trees)
    want = _run(REF, fasta, newick, iterations, optimize_new_edges, 2)
    got = _run(B200, fasta, newick, iterations, optimize_new_edges, 2)
    assert [w[0] for w in want] == [g[0] for g in got]
    n_scored = 0
    for w, g in zip(want, got):
        key = w[0]
        if key in ("dag", "dag_after", "iteration", "accepted"):
            assert w == g, (w, g)                      # sizes, NNI identities, acceptance: exact
        elif key == "scored":
            assert w[1] == g[1], (w, g)                # the same NNI at the same place of the ordered map
            a, b = float(w[2]), float(g[2])
            tol = 1e-7 if optimize_new_edges else 1e-9
            assert abs(a - b) <= tol * max(1.0, abs(a)), (w, g)
            n_scored += 1
        elif key in ("initial_branch_lengths", "branch_lengths"):
            a, b = np.array(w[2:], dtype=float), np.array(g[2:], dtype=float)
            assert w[1] == g[1] and np.max(np.abs(a - b)) <= 1e-6, (key, np.max(np.abs(a - b)))
        elif key in ("initial_per_gpcsp_llh", "final_per_gpcsp_llh", "final_log_marginal"):
            a, b = np.array(w[1:], dtype=float), np.array(g[1:], dtype=float)
            assert np.max(np.abs(a - b) / np.maximum(1.0, np.abs(a))) <= 1e-7, key
        else:
            raise AssertionError(f"unexpected line {w[:2]}")
    assert n_scored > 0
