"""The reference's NNI search in TP mode (test/nni_search.py --tp: NNIEngine + NNIEvalEngineViaTP + TPEngine, all
unmodified objects) with the TPEngine's likelihood evaluator swapped for TPEvalEngineOverGPEngine
(bito_b200/host/tp_eval_engine_b200.hpp), a subclass of the reference's TPEvalEngineViaLikelihood that serves its
virtual interface with GP op lists (tp_likelihood_plan.hpp) on a GP engine.

oracle/_ref/tp_search_parity (tests/cpp/tp_search_parity.cpp, `make -C oracle tpsearchparity`) runs three search
iterations (top-1 filter, new edges optimised) three ways on DAGs built from the same trees:
 * the reference TPEngine as it is;
 * the swapped evaluator over the reference CPU GPEngine - no GPU needed: identical scored and accepted NNIs, and
   scores, top-tree log-likelihoods and branch lengths equal BIT FOR BIT after every iteration;
 * `--gpu`: the swapped evaluator over GPEngineB200 (the CUDA kernels): identical scored / accepted NNIs, scores and
   top-tree log-likelihoods to 1e-7 relative, branch lengths to 1e-6."""
import os
import subprocess

import pytest

from test_host_shim_gpu import _write_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARY = os.path.join(ROOT, "oracle", "_ref", "tp_search_parity")
CASES = [(5, 300, 3, 1), (8, 800, 6, 2), (16, 2000, 12, 2), (30, 3000, 40, 2)]


def _run(tmp_path, taxa, sites, trees, moves, *flags):
    fasta, newick = _write_case(tmp_path, taxa, sites, trees, moves, seed=taxa * 313 + trees)
    run = subprocess.run([BINARY, fasta, newick, *flags], capture_output=True, text=True, timeout=900)
    print(run.stdout[-3000:])
    print(run.stderr[-2000:])
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    lines = run.stdout.splitlines()
    assert lines[-1] == "TP SEARCH PARITY PASS"
    return lines


@pytest.mark.skipif(not os.path.exists(BINARY), reason="oracle/_ref/tp_search_parity not built (make -C oracle tpsearchparity)")
@pytest.mark.parametrize("taxa,sites,trees,moves", CASES)
def test_tp_mode_search_with_swapped_evaluator_on_cpu_engine(tmp_path, taxa, sites, trees, moves):
    lines = _run(tmp_path, taxa, sites, trees, moves)
    ok = [line for line in lines if line.startswith("ok  ")]
    assert len(ok) == 1 and "same NNIs 1" in ok[0]
    # bit for bit: the class issues the reference's own arithmetic in the reference's own order
    assert "scores 0.000e+00" in ok[0] and "top-tree llh 0.000e+00" in ok[0] and "|dBL| 0.000e+00" in ok[0]


@pytest.mark.gpu
@pytest.mark.parametrize("taxa,sites,trees,moves", CASES)
def test_tp_mode_search_with_swapped_evaluator_on_cuda_engine(cuda_engine_lib, tmp_path, taxa, sites, trees, moves):
    if not os.path.exists(BINARY):
        pytest.fail(f"{BINARY} is missing: run `make -C oracle tpsearchparity` in the build container "
                    "(needs /root/reference); the binary travels with the snapshot")
    lines = _run(tmp_path, taxa, sites, trees, moves, "--gpu")
    ok = [line for line in lines if line.startswith("ok  ")]
    assert len(ok) == 2 and all("same NNIs 1" in line for line in ok)
