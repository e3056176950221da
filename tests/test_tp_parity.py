"""The top-pruning (TP) likelihood evaluator as GP op lists.

oracle/_ref/tp_parity (tests/cpp/tp_parity.cpp, built by `make -C oracle tpparity`) holds the reference's unmodified
TPEngine, turns its TPChoiceMap into GPOperationVectors with bito_b200/host/tp_likelihood_plan.hpp and compares the
per-edge top-tree log-likelihoods (TPEngine::GetTopTreeLikelihoods) with the op lists' per-GPCSP log-likelihoods
 * on CPU: through the reference CPU GPEngine (the plan itself; runs in the build container and on the GPU box),
 * on GPU (`--gpu`): through GPEngineB200, i.e. the CUDA kernels, twice with different branch lengths.
Then up to eight NNIs adjacent to the DAG are scored as PROPOSED NNIs (GetTopTreeScoreWithProposedNNI: spare PVs and
edges, lengths from the pre-NNI, with and without five rounds of OptimizeBranchLength on the new edges) by the reference
and by TPLikelihoodPlan::ProposedNNIOps, and the whole-DAG BranchLengthOptimization (two calls of five rounds) by
BranchLengthOptimizationOps. Finally the reference's NNI search runs three iterations in TP mode (top-1 filter, new
edges optimised, nni_search.py:624-642) and the plan, rebuilt for the grown DAG, must reproduce the reference's
re-evaluation of it and the next round of proposed NNIs. BatchedProposedNNIOps scores all adjacent NNIs in one set of
lists (disjoint temps per NNI, so the engine batches the k-th step of every NNI into one level) against the reference
scoring them one at a time (CPU engine, and `--gpu-batched`: CUDA). Last, the search's own INCREMENTAL update
(TPEvalEngineViaLikelihood::UpdateEngineAfterModifyingDAG) is replayed by UpdateAfterModifyingDAGOps on engines that persist
across the DAG's growth (grown with the search's reindexers): scores of the refreshed edges and all branch lengths
must equal what the reference holds after its partial update - bit for bit through the CPU GPEngine. 1e-9 relative (through the CPU GPEngine the plan is bit-exact; CUDA with
optimisation: scores 1e-7, lengths 1e-6); inputs are generated here (the reference's data directory does not travel)."""
import os
import subprocess

import pytest

from test_host_shim_gpu import _write_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARY = os.path.join(ROOT, "oracle", "_ref", "tp_parity")
CASES = [(5, 300, 3, 1), (8, 800, 6, 2), (16, 2000, 12, 2), (30, 3000, 40, 2)]


def _run(tmp_path, taxa, sites, trees, moves, *flags):
    fasta, newick = _write_case(tmp_path, taxa, sites, trees, moves, seed=taxa * 313 + trees)
    run = subprocess.run([BINARY, fasta, newick, *flags], capture_output=True, text=True, timeout=600)
    print(run.stdout[-3000:])
    print(run.stderr[-2000:])
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    lines = run.stdout.splitlines()
    assert lines[-1] == "TP PARITY PASS"
    return lines


@pytest.mark.skipif(not os.path.exists(BINARY), reason="oracle/_ref/tp_parity not built (make -C oracle tpparity)")
@pytest.mark.parametrize("taxa,sites,trees,moves", CASES)
def test_tp_plan_matches_reference_tp_engine_on_cpu(tmp_path, taxa, sites, trees, moves):
    lines = _run(tmp_path, taxa, sites, trees, moves)
    # the per-edge pass; proposed NNIs: scores with fixed lengths, scores and lengths after optimisation; the
    # whole-DAG branch-length optimisation; the DAG grown by three iterations of the reference's TP-mode NNI search
    # (re-evaluated from scratch + its next proposed NNIs)
    # ... every adjacent NNI scored in ONE batch of lists (BatchedProposedNNIOps), and the search's incremental
    # updates replayed on persistent engines (scores, branch lengths)
    assert sum(line.startswith("ok  ") for line in lines) == 10
    assert any("incremental update: plan on CPU GPEngine vs TPEngine, scores" in line and "err 0.000e+00" in line
               for line in lines)  # the replay is the reference's own arithmetic in the reference's own order


@pytest.mark.gpu
@pytest.mark.parametrize("taxa,sites,trees,moves", CASES)
def test_tp_plan_through_the_cuda_engine_matches_reference_tp_engine(cuda_engine_lib, tmp_path, taxa, sites, trees, moves):
    if not os.path.exists(BINARY):
        pytest.fail(f"{BINARY} is missing: run `make -C oracle tpparity` in the build container "
                    "(needs /root/reference); the binary travels with the snapshot")
    lines = _run(tmp_path, taxa, sites, trees, moves, "--gpu")
    # CPU checks as above (5) + CUDA: two per-edge passes, proposed NNIs fixed / optimised / optimised lengths,
    # whole-DAG optimisation, grown DAG + its next proposed NNIs
    # + the incremental updates on a persistent CUDA engine (scores, branch lengths): 10 CPU + 10 CUDA checks
    assert sum(line.startswith("ok  ") for line in lines) == 20


@pytest.mark.gpu
@pytest.mark.parametrize("taxa,sites,trees,moves", CASES[1:3])
def test_tp_batched_proposed_nnis_through_the_cuda_engine(cuda_engine_lib, tmp_path, taxa, sites, trees, moves):
    """All adjacent NNIs in one set of op lists on the CUDA engine (the level scheduler batches the k-th step of
    every NNI into one launch) against the reference scoring them one at a time."""
    if not os.path.exists(BINARY):
        pytest.fail(f"{BINARY} is missing: run `make -C oracle tpparity` in the build container")
    lines = _run(tmp_path, taxa, sites, trees, moves, "--gpu-batched")
    assert any(line.startswith("ok  ") and "batched proposed NNIs (optimised): plan on CUDA" in line for line in lines)
