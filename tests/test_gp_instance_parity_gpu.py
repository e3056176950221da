"""The reference's own GPInstance (gp_instance.cpp, what bito's users and pybito.cpp drive) on top of the host class.

oracle/_ref/gp_instance_parity_ref and oracle/_ref/gp_instance_parity_b200 are ONE program
(tests/cpp/gp_instance_parity.cpp) built twice by `make -C oracle gpinstanceparity`: the reference's gp_instance.cpp
with its CPU GPEngine, and the SAME gp_instance.cpp, unchanged, compiled against bito_b200/host/gp_engine_b200.hpp
installed as gp_engine.hpp (in both builds fat_beagle.hpp is a stub: BEAGLE is not on the GP path). The program walks
BASELINE.json configs[0..2] through GPInstance: PopulatePLVs + ComputeLikelihoods, TakeFirst / HotStart branch lengths,
EstimateBranchLengths + ComputeMarginalLikelihood, EstimateSBNParameters, CalculateHybridMarginals, the CSV exporters.
Log-likelihoods 1e-9 relative at fixed branch lengths (1e-7 after optimisation), branch lengths 1e-6, SBN parameters 1e-6."""
import os
import subprocess

import numpy as np
import pytest

from test_host_shim_gpu import _write_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "gp_instance_parity_ref")
B200 = os.path.join(ROOT, "oracle", "_ref", "gp_instance_parity_b200")
KEYS = ["dag", "pass_per_gpcsp_llh", "pass_log_marginal", "take_first_branch_lengths", "hot_start_branch_lengths",
        "estimated_branch_lengths", "estimated_per_gpcsp_llh", "estimated_log_marginal", "sbn_parameters",
        "hybrid_marginals", "sbn_parameters_after_hybrid", "exporters"]


def _run(binary, *args):
    run = subprocess.run([binary, *map(str, args)], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, (binary, run.stdout[-2000:], run.stderr[-2000:])
    lines = {}
    for line in run.stdout.splitlines():
        parts = line.split()
        if parts and (parts[0] in KEYS or parts[0] == "flatness"):
            lines[parts[0]] = parts[1:]
    assert [k for k in lines if k != "flatness"] == KEYS, list(lines)
    return lines


def _values(fields, counted=True):
    return np.array(fields[1:] if counted else fields, dtype=float)


def _close(a, b, rtol=0.0, atol=0.0):
    # the reference's hybrid marginals include its spare edges (padded count); compare what both hold
    n = min(a.size, b.size)
    a, b = a[:n], b[:n]
    # -inf (an edge without hybrid marginal) and NaN (the reference's own softmax over such edges) must sit at
    # the same places
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    same_nan = np.isnan(a) & np.isnan(b)
    finite = np.isfinite(a) & np.isfinite(b)
    assert np.all(same_inf | same_nan | finite), "non-finite entries differ"
    return np.all(np.abs(a[finite] - b[finite]) <= atol + rtol * np.maximum(1.0, np.abs(b[finite])))


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/gp_instance_parity_ref not built")
def test_reference_build_walks_gp_instance(tmp_path):
    fasta, newick = _write_case(tmp_path, 7, 500, 4, 1, seed=7 * 131 + 4)
    out = _run(REF, fasta, newick)
    assert np.all(np.isfinite(_values(out["estimated_per_gpcsp_llh"])))
    assert float(out["estimated_log_marginal"][0]) > float(out["pass_log_marginal"][0])  # optimisation helped


@pytest.mark.gpu
@pytest.mark.parametrize("taxa,sites,trees,moves,threshold", [
    (6, 400, 3, 1, "1e-40"),
    (10, 1000, 8, 2, "1e-40"),
    (10, 1000, 8, 2, "0.5"),      # rescaling fires
    (24, 2500, 30, 2, "1e-40"),
])
def test_gp_instance_over_the_host_class_matches_reference(cuda_engine_lib, tmp_path, taxa, sites, trees, moves,
                                                           threshold):
    for b in (REF, B200):
        if not os.path.exists(b):
            pytest.fail(f"{b} is missing: run `make -C oracle gpinstanceparity` in the build container "
                        "(needs /root/reference); the binaries travel with the snapshot")
    fasta, newick = _write_case(tmp_path, taxa, sites, trees, moves, seed=taxa * 131 + trees)
    # one sweep of EstimateBranchLengths: the north-star tolerances
    want, got = _run(REF, fasta, newick, threshold, 1), _run(B200, fasta, newick, threshold, 1)
    assert want["dag"] == got["dag"]
    assert _close(_values(got["pass_per_gpcsp_llh"]), _values(want["pass_per_gpcsp_llh"]), rtol=1e-9)
    assert _close(_values(got["pass_log_marginal"], False), _values(want["pass_log_marginal"], False), rtol=1e-9)
    for key in ("take_first_branch_lengths", "hot_start_branch_lengths"):  # host-side arithmetic on both sides
        assert _close(_values(got[key]), _values(want[key]), atol=1e-12), key
    assert _close(_values(got["estimated_branch_lengths"]), _values(want["estimated_branch_lengths"]), atol=1e-6)
    assert _close(_values(got["estimated_per_gpcsp_llh"]), _values(want["estimated_per_gpcsp_llh"]), rtol=1e-7)
    assert _close(_values(got["estimated_log_marginal"], False), _values(want["estimated_log_marginal"], False),
                  rtol=1e-7)
    assert _close(_values(got["sbn_parameters"]), _values(want["sbn_parameters"]), atol=1e-6)
    assert _close(_values(got["hybrid_marginals"]), _values(want["hybrid_marginals"]), rtol=1e-7)
    assert _close(_values(got["sbn_parameters_after_hybrid"]), _values(want["sbn_parameters_after_hybrid"]), atol=1e-6)
    # Three sweeps from the hot start: near the optimum Brent compares objective values that differ by less than
    # their rounding noise, so a decision can flip and move a branch inside Brent's own tolerance. Two builds of
    # the UNMODIFIED reference (-O3 vs -O2 -march=native, `make -C oracle refvar`) disagree that way on the second
    # case here: 2 of 54 edges by up to 6.9e-6, log-likelihoods by 1.9e-8 relative. So: >= 90 % of the edges within
    # 1e-6, the rest within Brent's tolerance (2^-9 relative in log t, optimization.hpp:84-100), and the
    # optimised log marginal to 1e-6 relative.
    want, got = _run(REF, fasta, newick, threshold, 3), _run(B200, fasta, newick, threshold, 3)
    w, g = _values(want["estimated_branch_lengths"]), _values(got["estimated_branch_lengths"])
    off = np.abs(w - g) > 1e-6
    assert off.sum() <= 0.1 * w.size, int(off.sum())
    tol = 2.0 ** -9
    assert np.all(np.abs(np.log(g[off]) - np.log(w[off])) <= 4 * (tol * np.abs(np.log(w[off])) + tol / 4))
    # ... and every such edge must be one whose objective cannot tell the two lengths apart: the CUDA build, with its
    # own optimised PLVs in place, scores the edge at its own length and at the reference's (Likelihood(e) depends on
    # t_e only through M(t_e)); the two per-edge log-likelihoods agree to 1e-8 relative (round-1 verdict, item 7).
    # Not 1e-9: after three COUPLED Gauss-Seidel sweeps an early flip has moved the neighbours' PLVs, so the two
    # optimisers stop at different points inside Brent's x-tolerance of slightly different objectives; the
    # difference is second order in that tolerance (worst measured on B200: 1.3e-9 relative, 1.3e-5 absolute on
    # -9634; the two reference builds differ by 1.9e-8 relative in the same place).
    lengths = tmp_path / "reference_lengths.txt"
    lengths.write_text(" ".join(repr(float(x)) for x in w))
    flat = _run(B200, fasta, newick, threshold, 3, lengths)["flatness"]
    n_off = int(flat[0])
    rec = np.array(flat[1:], dtype=float).reshape(n_off, 3)
    assert sorted(rec[:, 0].astype(int).tolist()) == np.nonzero(off)[0].tolist()
    rel = np.abs(rec[:, 1] - rec[:, 2]) / np.maximum(1.0, np.abs(rec[:, 1]))
    assert np.all(rel <= 1e-8), (rec[np.argmax(rel)].tolist(), float(rel.max()) if n_off else 0.0)
    assert _close(_values(got["estimated_log_marginal"], False), _values(want["estimated_log_marginal"], False),
                  rtol=1e-6)
