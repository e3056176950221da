"""GPU test of the reference-side host class (bito_b200/host/gp_engine_b200.{hpp,cpp}): the C++
`GPEngine` surface of /root/reference/src/gp_engine.hpp:24-236 over the C-ABI.

oracle/_ref/gp_host_parity (tests/cpp/host_parity.cpp, built in the build container by
`make -C oracle hostparity` because it needs the reference's headers; the binary travels to the
GPU box) parses the files written below with the REFERENCE's parser, plans with the REFERENCE's
GPDAG, and runs the same std::variant GPOperationVectors through the unmodified reference CPU
GPEngine and through GPEngineB200, comparing every public read-back (1e-9 relative log-likelihoods,
1e-6 branch lengths, bit-exact rescaling counts and indices)."""
import os
import subprocess

import numpy as np
import pytest

from bito_b200.synthetic import nni_walk_trees, random_tree, simulate_alignment, tree_to_newick, write_fasta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARY = os.path.join(ROOT, "oracle", "_ref", "gp_host_parity")

pytestmark = pytest.mark.gpu


def _write_case(tmp_path, taxa, sites, trees, moves, seed):
    rng = np.random.default_rng(seed)
    names = [f"taxon{i:03d}" for i in range(taxa)]
    seed_tree = random_tree(taxa, rng, 0.08)
    sample = nni_walk_trees(seed_tree, trees, moves, rng)
    fasta, newick = str(tmp_path / "alignment.fasta"), str(tmp_path / "trees.nwk")
    write_fasta(fasta, simulate_alignment(seed_tree, sites, rng, gap_rate=0.02), names)
    with open(newick, "w") as f:
        for t in sample:
            f.write(tree_to_newick(t, names) + "\n")
    return fasta, newick


@pytest.mark.parametrize("taxa,sites,trees,moves,threshold,sweeps", [
    (6, 400, 4, 1, "1e-40", 2),
    (14, 1500, 12, 2, "1e-40", 2),
    (14, 1500, 12, 2, "0.5", 1),     # rescaling counts become non-zero
    (30, 3000, 40, 2, "1e-40", 1),
])
def test_host_class_matches_reference_engine(cuda_engine_lib, tmp_path, taxa, sites, trees, moves, threshold,
                                             sweeps):
    if not os.path.exists(BINARY):
        pytest.fail(f"{BINARY} is missing: run `make -C oracle hostparity` in the build container "
                    "(needs /root/reference headers); the binary travels with the snapshot")
    fasta, newick = _write_case(tmp_path, taxa, sites, trees, moves, seed=taxa * 1000 + trees)
    run = subprocess.run([BINARY, fasta, newick, threshold, str(sweeps)], capture_output=True, text=True,
                         timeout=600)
    print(run.stdout[-6000:])
    print(run.stderr[-2000:])
    assert run.returncode == 0, run.stdout[-3000:]
    assert "PASS" in run.stdout.splitlines()[-1]
