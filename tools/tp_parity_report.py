"""Runs oracle/_ref/tp_parity --gpu on the generated cases of tests/test_tp_parity.py and prints its report lines
(the per-check errors pytest hides when a test passes):  python tools/tp_parity_report.py"""
import os
import pathlib
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_host_shim_gpu import _write_case  # noqa: E402
from test_tp_parity import BINARY, CASES  # noqa: E402

for taxa, sites, trees, moves in CASES:
    tmp = pathlib.Path(tempfile.mkdtemp())
    fasta, newick = _write_case(tmp, taxa, sites, trees, moves, seed=taxa * 313 + trees)
    run = subprocess.run([BINARY, fasta, newick, "--gpu"], capture_output=True, text=True, timeout=600)
    print(f"== {taxa} taxa, {sites} sites, {trees} trees (rc {run.returncode})")
    print(run.stdout.strip())
    if run.returncode != 0:
        print(run.stderr[-1000:])
