"""CPU tests of the N > 1 host logic (gloo, world_size 2): rendezvous helpers, shard bounds, and
the identity the pattern-sharded engine relies on (global per-edge scalars = sum over shards)."""
import os
import subprocess
import sys

import pytest

from bito_b200.sharding import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance():
    for P in (1, 5, 934, 125_000, 1_000_000):
        for G in (1, 2, 3, 4, 8):
            b = [shard_bounds(P, G, r) for r in range(G)]
            assert b[0][0] == 0 and b[-1][1] == P
            assert all(b[i][1] == b[i + 1][0] for i in range(G - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "_dist_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=280)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "DIST-OK 2" in out.stdout
