"""Worker of tests/test_distributed.py: run under torchrun with the gloo backend (CPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch.distributed as dist  # noqa: E402

from bito_b200 import distributed as D  # noqa: E402
from bito_b200.sharding import shard_bounds  # noqa: E402
from gp_cases import Fixture  # noqa: E402
from oracle.port_engine import PortEngine  # noqa: E402  (the checker; stands in for the local engine)

rank, world, local = D.init("gloo")
assert world == int(os.environ["WORLD_SIZE"]) and D.world_size() == world

# 1. the 128-byte communicator id reaches every rank unchanged
payload = bytes(range(128)) if rank == 0 else None
got = D.broadcast_bytes(payload, 128, src=0)
assert got == bytes(range(128)), "broadcast_bytes"

# 2. timing reduction is a max, scalar reduction a sum
assert D.max_over_ranks(10.0 + rank) == 10.0 + world - 1
assert np.allclose(D.sum_over_ranks([1.0, rank]), [world, world * (world - 1) / 2])

# 2b. rank agreement as bench.py checks it: MAX of v and of -v agree exactly iff every rank holds the same values
same = np.array([0.125, 3.0, -7.5])
assert np.array_equal(D.max_over_ranks_array(same), -D.max_over_ranks_array(-same))
differ = np.array([0.125, 3.0 + rank, -7.5])
assert not np.array_equal(D.max_over_ranks_array(differ), -D.max_over_ranks_array(-differ))

# 3. shards tile [0, P) exactly, ragged P included
for P in (1, 2, 7, 934, 100_000, 1_000_003):
    lo, hi = D.my_shard(P)
    assert (lo, hi) == shard_bounds(P, world, rank)
    sizes = D.sum_over_ranks([hi - lo])
    assert int(sizes[0]) == P
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi))
    assert gathered[0][0] == 0 and gathered[-1][1] == P
    assert all(gathered[i][1] == gathered[i + 1][0] for i in range(world - 1))
    assert max(b - a for a, b in gathered) - min(b - a for a, b in gathered) <= 1

# 4. the sharded likelihood pass is the unsharded one: per-edge sums and the marginal are sums over
#    shards, per-pattern rows are slices, and the per-PLV rescale decision needs the MAX over shards
#    (thresholds 1e-40: none fires, every count is 0 on every shard).
fx = Fixture("fluA")
a = fx.engine_args(0)
lo, hi = D.my_shard(a["symbols"].shape[1])
e = PortEngine(a["symbols"][:, lo:hi], a["weights"][lo:hi], a["site_count"], a["node_count"], a["edge_count"],
               a["q"], a["unconditional"], a["inverted"], a["rescaling_threshold"])
e.set_branch_lengths(fx["initial_branch_lengths"])
e.process_operations(*fx.ops("populate_plvs"))
e.process_operations(*fx.ops("compute_likelihoods"))
per_edge = D.sum_over_ranks(e.per_gpcsp_log_likelihoods())
marginal = D.sum_over_ranks([e.log_marginal_likelihood()])[0]
assert np.max(np.abs(per_edge - fx["t0_pass_per_gpcsp_ll"]) / np.abs(fx["t0_pass_per_gpcsp_ll"])) < 1e-12
assert abs(marginal - fx["t0_pass_log_marginal"]) < 1e-9 * abs(fx["t0_pass_log_marginal"])
rows = fx["t0_pass_ll_rows"]
assert np.allclose(e.log_likelihood_matrix()[rows], fx["t0_pass_ll_matrix"][:, lo:hi], rtol=1e-12, atol=1e-12)
assert not e.rescaling_counts().any()

D.barrier()
if rank == 0:
    print("DIST-OK", world)
dist.destroy_process_group()
