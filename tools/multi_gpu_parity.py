"""Pattern-sharded parity check, one process per GPU (launch with torchrun):
every rank owns a contiguous slice of a golden fixture's site patterns, the engines join one NCCL
communicator, and the GLOBAL per-edge log-likelihoods, marginal, rescaling counts and optimised
branch lengths must equal the reference's single-process outputs (tests/golden/*.npz).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_parity.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bito_b200 import _lib  # noqa: E402
from bito_b200.gp_engine import GPEngine  # noqa: E402
from bito_b200.sharding import shard_bounds  # noqa: E402
from gp_cases import BL_ATOL, LL_RTOL, Fixture, make_cuda, rel_err  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
failures = []
collectives = [0, 0]  # all-reduces issued, of which over NVLink peer memory (k_peer_allreduce)


def check(name, ok):
    if not ok:
        failures.append(name)
    if rank == 0:
        print(("ok   " if ok else "FAIL ") + name, flush=True)


# hello_single_nucleotide has ONE pattern: every rank but the first owns an empty shard
for case, thresholds in (("ds1", (0, 2)), ("fluA", (0, 4)), ("five_taxon", (0, 1)), ("hello_single_nucleotide", (0,))):
    fx = Fixture(case)
    P = fx["symbols"].shape[1]
    lo, hi = shard_bounds(P, world, rank)
    for ti in thresholds:
        e = make_cuda(fx, ti, pattern_slice=slice(lo, hi), device=local)
        uid = [GPEngine.make_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        e.comm_init(world, rank, uid[0])
        key = f"t{ti}"
        e.process_operations(*fx.ops("populate_plvs"))
        e.process_operations(*fx.ops("compute_likelihoods"))
        check(f"{case} thr#{ti} per-edge ll", rel_err(e.per_gpcsp_log_likelihoods(), fx[f"{key}_pass_per_gpcsp_ll"]) <= LL_RTOL)
        check(f"{case} thr#{ti} marginal", rel_err(e.log_marginal_likelihood(), fx[f"{key}_pass_log_marginal"]) <= LL_RTOL)
        want = fx[f"{key}_pass_counts"]
        check(f"{case} thr#{ti} rescaling counts", np.array_equal(e.rescaling_counts()[:want.size], want))
        rows = fx[f"{key}_pass_ll_rows"]
        check(f"{case} thr#{ti} per-pattern rows (shard)",
              rel_err(e.log_likelihood_matrix()[rows], fx[f"{key}_pass_ll_matrix"][:, lo:hi]) <= LL_RTOL)
        e.process_operations(*fx.ops("optimize_sbn_parameters"))
        check(f"{case} thr#{ti} sbn q", np.max(np.abs(e.sbn_parameters() - fx[f"{key}_sbn_q"])) <= 1e-6)
        e.close()
        # Gauss-Seidel sweeps with Brent (the default method): once with the engine's own choice (clusters that
        # exchange their sums over NVLink inside the kernel), once with the streamed Taylor-model scheme
        # (the batched sweep's path: 16 all-reduced sums per edge and pass)
        for scheme, flags in (("cluster", 0), ("streamed", _lib.FLAG_NO_ONCHIP_OPTIMIZER)):
            e = make_cuda(fx, ti, pattern_slice=slice(lo, hi), device=local, flags=flags)
            dist.broadcast_object_list(uid := [GPEngine.make_unique_id() if rank == 0 else None], src=0)
            e.comm_init(world, rank, uid[0])
            skey = f"{key}_sweep_brent"
            e.set_optimization_method("brent")
            e.reset_optimization_count()
            e.process_operations(*fx.ops("populate_plvs"))
            e.process_operations(*fx.ops("marginal_likelihood"))
            worst = 0.0
            for s in range(int(fx["sweeps"])):
                e.process_operations(*fx.ops("branch_length_optimization"))
                e.process_operations(*fx.ops("populate_plvs"))
                e.process_operations(*fx.ops("marginal_likelihood"))
                worst = max(worst, float(np.max(np.abs(e.branch_lengths() - fx[skey + "_bl"][s]))))
                e.increment_optimization_count()
            check(f"{case} thr#{ti} brent sweeps ({scheme}) |dBL| {worst:.1e}", worst <= BL_ATOL)
            want = fx[skey + "_counts"]
            check(f"{case} thr#{ti} counts after sweeps ({scheme})", np.array_equal(e.rescaling_counts()[:want.size], want))
            st = e.stats()
            collectives[0] += st["collective_calls"]
            collectives[1] += st["peer_collective_calls"]
            check(f"{case} thr#{ti} no device-side assert / peer timeout ({scheme}, optimiser scheme "
                  f"{st['optimizer_scheme']}, {st['objective_passes']} passes for {st['objective_evaluations']} evaluations)",
                  st["device_status_bits"] == 0)
            e.close()

t = torch.tensor([len(failures)], device="cuda")
dist.all_reduce(t)
if rank == 0:
    print(f"all-reduces: {collectives[0]}, of which over peer memory: {collectives[1]} "
          f"(BITO_GP_PEER_ALLREDUCE={os.environ.get('BITO_GP_PEER_ALLREDUCE', 'unset')})", flush=True)
    print("MULTI-GPU PARITY", "PASSED" if t.item() == 0 else f"FAILED ({failures})", f"on {world} GPUs", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 0 else 1)
