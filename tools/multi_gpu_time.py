"""Times the pattern-sharded engine on N GPUs (launch with torchrun): full pass, batched sweep and the reference's
Gauss-Seidel sweep, with the scalar all-reduces over NVLink peer memory (default) or NCCL (BITO_GP_PEER_ALLREDUCE=0).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_time.py [workload] [patterns_total]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bito_b200 import _lib  # noqa: E402
from bito_b200.gp_engine import GPEngine  # noqa: E402
from bito_b200.sharding import shard_bounds  # noqa: E402
from bito_b200.synthetic import make_named_workload  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "synthetic-200taxa-100kpat-1000trees"
total = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
wl = make_named_workload(name, pattern_count=total)
dag = wl.dag
lo, hi = shard_bounds(wl.pattern_count, world, rank)
stream = torch.cuda.Stream()  # not torch's default stream: its handle is NULL = "the engine's own stream" to set_stream
torch.cuda.set_stream(stream)
eng = GPEngine(np.ascontiguousarray(wl.symbols[:, lo:hi]), np.ascontiguousarray(wl.weights[lo:hi]), wl.site_count,
               dag.node_count, dag.edge_count, sbn_prior=wl.sbn_prior, unconditional_node_probabilities=wl.unconditional,
               inverted_sbn_prior=wl.inverted, flags=_lib.FLAG_NO_LOGLIK_MATRIX, device=local)
uid = [GPEngine.make_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
eng.comm_init(world, rank, uid[0])
eng.set_stream(stream.cuda_stream)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


pop, lik = wl.ops("populate_plvs"), wl.ops("compute_likelihoods")
t_pass = timed(lambda: (eng.process_operations(*pop), eng.process_operations(*lik)), 5)


def sweep(ops):
    eng.set_branch_lengths_to_constant(0.1)
    eng.reset_optimization_count()
    eng.process_operations(*pop)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    eng.process_operations(*ops)
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


t_batched = min(sweep(wl.ops("batched_branch_length_optimization")) for _ in range(2))
t_gs = min(sweep(wl.ops("branch_length_optimization")) for _ in range(2))
eng.process_operations(*pop)
eng.process_operations(*wl.ops("marginal_likelihood"))
st = eng.stats()
if rank == 0:
    print(f"{world} GPUs, {name}, {wl.pattern_count} patterns total, BITO_GP_PEER_ALLREDUCE="
          f"{os.environ.get('BITO_GP_PEER_ALLREDUCE', 'unset')}: pass {t_pass:.3f} ms, batched sweep {t_batched:.3f} ms, "
          f"Gauss-Seidel sweep {t_gs:.1f} ms, all-reduces {st['collective_calls']} (peer memory {st['peer_collective_calls']}), "
          f"status bits {st['device_status_bits']}, log marginal {eng.get_log_marginal_likelihood():.6f}", flush=True)
eng.close()
dist.destroy_process_group()
