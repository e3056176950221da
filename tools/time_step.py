"""Times the bench step (pass + batched sweep) back to back, and its parts, with per-step CUDA events:
    python tools/time_step.py [workload] [patterns] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bito_b200 import _lib  # noqa: E402
from bito_b200.gp_engine import GPEngine  # noqa: E402
from bito_b200.synthetic import make_named_workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "synthetic-1000taxa-1Mpat-5000trees"
patterns = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "-" else None
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
wl = make_named_workload(name, pattern_count=patterns)
dag = wl.dag
flags = _lib.FLAG_NO_LOGLIK_MATRIX if dag.edge_count * wl.pattern_count * 8 > 16e9 else 0
eng = GPEngine(wl.symbols, wl.weights, wl.site_count, dag.node_count, dag.edge_count, sbn_prior=wl.sbn_prior,
               unconditional_node_probabilities=wl.unconditional, inverted_sbn_prior=wl.inverted, flags=flags)
stream = torch.cuda.Stream()  # not torch's default stream: its handle is NULL = "the engine's own stream" to set_stream
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
pop, lik = wl.ops("populate_plvs"), wl.ops("compute_likelihoods")
blo = wl.ops("batched_branch_length_optimization")


def step():
    eng.set_branch_lengths_to_constant(0.1)
    eng.reset_optimization_count()
    eng.process_operations(*pop)
    eng.process_operations(*lik)
    eng.process_operations(*blo)


def parts():
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    eng.set_branch_lengths_to_constant(0.1)
    eng.reset_optimization_count()
    evs[0].record(stream)
    eng.process_operations(*pop)
    evs[1].record(stream)
    eng.process_operations(*lik)
    evs[2].record(stream)
    eng.process_operations(*blo)
    evs[3].record(stream)
    return evs


for _ in range(3):
    step()
torch.cuda.synchronize()
all_evs = [parts() for _ in range(reps)]
torch.cuda.synchronize()
for evs in all_evs:
    print("populate %.2f  likelihoods %.2f  sweep %.2f  ms" % tuple(evs[i].elapsed_time(evs[i + 1]) for i in range(3)))
print("whole: %.2f ms per step" % (all_evs[0][0].elapsed_time(all_evs[-1][3]) / reps))
print(eng.stats())
