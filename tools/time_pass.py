"""Times the bench workload's likelihood pass (device-resident, CUDA graphs on) and prints one line:
    python tools/time_pass.py [workload] [patterns] [reps]
Tuning knobs are read from the environment by the library (BITO_GP_TILES_PER_BLOCK, BITO_GP_NODE_OCC)."""
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
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
wl = make_named_workload(name, pattern_count=patterns)
dag = wl.dag
flags = _lib.FLAG_NO_LOGLIK_MATRIX if dag.edge_count * wl.pattern_count * 8 > 16e9 else 0
eng = GPEngine(wl.symbols, wl.weights, wl.site_count, dag.node_count, dag.edge_count, sbn_prior=wl.sbn_prior,
               unconditional_node_probabilities=wl.unconditional, inverted_sbn_prior=wl.inverted, flags=flags)
stream = torch.cuda.Stream()  # not torch's default stream: its handle is NULL = "the engine's own stream" to set_stream
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
pop, lik = wl.ops("populate_plvs"), wl.ops("compute_likelihoods")


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(n):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_pop = timed(lambda: eng.process_operations(*pop), reps)
t_lik = timed(lambda: eng.process_operations(*lik), reps)
out = f"populate {t_pop:.3f} ms  likelihoods {t_lik:.3f} ms  pass {t_pop + t_lik:.3f} ms"
if "sweep" in sys.argv:
    blo = wl.ops("batched_branch_length_optimization")
    eng.reset_optimization_count()
    eng.process_operations(*pop)
    f0 = eng.stats()["objective_evaluations"]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    eng.process_operations(*blo)
    e1.record(stream)
    torch.cuda.synchronize()
    out += f"  sweep {e0.elapsed_time(e1):.3f} ms ({eng.stats()['objective_evaluations'] - f0} evals)"
    eng.process_operations(*pop)
    eng.process_operations(*lik)
print(out + f"  log marginal {eng.get_log_marginal_likelihood():.6f}  env " +
      " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("BITO_GP_")), flush=True)
