"""GTR and HKY on the CUDA engine: bito_gp_set_substitution_model installs the eigensystem the
reference's GTRModel / HKYModel computed; outputs are compared with the reference GPEngine running on those models
(tests/golden/model_*.npz, see tests/test_models.py) - log-likelihoods 1e-9, branch lengths 1e-6, rescaling counts
bit-exact - through every optimiser path that serves models with other than two distinct eigenvalues: one block per
edge, and the round-per-launch scheme with per-eigenvalue coefficients (k_opt_eval<3>, k_opt_eval<4>)."""
import numpy as np
import pytest

from gp_cases import BL_ATOL, LL_RTOL, Fixture, check_pass, check_sbn, check_sweeps, make_cuda, rel_err
from test_gp_engine_gpu import _launches_of, _random_problem
from test_models import MODEL_CASES, sweep_atol

pytestmark = pytest.mark.gpu


def install(engine, fx):
    engine.set_substitution_model(fx["eigenvectors"], fx["inverse_eigenvectors"], fx["eigenvalues"], fx["frequencies"])


@pytest.mark.parametrize("scheme", ["on_chip", "rounds"])
@pytest.mark.parametrize("case", MODEL_CASES)
def test_cuda_matches_reference_under_gtr_and_hky(cuda_engine_lib, case, scheme):
    from bito_b200 import _lib
    flags = 0 if scheme == "on_chip" else _lib.FLAG_NO_ONCHIP_OPTIMIZER
    fx = Fixture(case)
    with make_cuda(fx, 0) as e:
        install(e, fx)
        for t, want in zip(fx["transition_matrix_times"], fx["transition_matrices"]):
            assert np.max(np.abs(e.get_transition_matrix(float(t)) - want)) < 1e-14
    for ti in range(len(fx.thresholds)):
        with make_cuda(fx, ti, flags=flags) as e:
            install(e, fx)
            check_pass(e, fx, ti, rtol=LL_RTOL)
            check_sbn(e, fx, ti)
        for method in fx.methods:
            with make_cuda(fx, ti, flags=flags) as e:
                install(e, fx)
                check_sweeps(e, fx, ti, method, atol=sweep_atol(method))


def test_models_can_be_switched_on_one_engine(cuda_engine_lib):
    """JC69 -> GTR -> JC69 on the same engine: the cached programs (and graphs) stay valid, the results follow the
    model; two engines with different models alternate in one process."""
    jc, gtr = Fixture("five_taxon"), Fixture("model_gtr_five_taxon")
    with make_cuda(jc, 0) as e, make_cuda(jc, 0) as other:
        check_pass(e, jc, 0)
        install(e, gtr)
        check_pass(e, gtr, 0)
        check_pass(other, jc, 0)      # the other engine still computes under JC69
        check_pass(e, gtr, 0)
        from bito_b200.gp_engine import GPEngine  # noqa: F401
        e.set_substitution_model(np.array([[1.0, 2.0, 0.0, 0.5], [1.0, -2.0, 0.5, 0.0], [1.0, 2.0, 0.0, -0.5],
                                           [1.0, -2.0, -0.5, 0.0]]),
                                 np.array([[0.25, 0.25, 0.25, 0.25], [0.125, -0.125, 0.125, -0.125],
                                           [0.0, 1.0, 0.0, -1.0], [1.0, 0.0, -1.0, 0.0]]),
                                 np.array([0.0, -4.0 / 3, -4.0 / 3, -4.0 / 3]), np.full(4, 0.25))
        check_pass(e, jc, 0)
        with pytest.raises(RuntimeError, match="frequencies"):
            e.set_substitution_model(np.eye(4), np.eye(4), np.zeros(4), np.array([0.5, 0.5, 0.5, 0.5]))


@pytest.mark.parametrize("model", ["model_gtr_five_taxon", "model_hky_five_taxon"])
@pytest.mark.parametrize("taxa,patterns,thr", [(9, 257, 0.5), (30, 3001, 1e-40), (64, 20000, 1e-40)])
def test_random_trees_under_gtr_and_hky_match_oracle(cuda_engine_lib, model, taxa, patterns, thr):
    """Ragged sizes and a batched optimisation of every edge at once, CUDA vs the plain-C oracle with the same
    eigensystem; at 20 000 patterns the optimiser streams per-eigenvalue coefficients (k_opt_eval<G>)."""
    from bito_b200.gp_engine import GPEngine
    from oracle import port_engine
    from oracle.port_engine import PortEngine
    fx = Fixture(model)
    rng = np.random.default_rng(taxa * 1000 + patterns)
    pb = _random_problem(rng, taxa, patterns)
    site_count = int(pb["weights"].sum())
    port_engine.set_model(fx["eigenvectors"], fx["inverse_eigenvectors"], fx["eigenvalues"], fx["frequencies"])
    try:
        cpu = PortEngine(pb["symbols"], pb["weights"], site_count, pb["node_count"], pb["edge_count"],
                         rescaling_threshold=thr)
        with GPEngine(pb["symbols"], pb["weights"], site_count, pb["node_count"], pb["edge_count"], thr) as gpu:
            install(gpu, fx)
            gpu.set_profiling(True)
            for e in (cpu, gpu):
                e.set_branch_lengths(pb["branch_lengths"])
                e.process_operations(*pb["populate"])
                e.process_operations(*pb["likelihoods"])
            assert rel_err(gpu.get_log_likelihood_matrix(), cpu.log_likelihood_matrix()) <= LL_RTOL
            assert rel_err(gpu.get_log_marginal_likelihood(), cpu.log_marginal_likelihood()) <= LL_RTOL
            assert np.array_equal(gpu.get_rescaling_counts(), cpu.rescaling_counts())
            for method in ("brent", "newton"):
                for e in (cpu, gpu):
                    e.set_branch_lengths(pb["branch_lengths"])
                    e.set_optimization_method(method)
                    e.reset_optimization_count()
                    e.process_operations(*pb["populate"])
                    e.process_operations(*pb["optimize"])
                assert np.max(np.abs(gpu.get_branch_lengths() - cpu.branch_lengths())) <= BL_ATOL, method
            if patterns >= 20000:
                assert _launches_of(gpu, "k_opt_eval") > 0 and _launches_of(gpu, "k_opt_cluster") == 0
        cpu.close()
    finally:
        port_engine.set_model()
