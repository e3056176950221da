"""GTR and HKY through the GP path, CPU part: the plain-C oracle, given the eigensystem the
reference's own GTRModel / HKYModel computed (substitution_model.cpp:79-186), against outputs of the reference
GPEngine running on those models (tests/golden/model_*.npz, written by tests/golden/make_golden_models.py from
oracle/_ref/libbito_gp_ref_model.so: the unmodified reference with the engine's hard-wired JC69Model member type
swapped by oracle/model_patch.hpp). This pins the oracle for models with 3 and 4 distinct eigenvalues."""
import numpy as np
import pytest

from gp_cases import BL_ATOL, LL_RTOL, Fixture, check_pass, check_sbn, check_sweeps, make_port

MODEL_CASES = ["model_gtr_five_taxon", "model_gtr_ds1_reduced_5", "model_hky_five_taxon", "model_hky_ds1_reduced_5",
               "model_hky4_five_taxon", "model_hky4_ds1_reduced_5"]


def sweep_atol(method):
    # BrentOptimizationWithGradients steps by 1.0005 * t * dl/dt (optimization.hpp:286-289): the rounding noise of
    # the derivative lands in the branch length; with general eigenvectors the C port's and
    # Eigen's 4x4 products round differently, which this method amplifies to ~1e-5 on the 4-pattern fixture
    return 5e-5 if method == "brent_with_gradients" else BL_ATOL


@pytest.fixture
def port_model():
    from oracle import port_engine

    def install(fx):
        port_engine.set_model(fx["eigenvectors"], fx["inverse_eigenvectors"], fx["eigenvalues"], fx["frequencies"])
    yield install
    port_engine.set_model()  # back to JC69 for every other test of the process


def test_fixtures_hold_three_and_four_distinct_eigenvalues():
    distinct = {name: len(set(Fixture(name)["eigenvalues"].tolist())) for name in MODEL_CASES}
    assert distinct["model_gtr_five_taxon"] == 4 and distinct["model_hky4_five_taxon"] == 4
    assert distinct["model_hky_five_taxon"] == 3
    for name in MODEL_CASES:
        fx = Fixture(name)
        v, vinv = fx["eigenvectors"], fx["inverse_eigenvectors"]
        assert np.max(np.abs(v @ vinv - np.eye(4))) < 1e-14
        assert abs(fx["frequencies"].sum() - 1.0) < 1e-12


@pytest.mark.parametrize("case", MODEL_CASES)
def test_oracle_matches_reference_under_gtr_and_hky(case, port_model):
    from oracle import port_engine
    fx = Fixture(case)
    port_model(fx)
    for t, want in zip(fx["transition_matrix_times"], fx["transition_matrices"]):
        assert np.max(np.abs(port_engine.transition_matrix(float(t)) - want)) < 1e-14
    for ti in range(len(fx.thresholds)):
        e = make_port(fx, ti)
        check_pass(e, fx, ti, rtol=LL_RTOL)
        check_sbn(e, fx, ti)
        e.close()
        for method in fx.methods:
            e = make_port(fx, ti)
            check_sweeps(e, fx, ti, method, atol=sweep_atol(method))
            e.close()
