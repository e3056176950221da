"""CPU tests (no GPU): pin the plain-C oracle (oracle/gp_oracle.c) against the reference's own
golden values and against outputs of the unmodified reference GPEngine (tests/golden/*.npz,
plus oracle/_ref live when it has been built in this container)."""
import math

import numpy as np
import pytest

from oracle import port_engine, ref_engine
from gp_cases import (ALL_CASES, QUARTET_CASES, SMALL_CASES, Fixture, check_pass, check_quartet_hybrid, check_sbn,
                      check_sweeps, make_port, rel_err)


def test_jc69_transition_matrix_golden():
    # /root/reference/src/gp_engine.hpp:382-393
    m = port_engine.transition_matrix(0.75)
    assert abs(0.52590958087 - m[0, 0]) < 1e-10
    assert abs(0.1580301397 - m[0, 1]) < 1e-10
    assert np.allclose(m.sum(axis=1), 1.0, atol=1e-14)


def test_log_add_goldens():
    # /root/reference/src/numerical_utils.hpp:78-113
    assert abs(port_engine.log_add(math.log(2.0), math.log(3.0)) - math.log(5.0)) < 1e-12
    assert port_engine.log_add(-math.inf, -math.inf) == -math.inf
    assert port_engine.log_add(-math.inf, 1.5) == 1.5
    assert port_engine.log_add(0.0, -100.0) == 0.0  # below log(eps): the small term is dropped


def test_hello_classical_likelihood_golden():
    # gp_doctest.cpp:119-131: every per-edge log-likelihood and the marginal equal -84.77961943
    fx = Fixture("hello")
    e = make_port(fx)
    e.process_operations(*fx.ops("populate_plvs"))
    e.process_operations(*fx.ops("compute_likelihoods"))
    assert np.max(np.abs(e.per_gpcsp_log_likelihoods() - -84.77961943)) < 1e-6
    assert abs(e.log_marginal_likelihood() - -84.77961943) < 1e-6


def test_hello_gradient_goldens():
    # gp_doctest.cpp:257-306. PLV ids: P(jupiter)=0*5+0, RLeft(rootsplit)=5*5+4, edge 2.
    fx = Fixture("hello_single_nucleotide")
    e = make_port(fx)
    e.process_operations(*fx.ops("populate_plvs"))
    e.process_operations(*fx.ops("compute_likelihoods"))
    ll, d1 = e.log_likelihood_and_derivatives(2, 29, 0)
    assert abs(ll - -4.806671945) < 1e-6 and abs(d1 - -0.6109379521) < 1e-6
    fx = Fixture("hello")
    e = make_port(fx)
    e.process_operations(*fx.ops("populate_plvs"))
    e.process_operations(*fx.ops("compute_likelihoods"))
    ll, d1, d2 = e.log_likelihood_and_derivatives(2, 29, 0, two=True)
    assert abs(ll - -84.77961943) < 1e-6
    assert abs(d1 - -18.22479569) < 1e-6
    assert abs(d2 - -5.4460787413) < 1e-6


def _estimate_branch_lengths(e, fx, tol, max_iter):
    # GPInstance::EstimateBranchLengths, gp_instance.cpp:241-308
    e.reset_optimization_count()
    e.process_operations(*fx.ops("populate_plvs"))
    e.process_operations(*fx.ops("marginal_likelihood"))
    for _ in range(max_iter):
        e.process_operations(*fx.ops("branch_length_optimization"))
        e.process_operations(*fx.ops("populate_plvs"))
        e.process_operations(*fx.ops("marginal_likelihood"))
        if np.mean(e.branch_length_differences()) < tol:
            break
        e.increment_optimization_count()


def test_hello_newton_optimised_branch_length_golden():
    # gp_doctest.cpp:310-346: PCSP 100|011|001 (rootsplit -> venus = mars|saturn, edge 1) -> 0.0694244266
    fx = Fixture("hello")
    lengths = {}
    for method in ("brent", "newton"):
        e = make_port(fx)
        e.set_optimization_method(method)
        _estimate_branch_lengths(e, fx, 0.0001, 100)
        lengths[method] = e.branch_lengths()[1]
    true_length = 0.0694244266
    assert abs(lengths["newton"] - true_length) < 1e-6
    assert abs(lengths["newton"] - true_length) < abs(lengths["brent"] - true_length)


def test_fluA_rescaling_invariance_golden():
    # gp_doctest.cpp:348-359: marginal with threshold 1e-40 equals the one with 1e-4
    fx = Fixture("fluA")
    vals = []
    for ti in (0, 1):
        e = make_port(fx, ti)
        e.process_operations(*fx.ops("populate_plvs"))
        e.process_operations(*fx.ops("compute_likelihoods"))
        vals.append(e.log_marginal_likelihood())
    assert abs(vals[0] - vals[1]) < 1e-10


@pytest.mark.parametrize("case", ALL_CASES)
def test_port_matches_reference_pass(case):
    fx = Fixture(case)
    for ti in range(len(fx.thresholds)):
        e = make_port(fx, ti)
        check_pass(e, fx, ti, rtol=1e-12)
        check_sbn(e, fx, ti)


@pytest.mark.parametrize("case", ALL_CASES)
def test_port_matches_reference_sweeps(case):
    fx = Fixture(case)
    for ti in range(len(fx.thresholds)):
        for method in fx.methods:
            # Brent-with-gradients steps along -t*dl/dt, which is ill-conditioned for t -> 0
            # (rows of Q sum to 0): rounding noise is amplified to ~1e-7 there.
            atol = 1e-6 if method == "brent_with_gradients" else 1e-8
            check_sweeps(make_port(fx, ti), fx, ti, method, atol=atol)


def test_rescaling_counts_are_exercised():
    # thresholds {0.1, 0.5, 0.9} are the ones that make counts non-zero.
    fx = Fixture("fluA")
    assert fx["t0_pass_counts"].max() == 0 and fx["t1_pass_counts"].max() == 0
    assert fx["t2_pass_counts"].max() >= 1 and fx["t4_pass_counts"].max() >= 20
    assert Fixture("ds1")["t2_pass_counts"].max() >= 30


def test_brent_minimize_on_a_parabola():
    x, fx = port_engine.brent_minimize(lambda v: (v - 0.3) ** 2 + 1.0, 0.0, -2.0, 2.0, 20, 100)
    assert abs(x - 0.3) < 1e-5 and abs(fx - 1.0) < 1e-9


@pytest.mark.skipif(not ref_engine.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("case", ["hello_two_trees", "seven_taxon"])
def test_port_matches_live_reference(case):
    import os
    from golden.make_golden import CASES, DATA, open_case
    if not os.path.exists(DATA):
        pytest.skip("/root/reference absent")
    fx = Fixture(case)
    r = open_case(CASES[case], fx.thresholds[0])
    p = make_port(fx)
    for name in ("populate_plvs", "compute_likelihoods", "branch_length_optimization", "populate_plvs",
                 "compute_likelihoods"):
        r.process_operations(*fx.ops(name))
        p.process_operations(*fx.ops(name))
    assert np.max(np.abs(p.branch_lengths() - r.branch_lengths())) < 1e-9
    assert rel_err(p.log_likelihood_matrix(), r.log_likelihood_matrix()) < 1e-8
    assert np.array_equal(p.rescaling_counts(), r.rescaling_counts())


@pytest.mark.parametrize("case", QUARTET_CASES)
def test_port_quartet_hybrid_marginals_match_reference(case):
    fx = Fixture(case)
    check_quartet_hybrid(make_port(fx), fx)
