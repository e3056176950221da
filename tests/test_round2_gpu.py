"""GPU parity tests added in round 2 (all through the C-ABI, checker = the plain-C oracle port or the
reference-generated goldens):
  * the pipelined cluster optimiser (scheme 3: streaming producer + one cluster per edge reading rho),
    every fixture and shape, multi-chunk rings, weight classes;
  * non-power-of-two cluster shapes of the cluster-resident optimiser;
  * the headline bench DAG (BASELINE.json configs[4]: 1000 taxa, 5657 nodes, 11 144 edges) on a pattern
    subsample against the oracle: pass, rescaling counts at the default 1e-40 threshold (they do fire on
    this deep DAG) and at 0.5, and the batched sweep under the rounds and pipelined schemes;
  * accumulate groups of more than kItemChunk (64) increments: k_node's multi-chunk branch;
  * regressions for the advisor's findings (op order in the zero-elision decision, re-upload with
    graphs on, program-cache eviction)."""
import numpy as np
import pytest

from gp_cases import ALL_CASES, BL_ATOL, LL_RTOL, Fixture, check_sweeps, make_cuda, rel_err
from test_gp_engine_gpu import _launches_of, _random_problem, _random_tree_case

pytestmark = pytest.mark.gpu


@pytest.fixture
def pipelined(request, monkeypatch):
    """BITO_GP_OPT_SCHEME=3 forces the pipelined cluster scheme wherever plain Brent runs; "C" or "CxT"
    fixes the cluster shape; BITO_GP_OPT_RING_EDGES=R makes chunks of R edges (several chunks, both ring
    halves reused) even on small fixtures."""
    shape, _, ring = str(request.param).partition("/")
    c, _, t = shape.partition("x")
    monkeypatch.setenv("BITO_GP_OPT_SCHEME", "3")
    monkeypatch.setenv("BITO_GP_OPT_CLUSTER", c)
    if t:
        monkeypatch.setenv("BITO_GP_OPT_CLUSTER_THREADS", t)
    if ring:
        monkeypatch.setenv("BITO_GP_OPT_RING_EDGES", ring)
    return request.param


@pytest.mark.parametrize("pipelined", ["1", "3", "16", "5x512", "12x1024"], indirect=True)
@pytest.mark.parametrize("case", ALL_CASES)
def test_pipelined_optimizer_sweeps_match_reference(cuda_engine_lib, case, pipelined):
    fx = Fixture(case)
    if "brent" not in fx.methods:
        pytest.skip("fixture has no plain-Brent sweep")
    for ti in range(len(fx.thresholds)):
        with make_cuda(fx, ti) as e:
            check_sweeps(e, fx, ti, "brent")
            st = e.stats()
        assert st["objective_evaluations"] > 0 and st["graph_launches"] > 0 and st["optimizer_scheme"] == 3
    with make_cuda(fx, 0) as e:  # the kernels that ran really were the producer and the cluster consumer
        e.set_optimization_method("brent")
        e.set_profiling(True)
        e.process_operations(*fx.ops("populate_plvs"))
        e.process_operations(*fx.ops("branch_length_optimization"))
        assert _launches_of(e, "k_opt_cluster") > 0 and _launches_of(e, "k_opt_prepare") > 0
        assert _launches_of(e, "k_opt_block") == 0 and _launches_of(e, "k_opt_eval") == 0


@pytest.mark.parametrize("pipelined", ["2/3", "16/7", "4x1024/5", "7x512/1"], indirect=True)
@pytest.mark.parametrize("taxa,patterns,thr", [(4, 1, 1e-40), (7, 255, 1e-40), (9, 256, 0.5), (12, 257, 0.9),
                                                (30, 3001, 0.7), (64, 20000, 1e-40)])
def test_pipelined_optimizer_random_trees_match_oracle(cuda_engine_lib, taxa, patterns, thr, pipelined):
    """A batched optimisation of every edge at once in chunks of a few edges: the ring halves are handed
    back and forth between the producer and the consumer stream several times per level."""
    from oracle.port_engine import PortEngine
    rng = np.random.default_rng(taxa * 1000 + patterns)
    pb = _random_problem(rng, taxa, patterns)
    site_count = int(pb["weights"].sum())
    cpu = PortEngine(pb["symbols"], pb["weights"], site_count, pb["node_count"], pb["edge_count"],
                     rescaling_threshold=thr)
    _random_tree_case(pb, cpu, site_count, thr, 0)


@pytest.mark.parametrize("pipelined", ["1/4", "16/4", "3x512/2"], indirect=True)
def test_pipelined_optimizer_weight_classes(cuda_engine_lib, pipelined):
    """Every weight class of the cluster layout through the producer's scatter (pattern -> position), then
    other weights on the same engine: the ring's padding positions must be re-zeroed for the new layout."""
    from bito_b200.gp_engine import GPEngine
    from oracle.port_engine import PortEngine
    rng = np.random.default_rng(12)
    pb = _random_problem(rng, 10, 2500)
    w = pb["weights"].copy()
    w[::5] = rng.integers(2, 8, size=w[::5].size)
    w[3::7] = rng.uniform(0.25, 3.5, size=w[3::7].size)
    w[5::31] = rng.integers(8, 400, size=w[5::31].size)
    site_count = int(round(w.sum()))
    cpu = PortEngine(pb["symbols"], w, site_count, pb["node_count"], pb["edge_count"])
    with GPEngine(pb["symbols"], w, site_count, pb["node_count"], pb["edge_count"]) as gpu:
        for e in (cpu, gpu):
            e.set_branch_lengths(pb["branch_lengths"])
            e.process_operations(*pb["populate"])
            e.process_operations(*pb["likelihoods"])
            e.process_operations(*pb["optimize"])
        assert gpu.stats()["optimizer_scheme"] == 3
        assert np.max(np.abs(gpu.get_branch_lengths() - cpu.branch_lengths())) <= BL_ATOL
        w2 = np.where(np.arange(w.size) % 3 == 0, 2.0, 1.0)  # other class boundaries, same pattern count
        cpu2 = PortEngine(pb["symbols"], w2, int(w2.sum()), pb["node_count"], pb["edge_count"])
        gpu.set_site_patterns(pb["symbols"], w2)
        for e in (cpu2, gpu):
            e.set_branch_lengths(pb["branch_lengths"])
            e.reset_optimization_count()
            e.process_operations(*pb["populate"])
            e.process_operations(*pb["optimize"])
        assert np.max(np.abs(gpu.get_branch_lengths() - cpu2.branch_lengths())) <= BL_ATOL


@pytest.fixture
def cluster_shape(request, monkeypatch):
    c, _, t = str(request.param).partition("x")
    monkeypatch.setenv("BITO_GP_OPT_CLUSTER", c)
    if t:
        monkeypatch.setenv("BITO_GP_OPT_CLUSTER_THREADS", t)
    return request.param


@pytest.mark.parametrize("cluster_shape", ["3", "6x512", "12x1024", "9"], indirect=True)
@pytest.mark.parametrize("case", ["hello", "five_taxon", "ds1_reduced_5", "fluA", "ds1"])
def test_cluster_optimizer_any_cluster_size(cuda_engine_lib, case, cluster_shape):
    """Cluster sizes need not be powers of two (a GPC's SMs rarely divide by 16): 3, 6, 9, 12 blocks,
    512-thread blocks."""
    fx = Fixture(case)
    if "brent" not in fx.methods:
        pytest.skip("fixture has no plain-Brent sweep")
    with make_cuda(fx, 0) as e:
        check_sweeps(e, fx, 0, "brent")
        assert e.stats()["optimizer_scheme"] == 2


def _check_jacobi_sweep(gpu, cpu, blo):
    """One batched (Jacobi) Brent sweep, CUDA vs oracle: branch lengths within 1e-6, except edges whose
    first parabolic step sits on an acceptance boundary (two builds of the unmodified reference disagree
    on those too, oracle/ref_jacobi_sweep_sensitivity.py): for each such edge the two lengths must lie
    inside Brent's own tolerance AND give the same objective value to 1e-9 (a flat objective)."""
    for eng in (cpu, gpu):
        eng.process_operations(*blo)
    got, want = gpu.get_branch_lengths(), cpu.branch_lengths()
    off = np.nonzero(np.abs(got - want) > BL_ATOL)[0]
    assert off.size <= 0.005 * want.size, off.size
    tol = 2.0 ** -9
    assert np.all(np.abs(np.log(got[off]) - np.log(want[off])) <= 4 * (tol * np.abs(np.log(want[off])) + tol / 4))
    by_edge = {int(r[3]): (int(r[1]), int(r[2])) for r in blo[0]}
    for g in off:
        leafward, rootward = by_edge[int(g)]
        values = []
        for t in (got[g], want[g]):
            bl = want.copy()
            bl[g] = t
            cpu.set_branch_lengths(bl)
            values.append(cpu.log_likelihood_and_derivatives(int(g), rootward, leafward)[0])
        assert abs(values[0] - values[1]) <= 1e-9 * abs(values[1])
    return int(off.size)


@pytest.mark.parametrize("thr", [1e-40, 0.5])
def test_headline_dag_subsample_matches_oracle(cuda_engine_lib, thr, monkeypatch):
    """The DAG bench.py's headline number is measured on (configs[4]: 1000 taxa, DAG from 5000 trees,
    5657 nodes, 11 144 edges, 72+ dependency levels) on the first 768 patterns of the rank-0 shard, against
    the oracle: per-edge and per-pattern log-likelihoods, marginal (1e-9), rescaling counts bit-exact -
    at the DEFAULT threshold 1e-40 they are non-zero on this DAG - then a batched sweep (rounds scheme)
    and the same sweep under the pipelined cluster scheme."""
    from bito_b200.gp_engine import GPEngine
    from bito_b200.synthetic import make_named_workload
    from oracle.port_engine import PortEngine
    wl = make_named_workload("synthetic-1000taxa-1Mpat-5000trees", pattern_count=768)
    dag = wl.dag
    assert dag.node_count == 5657 and dag.edge_count == 11144
    pop, lik = wl.ops("populate_plvs"), wl.ops("compute_likelihoods")
    blo = wl.ops("batched_branch_length_optimization")

    def make_cpu():
        e = PortEngine(wl.symbols, wl.weights, wl.site_count, dag.node_count, dag.edge_count, wl.sbn_prior,
                       wl.unconditional, wl.inverted, thr)
        e.process_operations(*pop)
        e.process_operations(*lik)
        return e

    cpu = make_cpu()
    counts = cpu.rescaling_counts()
    assert counts.max() > 0, "this DAG is deep enough to rescale at any threshold"
    for scheme in ("0", "3"):
        monkeypatch.setenv("BITO_GP_OPT_SCHEME", scheme)
        monkeypatch.setenv("BITO_GP_OPT_RING_EDGES", "500")
        with GPEngine(wl.symbols, wl.weights, wl.site_count, dag.node_count, dag.edge_count, thr,
                      sbn_prior=wl.sbn_prior, unconditional_node_probabilities=wl.unconditional,
                      inverted_sbn_prior=wl.inverted) as gpu:
            gpu.process_operations(*pop)
            gpu.process_operations(*lik)
            assert gpu.stats()["levels_last"] >= 1
            assert rel_err(gpu.get_per_gpcsp_log_likelihoods(), cpu.per_gpcsp_log_likelihoods()) <= LL_RTOL
            assert rel_err(gpu.get_log_marginal_likelihood(), cpu.log_marginal_likelihood()) <= LL_RTOL
            assert rel_err(gpu.get_log_likelihood_matrix(), cpu.log_likelihood_matrix()) <= LL_RTOL
            assert rel_err(gpu.get_per_pattern_log_marginal(), cpu.per_pattern_log_marginal()) <= LL_RTOL
            assert np.array_equal(gpu.get_rescaling_counts(), counts)
            assert gpu.stats()["device_status_bits"] == 0
            _check_jacobi_sweep(gpu, cpu, blo)
            assert gpu.stats()["optimizer_scheme"] == int(scheme)
        cpu.close()
        cpu = make_cpu()
    cpu.close()


def test_accumulate_groups_larger_than_one_matrix_chunk(cuda_engine_lib):
    """k_node stages 64 transition matrices at a time; a subsplit with more than 64 child edges in one
    clade (or a fused node whose two groups total more) takes its multi-chunk branch. Star-shaped lists:
    100 increments into one PLV, then a fused node of 70 + 45 increments and the Multiply of the two."""
    from bito_b200.gp_engine import GPEngine
    from bito_b200.gp_operation import GPOperationVector
    from oracle.port_engine import PortEngine
    rng = np.random.default_rng(77)
    taxa, patterns = 120, 777
    sym = rng.integers(0, 5, size=(taxa, patterns)).astype(np.uint8)
    w = rng.integers(1, 4, size=patterns).astype(np.float64)
    N = taxa + 4
    n_edges = taxa + 8
    P_, PHR, PHL = 0, N, 2 * N
    a, b = taxa, taxa + 1  # two internal nodes
    ops = GPOperationVector()
    for n in (a, b):
        for t in (P_, PHR, PHL):
            ops.zero_plv(t + n)
    ops.append_after_prep_for_marginalization([(PHR + a, 1 + k, P_ + k) for k in range(100)])
    ops.append_after_prep_for_marginalization([(PHR + b, 1 + k, P_ + k) for k in range(70)])
    ops.append_after_prep_for_marginalization([(PHL + b, 1 + 70 + k, P_ + 70 + k) for k in range(45)])
    ops.multiply(P_ + b, PHR + b, PHL + b)
    lik = GPOperationVector()
    lik.likelihood(0, P_ + b, PHR + a)  # any pair of dense PLVs
    bl = rng.uniform(0.01, 0.4, size=n_edges)
    q = rng.uniform(0.2, 1.0, size=n_edges)
    for thr in (1e-40, 0.5):
        cpu = PortEngine(sym, w, int(w.sum()), N, n_edges, rescaling_threshold=thr)
        with GPEngine(sym, w, int(w.sum()), N, n_edges, thr) as gpu:
            for e in (cpu, gpu):
                e.set_sbn_parameters(q)
                e.set_branch_lengths(bl)
                e.process_operations(*ops.arrays())
                e.process_operations(*lik.arrays())
            for plv in (PHR + a, PHR + b, PHL + b, P_ + b):
                want = cpu.get_plv(plv)
                assert np.max(np.abs(gpu.get_plv(plv) - want)) <= 1e-12 * np.max(np.abs(want)), plv
            assert np.array_equal(gpu.get_rescaling_counts(), cpu.rescaling_counts())
            assert rel_err(gpu.get_per_gpcsp_log_likelihoods(), cpu.per_gpcsp_log_likelihoods()) <= LL_RTOL
        cpu.close()


def test_multiply_before_its_operand_is_written(cuda_engine_lib):
    """Multiply(a, b, c) with b still empty, followed by an Increment into b in the same
    list. The zero-elision decision must look at the whole list (b does become dense), not at the PLV
    kinds at the moment the Multiply is seen: a must get storage and the product must be zero at the
    time of the Multiply (b is empty then), exactly as the reference computes it."""
    from bito_b200.gp_engine import GPEngine
    from bito_b200.gp_operation import GPOperationVector
    from oracle.port_engine import PortEngine
    rng = np.random.default_rng(5)
    taxa, patterns = 4, 300
    sym = rng.integers(0, 4, size=(taxa, patterns)).astype(np.uint8)
    w = np.ones(patterns)
    N, n_edges = 7, 7
    a, b = N + 4, N + 5  # PHatRight of nodes 4 and 5: never written on a fresh engine
    ops = GPOperationVector()
    ops.multiply(a, b, 0)                                 # b is empty: a = 0
    ops.increment_with_weighted_evolved_plv(b, 1, 1)      # ... and only now b gets content
    ops.multiply(2 * N + 4, b, 0)                         # a real product
    cpu = PortEngine(sym, w, patterns, N, n_edges)
    with GPEngine(sym, w, patterns, N, n_edges) as gpu:
        for e in (cpu, gpu):
            e.process_operations(*ops.arrays())
        for plv in (a, b, 2 * N + 4):
            assert np.allclose(gpu.get_plv(plv), cpu.get_plv(plv), rtol=1e-13, atol=0), plv
        assert not gpu.get_plv(a).any() and gpu.get_plv(2 * N + 4).any()
        # the engine survives (no illegal address) and runs the list again
        gpu.process_operations(*ops.arrays())
        assert np.array_equal(gpu.get_rescaling_counts(), cpu.rescaling_counts())
    cpu.close()


def test_reupload_with_other_weights_replays_fresh_graphs(cuda_engine_lib, monkeypatch):
    """new weights with the same number of rho rows but other class boundaries, graphs ON and
    the cluster optimiser in use: the captured launches hold the class segments by value, so the graphs
    must be dropped; branch lengths must follow the new weights."""
    from bito_b200.gp_engine import GPEngine
    from oracle.port_engine import PortEngine
    monkeypatch.setenv("BITO_GP_OPT_CLUSTER", "2")
    rng = np.random.default_rng(21)
    pb = _random_problem(rng, 9, 500)
    w1 = np.where(np.arange(500) < 300, 1.0, 2.0)
    w2 = np.where(np.arange(500) < 200, 1.0, 2.0)
    with GPEngine(pb["symbols"], w1, int(w1.sum()), pb["node_count"], pb["edge_count"]) as gpu:
        for w in (w1, w2, w1):
            cpu = PortEngine(pb["symbols"], w, int(w.sum()), pb["node_count"], pb["edge_count"])
            gpu.set_site_patterns(pb["symbols"], w)
            for e in (cpu, gpu):
                e.set_branch_lengths(pb["branch_lengths"])
                e.reset_optimization_count()
                e.process_operations(*pb["populate"])
                e.process_operations(*pb["optimize"])
            assert np.max(np.abs(gpu.get_branch_lengths() - cpu.branch_lengths())) <= BL_ATOL
            cpu.close()
        assert gpu.stats()["graph_launches"] > 0 and gpu.stats()["optimizer_scheme"] == 2


def test_stale_programs_are_evicted(cuda_engine_lib):
    """after a resize every cached program is dead; they must be freed (the NNI search grows
    the DAG and issues new lists every iteration), and the live cache stays bounded."""
    fx = Fixture("five_taxon")
    with make_cuda(fx) as e:
        pop, lik = fx.ops("populate_plvs"), fx.ops("compute_likelihoods")
        e.process_operations(*pop)
        e.process_operations(*lik)      # allocates the log-likelihood rows: the first program is stale and freed
        e.process_operations(*pop)
        e.process_operations(*lik)
        st = e.stats()
        assert st["programs_cached"] == 2 and st["programs_evicted"] == 1 and st["programs_compiled"] == 3
        e.grow_spare_plvs(24)          # invalidates every program
        e.process_operations(*pop)      # recompiled; the two stale ones are freed
        st = e.stats()
        # (the recompiled list replaces its own stale entry in place; the other stale program is evicted)
        assert st["programs_cached"] == 1 and st["programs_evicted"] == 2
        # many distinct lists: the cache is bounded (least recently used go first)
        for k in range(60):
            ops = np.array([[0, fx["node_count"] + 1 + (k % 3), 0, 0, 0, 0]] * (k + 1), dtype=np.int64)
            e.process_operations(ops)
        st = e.stats()
        assert st["programs_cached"] <= 48 and st["programs_evicted"] >= 2 + 61 - 48
        e.process_operations(*pop)
        e.process_operations(*lik)
        assert rel_err(e.get_log_marginal_likelihood(), fx["t0_pass_log_marginal"]) <= LL_RTOL


# ---- Taylor-model Brent (gp_types.h, OptPass): the streamed and the cluster-resident searches ---------------
@pytest.fixture
def opt_env(request, monkeypatch):
    """BITO_GP_* settings read when an engine is created: "KEY=VAL,KEY=VAL"."""
    for kv in str(request.param).split(","):
        if kv:
            k, _, v = kv.partition("=")
            monkeypatch.setenv("BITO_GP_" + k, v)
    return request.param


def _weighted_problem(seed, taxa, patterns):
    """Random tree problem whose weights populate every class of the optimiser's layouts: 1, 2..7, general."""
    rng = np.random.default_rng(seed)
    pb = _random_problem(rng, taxa, patterns)
    w = pb["weights"].copy()
    w[::5] = rng.integers(2, 8, size=w[::5].size)
    w[3::7] = rng.uniform(0.25, 3.5, size=w[3::7].size)
    w[5::31] = rng.integers(8, 400, size=w[5::31].size)
    pb["weights"] = w
    return pb


@pytest.mark.parametrize("opt_env", ["OPT_SCHEME=0", "OPT_SCHEME=0,OPT_CHUNK_MB=1", "OPT_SCHEME=0,OPT_SEGMENTS=3",
                                     "OPT_CLUSTER=4", "OPT_CLUSTER=16,OPT_CLUSTER_THREADS=512",
                                     "OPT_CLUSTER=7,OPT_CLUSTER_THREADS=1024"], indirect=True)
@pytest.mark.parametrize("taxa,patterns", [(6, 1), (9, 257), (14, 4099), (40, 30000)])
def test_taylor_model_sweep_matches_oracle(cuda_engine_lib, taxa, patterns, opt_env):
    """A batched Brent optimisation of every edge (weights of every class) under the streamed scheme (one and
    several chunks, other segment counts) and under the cluster-resident model kernel, against the plain-C
    oracle; most objective evaluations must have been answered without a pass."""
    from bito_b200.gp_engine import GPEngine
    from oracle.port_engine import PortEngine
    pb = _weighted_problem(taxa * 7919 + patterns, taxa, patterns)
    w = pb["weights"]
    site_count = int(round(w.sum()))
    cpu = PortEngine(pb["symbols"], w, site_count, pb["node_count"], pb["edge_count"])
    with GPEngine(pb["symbols"], w, site_count, pb["node_count"], pb["edge_count"]) as gpu:
        for e in (cpu, gpu):
            e.set_branch_lengths(pb["branch_lengths"])
            e.process_operations(*pb["populate"])
            e.process_operations(*pb["likelihoods"])
            e.process_operations(*pb["optimize"])
        st = gpu.stats()
        assert st["optimizer_scheme"] == (0 if "OPT_SCHEME=0" in opt_env else 2)
        assert np.max(np.abs(gpu.get_branch_lengths() - cpu.branch_lengths())) <= BL_ATOL
        assert 0 < st["objective_passes"] <= st["objective_evaluations"] // 2
        assert st["device_status_bits"] == 0
        # a second optimisation from the optimised lengths (IsFirstOptimization is false now: the first pass
        # centres its model on the start)
        cpu.increment_optimization_count()
        gpu.increment_optimization_count()
        for e in (cpu, gpu):
            e.process_operations(*pb["populate"])
            e.process_operations(*pb["optimize"])
        # Second Jacobi sweeps of random trees are ill-conditioned on saturated (t ~ 3) and vanishing (t ~ 1e-6)
        # edges - the plain-C port and the reference itself disagree there - so an edge outside 1e-6 must have
        # the same objective value at both lengths (the CPU engine still holds the PLVs the sweep optimised against).
        got, want = gpu.get_branch_lengths(), cpu.branch_lengths().copy()
        off = np.nonzero(np.abs(got - want) > BL_ATOL)[0]
        assert off.size <= max(2, want.size // 4)
        by_edge = {int(r[3]): (int(r[1]), int(r[2])) for r in pb["optimize"][0]}
        for g in off:
            leafward, rootward = by_edge[int(g)]
            values = []
            for t in (got[g], want[g]):
                bl = want.copy()
                bl[g] = t
                cpu.set_branch_lengths(bl)
                values.append(cpu.log_likelihood_and_derivatives(int(g), rootward, leafward)[0])
            assert abs(values[0] - values[1]) <= 1e-9 * max(1.0, abs(values[1])), (int(g), got[g], want[g], values)


@pytest.mark.parametrize("opt_env", ["OPT_SCHEME=0"], indirect=True)
def test_taylor_model_agrees_with_one_pass_per_evaluation(cuda_engine_lib, opt_env, monkeypatch):
    """The same streamed sweep with the model switched off (BITO_GP_OPT_MODEL=0: every objective evaluation is
    a pass over rho): same number of evaluations, same branch lengths up to Brent decisions that sit on a
    boundary (an edge may differ only where the objective is flat to 1e-9)."""
    from bito_b200.gp_engine import GPEngine
    from bito_b200.synthetic import make_named_workload
    wl = make_named_workload("synthetic-small", pattern_count=12000)
    dag = wl.dag
    pop, blo = wl.ops("populate_plvs"), wl.ops("batched_branch_length_optimization")
    out = {}
    for model in ("1", "0"):
        monkeypatch.setenv("BITO_GP_OPT_MODEL", model)
        with GPEngine(wl.symbols, wl.weights, wl.site_count, dag.node_count, dag.edge_count, sbn_prior=wl.sbn_prior,
                      unconditional_node_probabilities=wl.unconditional, inverted_sbn_prior=wl.inverted) as e:
            e.process_operations(*pop)
            e.process_operations(*blo)
            st = e.stats()
            out[model] = (e.get_branch_lengths(), st["objective_evaluations"], st["objective_passes"])
    bl1, ev1, passes1 = out["1"]
    bl0, ev0, passes0 = out["0"]
    assert passes0 == 0 and 0 < passes1 < (ev1 * 3) // 5
    assert abs(ev1 - ev0) <= max(4, ev0 // 200)
    off = np.abs(bl1 - bl0) > BL_ATOL
    assert off.sum() <= max(1, bl0.size // 100)
    tol = 2.0 ** -9
    assert np.all(np.abs(np.log(bl1[off]) - np.log(bl0[off])) <= 4 * (tol * np.abs(np.log(bl0[off])) + tol / 4))
