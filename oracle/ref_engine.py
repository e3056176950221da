"""ctypes binding of oracle/_ref/libbito_gp_ref.so — TEST INFRASTRUCTURE ONLY.

The shared library is the UNMODIFIED reference GP path (GPDAG + CPU GPEngine,
/root/reference/src/gp_dag.cpp, gp_engine.cpp) built by oracle/Makefile plus our C-ABI
driver oracle/ref_driver.cpp. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product path never does.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BITO_REF_LIB selects another build of the same driver: `make -C oracle refmodel` (libbito_gp_ref_model.so, the
# reference GPEngine over its own GTR / HKY models, chosen by BITO_REF_MODEL when an engine is constructed) or
# `make -C oracle refvar` (other floating-point code generation).
LIB_PATH = os.environ.get("BITO_REF_LIB") or os.path.join(_HERE, "_ref", "libbito_gp_ref.so")
MODEL_LIB_PATH = os.path.join(_HERE, "_ref", "libbito_gp_ref_model.so")

OPLISTS = {
    "populate_plvs": 0,
    "compute_likelihoods": 1,
    "marginal_likelihood": 2,
    "branch_length_optimization": 3,
    "optimize_sbn_parameters": 4,
    "rootward_pass": 5,
    "leafward_pass": 6,
    "set_rootward_zero": 7,
    "set_leafward_zero": 8,
    "set_rhat_to_stationary": 9,
    "approximate_branch_length_optimization": 10,
}

# Optimization::OptimizationMethod order, /root/reference/src/optimization.hpp:28-34
OPTIMIZATION_METHODS = {
    "brent": 0,
    "brent_with_gradients": 1,
    "gradient_ascent": 2,
    "logspace_gradient_ascent": 3,
    "newton": 4,
}


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    lib = C.CDLL(LIB_PATH)
    vp, i64, f64, i32 = C.c_void_p, C.c_int64, C.c_double, C.c_int
    P = C.POINTER
    lib.ref_last_error.restype = C.c_char_p
    lib.ref_open.restype = vp
    lib.ref_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, f64, i32]
    lib.ref_open_raw.restype = vp
    lib.ref_open_raw.argtypes = [i64, i64, vp, vp, i64, i64, i64, C.c_char_p, f64, vp, vp, vp, i32]
    lib.ref_close.argtypes = [vp]
    lib.ref_sizes.argtypes = [vp, vp]
    lib.ref_patterns.argtypes = [vp, vp, vp]
    lib.ref_priors.argtypes = [vp, vp, vp, vp]
    lib.ref_edges.argtypes = [vp, vp, vp, vp]
    lib.ref_node_bitsets.argtypes = [vp, vp]
    lib.ref_node_count_with_root.restype = i64
    lib.ref_node_count_with_root.argtypes = [vp]
    lib.ref_taxon_names.restype = i64
    lib.ref_taxon_names.argtypes = [vp, vp, i64]
    lib.ref_oplist.restype = i64
    lib.ref_oplist.argtypes = [vp, i32, vp, i64, vp, i64, P(i64)]
    lib.ref_run.argtypes = [vp, vp, i64, vp]
    lib.ref_time_run.argtypes = [vp, vp, i64, vp, i32, vp]
    lib.ref_get_plv.argtypes = [vp, i64, vp]
    lib.ref_set_plv.argtypes = [vp, i64, vp, i32]
    lib.ref_get_counts.argtypes = [vp, vp]
    lib.ref_get_loglik_matrix.argtypes = [vp, vp]
    lib.ref_get_per_pattern_marginal.argtypes = [vp, vp]
    lib.ref_get_per_gpcsp_loglik.argtypes = [vp, vp]
    lib.ref_get_per_gpcsp_components.argtypes = [vp, vp]
    lib.ref_get_log_marginal.restype = f64
    lib.ref_get_log_marginal.argtypes = [vp]
    lib.ref_get_q.argtypes = [vp, vp]
    lib.ref_set_q.argtypes = [vp, vp]
    lib.ref_get_branch_lengths.argtypes = [vp, vp]
    lib.ref_set_branch_lengths.argtypes = [vp, vp]
    lib.ref_set_branch_lengths_constant.argtypes = [vp, f64]
    lib.ref_get_branch_differences.argtypes = [vp, vp]
    lib.ref_set_optimization_method.argtypes = [vp, i32]
    lib.ref_use_gradient_optimization.argtypes = [vp, i32]
    lib.ref_set_significant_digits.argtypes = [vp, i32]
    lib.ref_reset_optimization_count.argtypes = [vp]
    lib.ref_increment_optimization_count.argtypes = [vp]
    lib.ref_get_optimization_count.restype = i64
    lib.ref_get_optimization_count.argtypes = [vp]
    lib.ref_set_null_prior.argtypes = [vp]
    lib.ref_loglik_and_derivatives.argtypes = [vp, i64, i64, i64, i32, vp]
    lib.ref_transition_matrix.argtypes = [vp, f64, vp]
    lib.ref_quartet_requests.restype = i64
    lib.ref_quartet_requests.argtypes = [vp, vp, vp, vp, i64, i64, P(i64)]
    lib.ref_quartet_likelihoods.argtypes = [vp, i64, vp, vp, vp]
    lib.ref_process_quartet_requests.argtypes = [vp, i64, vp, vp, vp]
    lib.ref_get_hybrid_marginals.argtypes = [vp, vp]
    if hasattr(lib, "ref_model_eigensystem"):
        lib.ref_model_eigensystem.argtypes = [vp, vp, vp, vp, vp]
    _lib = lib
    return lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class RefEngine:
    """The reference GPEngine (+ GPDAG when opened from files), driven through ctypes."""

    def __init__(self, handle, tmpdir):
        self._h = handle
        self._tmpdir = tmpdir
        s = np.zeros(10, dtype=np.int64)
        _load().ref_sizes(self._h, _ptr(s))
        (self.taxon_count, self.pattern_count, self.site_count, self.node_count, self.edge_count,
         self.rootsplit_count, self.plv_count, self.padded_plv_count, self.topology_count,
         has_dag) = (int(x) for x in s)
        self.has_dag = bool(has_dag)

    @classmethod
    def from_files(cls, fasta, newick, rescaling_threshold=1e-40, use_gradients=False):
        lib = _load()
        tmpdir = tempfile.TemporaryDirectory(prefix="bito_ref_")
        mmap_path = os.path.join(tmpdir.name, "plv.gp")
        h = lib.ref_open(os.fsencode(fasta), os.fsencode(newick), os.fsencode(mmap_path),
                         float(rescaling_threshold), int(use_gradients))
        if not h:
            raise RuntimeError(lib.ref_last_error().decode())
        return cls(h, tmpdir)

    @classmethod
    def from_arrays(cls, symbols, weights, site_count, node_count, edge_count, q, unconditional,
                    inverted, rescaling_threshold=1e-40, use_gradients=False):
        lib = _load()
        symbols = np.ascontiguousarray(symbols, dtype=np.uint8)
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        unconditional = np.ascontiguousarray(unconditional, dtype=np.float64)
        inverted = np.ascontiguousarray(inverted, dtype=np.float64)
        taxa, P = symbols.shape
        tmpdir = tempfile.TemporaryDirectory(prefix="bito_ref_")
        mmap_path = os.path.join(tmpdir.name, "plv.gp")
        h = lib.ref_open_raw(taxa, P, _ptr(symbols), _ptr(weights), int(site_count), int(node_count),
                             int(edge_count), os.fsencode(mmap_path), float(rescaling_threshold),
                             _ptr(q), _ptr(unconditional), _ptr(inverted), int(use_gradients))
        if not h:
            raise RuntimeError(lib.ref_last_error().decode())
        return cls(h, tmpdir)

    def close(self):
        if self._h:
            _load().ref_close(self._h)
            self._h = None
        if self._tmpdir is not None:
            self._tmpdir.cleanup()
            self._tmpdir = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(_load().ref_last_error().decode())

    # ---- inputs -------------------------------------------------------------------
    def patterns(self):
        sym = np.zeros((self.taxon_count, self.pattern_count), dtype=np.uint8)
        w = np.zeros(self.pattern_count, dtype=np.float64)
        _load().ref_patterns(self._h, _ptr(sym), _ptr(w))
        return sym, w

    def priors(self):
        q = np.zeros(self.edge_count)
        un = np.zeros(self.node_count)
        inv = np.zeros(self.edge_count)
        _load().ref_priors(self._h, _ptr(q), _ptr(un), _ptr(inv))
        return q, un, inv

    def edges(self):
        parent = np.zeros(self.edge_count, dtype=np.int64)
        child = np.zeros(self.edge_count, dtype=np.int64)
        on_left = np.zeros(self.edge_count, dtype=np.int32)
        self._check(_load().ref_edges(self._h, _ptr(parent), _ptr(child), _ptr(on_left)))
        return parent, child, on_left

    def node_bitsets(self):
        n = _load().ref_node_count_with_root(self._h)
        width = 2 * self.taxon_count
        buf = np.zeros(n * width, dtype=np.uint8)
        self._check(_load().ref_node_bitsets(self._h, _ptr(buf)))
        return [bytes(buf[i * width:(i + 1) * width]).decode() for i in range(n)]

    def taxon_names(self):
        buf = C.create_string_buffer(1 << 20)
        n = _load().ref_taxon_names(self._h, buf, len(buf))
        if n < 0:
            raise RuntimeError("taxon name buffer too small")
        return buf.value.decode().split("\n")[:-1]

    def oplist(self, which):
        lib = _load()
        which = OPLISTS[which] if isinstance(which, str) else int(which)
        vec_len = C.c_int64(0)
        n = lib.ref_oplist(self._h, which, None, 0, None, 0, C.byref(vec_len))
        if n < 0:
            raise RuntimeError(lib.ref_last_error().decode())
        ops = np.zeros((n, 6), dtype=np.int64)
        vec = np.zeros(max(1, vec_len.value), dtype=np.int64)
        n2 = lib.ref_oplist(self._h, which, _ptr(ops), n, _ptr(vec), vec.size, C.byref(vec_len))
        if n2 != n:
            raise RuntimeError(lib.ref_last_error().decode())
        return ops, vec[:vec_len.value]

    # ---- execution ----------------------------------------------------------------
    def process_operations(self, ops, vec=None):
        ops = np.ascontiguousarray(ops, dtype=np.int64).reshape(-1, 6)
        vec = np.ascontiguousarray(vec if vec is not None else np.zeros(1), dtype=np.int64)
        self._check(_load().ref_run(self._h, _ptr(ops), ops.shape[0], _ptr(vec)))

    def time_operations(self, ops, vec=None, repeats=1):
        ops = np.ascontiguousarray(ops, dtype=np.int64).reshape(-1, 6)
        vec = np.ascontiguousarray(vec if vec is not None else np.zeros(1), dtype=np.int64)
        out = np.zeros(repeats)
        self._check(_load().ref_time_run(self._h, _ptr(ops), ops.shape[0], _ptr(vec), repeats, _ptr(out)))
        return out

    # ---- state --------------------------------------------------------------------
    def get_plv(self, plv_id):
        out = np.zeros((self.pattern_count, 4))
        self._check(_load().ref_get_plv(self._h, int(plv_id), _ptr(out)))
        return out

    def set_plv(self, plv_id, values, count=0):
        values = np.ascontiguousarray(values, dtype=np.float64)
        assert values.shape == (self.pattern_count, 4)
        self._check(_load().ref_set_plv(self._h, int(plv_id), _ptr(values), int(count)))

    def rescaling_counts(self):
        out = np.zeros(self.padded_plv_count, dtype=np.int32)
        _load().ref_get_counts(self._h, _ptr(out))
        return out

    def log_likelihood_matrix(self):
        out = np.zeros((self.edge_count, self.pattern_count))
        _load().ref_get_loglik_matrix(self._h, _ptr(out))
        return out

    def per_pattern_log_marginal(self):
        out = np.zeros(self.pattern_count)
        _load().ref_get_per_pattern_marginal(self._h, _ptr(out))
        return out

    def per_gpcsp_log_likelihoods(self):
        out = np.zeros(self.edge_count)
        _load().ref_get_per_gpcsp_loglik(self._h, _ptr(out))
        return out

    def per_gpcsp_components_of_full_log_marginal(self):
        out = np.zeros(self.edge_count)
        _load().ref_get_per_gpcsp_components(self._h, _ptr(out))
        return out

    def log_marginal_likelihood(self):
        return float(_load().ref_get_log_marginal(self._h))

    def sbn_parameters(self):
        out = np.zeros(self.edge_count)
        _load().ref_get_q(self._h, _ptr(out))
        return out

    def set_sbn_parameters(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.size == self.edge_count
        _load().ref_set_q(self._h, _ptr(q))

    def branch_lengths(self):
        out = np.zeros(self.edge_count)
        _load().ref_get_branch_lengths(self._h, _ptr(out))
        return out

    def set_branch_lengths(self, t):
        t = np.ascontiguousarray(t, dtype=np.float64)
        assert t.size == self.edge_count
        _load().ref_set_branch_lengths(self._h, _ptr(t))

    def set_branch_lengths_to_constant(self, t):
        _load().ref_set_branch_lengths_constant(self._h, float(t))

    def branch_length_differences(self):
        out = np.zeros(self.edge_count)
        _load().ref_get_branch_differences(self._h, _ptr(out))
        return out

    def set_optimization_method(self, method):
        method = OPTIMIZATION_METHODS[method] if isinstance(method, str) else int(method)
        _load().ref_set_optimization_method(self._h, method)

    def use_gradient_optimization(self, use):
        _load().ref_use_gradient_optimization(self._h, int(use))

    def set_significant_digits_for_optimization(self, digits):
        _load().ref_set_significant_digits(self._h, int(digits))

    def reset_optimization_count(self):
        _load().ref_reset_optimization_count(self._h)

    def increment_optimization_count(self):
        _load().ref_increment_optimization_count(self._h)

    def optimization_count(self):
        return int(_load().ref_get_optimization_count(self._h))

    def set_null_prior(self):
        _load().ref_set_null_prior(self._h)

    def log_likelihood_and_derivatives(self, gpcsp, rootward, leafward, two=False):
        out = np.zeros(3)
        self._check(_load().ref_loglik_and_derivatives(self._h, int(gpcsp), int(rootward),
                                                       int(leafward), int(two), _ptr(out)))
        return tuple(out[:3 if two else 2])

    def model_eigensystem(self):
        """(V, V^-1, eigenvalues, frequencies) of the substitution model the engine was built with."""
        v, vinv, lam, pi = np.zeros((4, 4)), np.zeros((4, 4)), np.zeros(4), np.zeros(4)
        _load().ref_model_eigensystem(self._h, _ptr(v), _ptr(vinv), _ptr(lam), _ptr(pi))
        return v, vinv, lam, pi

    def transition_matrix(self, t):
        out = np.zeros((4, 4))
        _load().ref_transition_matrix(self._h, float(t), _ptr(out))
        return out

    # ---- quartet hybrid marginals (gp_engine.cpp:748-816, gp_dag.cpp:413-458) ------------
    def quartet_requests(self):
        """Every request GPInstance::CalculateHybridMarginals issues, flattened: central (n),
        tip_counts (n x 4: rootward, sister, rotated, sorted), tips (total x 3: node, plv, gpcsp)."""
        lib = _load()
        n_tips = C.c_int64(0)
        n = lib.ref_quartet_requests(self._h, None, None, None, 0, 0, C.byref(n_tips))
        if n < 0:
            raise RuntimeError(lib.ref_last_error().decode())
        central = np.zeros(n, dtype=np.int64)
        counts = np.zeros((n, 4), dtype=np.int32)
        tips = np.zeros((max(1, n_tips.value), 3), dtype=np.int64)
        n2 = lib.ref_quartet_requests(self._h, _ptr(central), _ptr(counts), _ptr(tips), n, tips.shape[0],
                                      C.byref(n_tips))
        if n2 != n:
            raise RuntimeError(lib.ref_last_error().decode())
        return central, counts, tips[:n_tips.value]

    def calculate_quartet_hybrid_likelihoods(self, central, counts, tips):
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        tips = np.ascontiguousarray(tips, dtype=np.int64)
        out = np.zeros(int(np.prod(counts.astype(np.int64))))
        self._check(_load().ref_quartet_likelihoods(self._h, int(central), _ptr(counts), _ptr(tips), _ptr(out)))
        return out

    def process_quartet_hybrid_requests(self, central, counts, tips):
        central = np.ascontiguousarray(central, dtype=np.int64)
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        tips = np.ascontiguousarray(tips, dtype=np.int64)
        self._check(_load().ref_process_quartet_requests(self._h, central.size, _ptr(central), _ptr(counts),
                                                         _ptr(tips)))

    def hybrid_marginals(self):
        out = np.zeros(self.edge_count)
        _load().ref_get_hybrid_marginals(self._h, _ptr(out))
        return out
