"""ctypes binding of oracle/libgp_oracle.so (the plain-C restatement, oracle/gp_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. Same method names as oracle.ref_engine.RefEngine so a
test can run against either.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgp_oracle.so")

OPTIMIZATION_METHODS = {
    "brent": 0,
    "brent_with_gradients": 1,
    "gradient_ascent": 2,
    "logspace_gradient_ascent": 3,
    "newton": 4,
}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gp_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "port"])
    return LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB_PATH)
    vp, i64, f64, i32 = C.c_void_p, C.c_int64, C.c_double, C.c_int
    lib.gpo_last_error.restype = C.c_char_p
    lib.gpo_create.restype = vp
    lib.gpo_create.argtypes = [i64, i64, vp, vp, i64, i64, i64, f64, vp, vp, vp, i32]
    lib.gpo_destroy.argtypes = [vp]
    lib.gpo_run.argtypes = [vp, vp, i64, vp]
    lib.gpo_plv_count.restype = i64
    lib.gpo_plv_count.argtypes = [vp]
    lib.gpo_padded_plv_count.restype = i64
    lib.gpo_padded_plv_count.argtypes = [vp]
    lib.gpo_get_plv.argtypes = [vp, i64, vp]
    lib.gpo_set_plv.argtypes = [vp, i64, vp, i32]
    for name in ("gpo_get_counts", "gpo_get_loglik_matrix", "gpo_get_per_pattern_marginal",
                 "gpo_get_per_gpcsp_loglik", "gpo_get_per_gpcsp_components", "gpo_get_q", "gpo_set_q",
                 "gpo_get_branch_lengths", "gpo_set_branch_lengths", "gpo_get_branch_differences"):
        getattr(lib, name).argtypes = [vp, vp]
    lib.gpo_get_log_marginal.restype = f64
    lib.gpo_get_log_marginal.argtypes = [vp]
    lib.gpo_set_branch_lengths_constant.argtypes = [vp, f64]
    lib.gpo_set_optimization_method.argtypes = [vp, i32]
    lib.gpo_use_gradient_optimization.argtypes = [vp, i32]
    lib.gpo_set_significant_digits.argtypes = [vp, i32]
    lib.gpo_reset_optimization_count.argtypes = [vp]
    lib.gpo_increment_optimization_count.argtypes = [vp]
    lib.gpo_get_optimization_count.restype = i64
    lib.gpo_get_optimization_count.argtypes = [vp]
    lib.gpo_set_null_prior.argtypes = [vp]
    lib.gpo_loglik_and_derivatives.argtypes = [vp, i64, i64, i64, vp]
    lib.gpo_transition_matrix.argtypes = [f64, vp]
    lib.gpo_quartet_likelihoods.argtypes = [vp, i64, vp, vp, vp]
    lib.gpo_process_quartet_requests.argtypes = [vp, i64, vp, vp, vp]
    lib.gpo_get_hybrid_marginals.argtypes = [vp, vp]
    lib.gpo_feval_count.restype = i64
    lib.gpo_feval_count.argtypes = [vp]
    lib.gpo_set_strict.argtypes = [vp, i32]
    lib.gpo_log_add.restype = f64
    lib.gpo_log_add.argtypes = [f64, f64]
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def log_add(x, y):
    return float(_load().gpo_log_add(float(x), float(y)))


BRENT_FUNC = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)


def brent_minimize(f, guess, lo, hi, significant_digits, max_iter):
    """Optimization::BrentMinimize (optimization.hpp:71-188) on a Python callable."""
    lib = _load()
    lib.gpo_brent_minimize.argtypes = [BRENT_FUNC, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                       C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    x = C.c_double()
    fx = C.c_double()
    cb = BRENT_FUNC(lambda v, _ctx: float(f(v)))
    lib.gpo_brent_minimize(cb, None, guess, lo, hi, significant_digits, max_iter, C.byref(x), C.byref(fx))
    return x.value, fx.value


def set_model(eigenvectors=None, inverse_eigenvectors=None, eigenvalues=None, frequencies=None):
    """Installs a substitution model for every PortEngine of this process (None: back to JC69)."""
    lib = _load()
    lib.gpo_set_model.argtypes = [C.c_void_p] * 4
    if eigenvectors is None:
        lib.gpo_set_model(None, None, None, None)
        return
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (eigenvectors, inverse_eigenvectors, eigenvalues,
                                                                  frequencies)]
    assert [a.size for a in arrs] == [16, 16, 4, 4]
    lib.gpo_set_model(*[_ptr(a) for a in arrs])


def transition_matrix(t):
    out = np.zeros((4, 4))
    _load().gpo_transition_matrix(float(t), _ptr(out))
    return out


class PortEngine:
    def __init__(self, symbols, weights, site_count, node_count, edge_count, q=None, unconditional=None,
                 inverted=None, rescaling_threshold=1e-40, use_gradients=False):
        lib = _load()
        symbols = np.ascontiguousarray(symbols, dtype=np.uint8)
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        self.taxon_count, self.pattern_count = symbols.shape
        self.site_count, self.node_count, self.edge_count = int(site_count), int(node_count), int(edge_count)

        def opt(a, n):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            assert a.size == n
            return a

        q, unconditional, inverted = opt(q, edge_count), opt(unconditional, node_count), opt(inverted, edge_count)
        self._h = lib.gpo_create(self.taxon_count, self.pattern_count, _ptr(symbols), _ptr(weights),
                                 self.site_count, self.node_count, self.edge_count, float(rescaling_threshold),
                                 _ptr(q), _ptr(unconditional), _ptr(inverted), int(use_gradients))
        if not self._h:
            raise RuntimeError(lib.gpo_last_error().decode())
        self.plv_count = int(lib.gpo_plv_count(self._h))
        self.padded_plv_count = int(lib.gpo_padded_plv_count(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _load().gpo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(_load().gpo_last_error().decode())

    def process_operations(self, ops, vec=None):
        ops = np.ascontiguousarray(ops, dtype=np.int64).reshape(-1, 6)
        vec = np.ascontiguousarray(vec if vec is not None else np.zeros(1), dtype=np.int64)
        self._check(_load().gpo_run(self._h, _ptr(ops), ops.shape[0], _ptr(vec)))

    def _vec(self, fn, n, dtype=np.float64):
        out = np.zeros(n, dtype=dtype)
        getattr(_load(), fn)(self._h, _ptr(out))
        return out

    def get_plv(self, plv_id):
        out = np.zeros((self.pattern_count, 4))
        self._check(_load().gpo_get_plv(self._h, int(plv_id), _ptr(out)))
        return out

    def set_plv(self, plv_id, values, count=0):
        values = np.ascontiguousarray(values, dtype=np.float64)
        assert values.shape == (self.pattern_count, 4)
        self._check(_load().gpo_set_plv(self._h, int(plv_id), _ptr(values), int(count)))

    def rescaling_counts(self):
        return self._vec("gpo_get_counts", self.padded_plv_count, np.int32)

    def log_likelihood_matrix(self):
        return self._vec("gpo_get_loglik_matrix", self.edge_count * self.pattern_count).reshape(
            self.edge_count, self.pattern_count)

    def per_pattern_log_marginal(self):
        return self._vec("gpo_get_per_pattern_marginal", self.pattern_count)

    def per_gpcsp_log_likelihoods(self):
        return self._vec("gpo_get_per_gpcsp_loglik", self.edge_count)

    def per_gpcsp_components_of_full_log_marginal(self):
        return self._vec("gpo_get_per_gpcsp_components", self.edge_count)

    def log_marginal_likelihood(self):
        return float(_load().gpo_get_log_marginal(self._h))

    def sbn_parameters(self):
        return self._vec("gpo_get_q", self.edge_count)

    def set_sbn_parameters(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.size == self.edge_count
        _load().gpo_set_q(self._h, _ptr(q))

    def branch_lengths(self):
        return self._vec("gpo_get_branch_lengths", self.edge_count)

    def set_branch_lengths(self, t):
        t = np.ascontiguousarray(t, dtype=np.float64)
        assert t.size == self.edge_count
        _load().gpo_set_branch_lengths(self._h, _ptr(t))

    def set_branch_lengths_to_constant(self, t):
        _load().gpo_set_branch_lengths_constant(self._h, float(t))

    def branch_length_differences(self):
        return self._vec("gpo_get_branch_differences", self.edge_count)

    def set_optimization_method(self, method):
        method = OPTIMIZATION_METHODS[method] if isinstance(method, str) else int(method)
        _load().gpo_set_optimization_method(self._h, method)

    def use_gradient_optimization(self, use):
        _load().gpo_use_gradient_optimization(self._h, int(use))

    def set_significant_digits_for_optimization(self, digits):
        _load().gpo_set_significant_digits(self._h, int(digits))

    def reset_optimization_count(self):
        _load().gpo_reset_optimization_count(self._h)

    def increment_optimization_count(self):
        _load().gpo_increment_optimization_count(self._h)

    def optimization_count(self):
        return int(_load().gpo_get_optimization_count(self._h))

    def set_null_prior(self):
        _load().gpo_set_null_prior(self._h)

    def log_likelihood_and_derivatives(self, gpcsp, rootward, leafward, two=False):
        out = np.zeros(3)
        _load().gpo_loglik_and_derivatives(self._h, int(gpcsp), int(rootward), int(leafward), _ptr(out))
        return tuple(out[:3 if two else 2])

    def transition_matrix(self, t):
        return transition_matrix(t)

    def set_strict(self, strict=True):
        _load().gpo_set_strict(self._h, int(strict))

    def feval_count(self):
        return int(_load().gpo_feval_count(self._h))

    # ---- quartet hybrid marginals (gp_engine.cpp:748-816) ----------------------------------
    def calculate_quartet_hybrid_likelihoods(self, central, counts, tips):
        counts = np.ascontiguousarray(counts, dtype=np.int32).reshape(4)
        tips = np.ascontiguousarray(tips, dtype=np.int64).reshape(-1, 3)
        out = np.zeros(int(np.prod(counts.astype(np.int64))))
        self._check(_load().gpo_quartet_likelihoods(self._h, int(central), _ptr(counts), _ptr(tips), _ptr(out)))
        return out

    def process_quartet_hybrid_requests(self, central, counts, tips):
        central = np.ascontiguousarray(central, dtype=np.int64).reshape(-1)
        counts = np.ascontiguousarray(counts, dtype=np.int32).reshape(-1, 4)
        tips = np.ascontiguousarray(tips, dtype=np.int64).reshape(-1, 3)
        self._check(_load().gpo_process_quartet_requests(self._h, central.size, _ptr(central), _ptr(counts),
                                                         _ptr(tips)))

    def hybrid_marginals(self):
        out = np.zeros(self.edge_count)
        _load().gpo_get_hybrid_marginals(self._h, _ptr(out))
        return out
