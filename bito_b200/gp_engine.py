"""Python mirror of the reference `GPEngine` (/root/reference/src/gp_engine.hpp:24-236) over
the C-ABI of include/bito_gp.h. Method names are the snake_case of the reference's; every
array crossing this boundary is a host numpy array (the engine stages it to/from HBM).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .gp_operation import as_arrays

DEFAULT_RESCALING_THRESHOLD = 1e-40  # gp_engine.hpp:285

# Optimization::OptimizationMethod, optimization.hpp:28-34
OPTIMIZATION_METHODS = {
    "brent": 0, "BrentOptimization": 0,
    "brent_with_gradients": 1, "BrentOptimizationWithGradients": 1,
    "gradient_ascent": 2, "GradientAscentOptimization": 2,
    "logspace_gradient_ascent": 3, "LogSpaceGradientAscentOptimization": 3,
    "newton": 4, "NewtonOptimization": 4,
}


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a, n=None, name="array"):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError(f"{name} has {a.size} entries, expected {n}")
    return a


class GPEngine:
    """B200 GPEngine. Constructor arguments follow GPEngine::GPEngine (gp_engine.hpp:26-29);
    `symbols`/`weights` are SitePattern::GetPatterns()/GetWeights() of THIS rank's shard."""

    def __init__(self, symbols, weights, site_count, node_count, gpcsp_count,
                 rescaling_threshold=DEFAULT_RESCALING_THRESHOLD, sbn_prior=None,
                 unconditional_node_probabilities=None, inverted_sbn_prior=None, use_gradients=False,
                 device=0, flags=0, max_device_bytes=0, spare_node_count=0, spare_gpcsp_count=0,
                 symbols_on_device=False):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        if symbols_on_device:
            taxa, patterns = symbols.shape  # torch tensors on the engine's device
        else:
            symbols = np.ascontiguousarray(symbols, dtype=np.uint8)
            weights = _f64(weights, name="weights")
            if symbols.ndim != 2 or weights.size != symbols.shape[1]:
                raise ValueError("symbols must be taxa x patterns and weights one per pattern")
            taxa, patterns = symbols.shape
        cfg = _lib.Config(abi_version=_lib.ABI_VERSION, device=int(device), taxon_count=int(taxa),
                          pattern_count=int(patterns), site_count=int(site_count),
                          node_count=int(node_count), gpcsp_count=int(gpcsp_count),
                          rescaling_threshold=float(rescaling_threshold),
                          use_gradients=int(bool(use_gradients)),
                          spare_node_count=int(spare_node_count),
                          spare_gpcsp_count=int(spare_gpcsp_count), flags=int(flags),
                          max_device_bytes=int(max_device_bytes))
        self._check(self._lib.bito_gp_create(C.byref(cfg), C.byref(self._h)))
        self.taxon_count, self.pattern_count, self.site_count = int(taxa), int(patterns), int(site_count)
        if symbols_on_device:
            self._check(self._lib.bito_gp_set_site_patterns_device(
                self._h, C.c_void_p(symbols.data_ptr()), C.c_void_p(weights.data_ptr())))
        else:
            self._check(self._lib.bito_gp_set_site_patterns(self._h, _ptr(symbols), _ptr(weights)))
        if sbn_prior is not None:
            self.initialize_priors(sbn_prior, unconditional_node_probabilities, inverted_sbn_prior)

    # ---- lifetime ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.bito_gp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self._lib.bito_gp_last_error().decode())

    # ---- counts (gp_engine.hpp:198-234) --------------------------------------------------
    @property
    def node_count(self):
        return int(self._lib.bito_gp_get_node_count(self._h))

    @property
    def plv_count(self):
        return int(self._lib.bito_gp_get_plv_count(self._h))

    @property
    def padded_plv_count(self):
        return int(self._lib.bito_gp_get_padded_plv_count(self._h))

    @property
    def gpcsp_count(self):
        return int(self._lib.bito_gp_get_gpcsp_count(self._h))

    edge_count = gpcsp_count

    @property
    def padded_gpcsp_count(self):
        return int(self._lib.bito_gp_get_padded_gpcsp_count(self._h))

    # ---- site patterns ------------------------------------------------------------------------
    def set_site_patterns(self, symbols, weights):
        """Re-upload this shard's alignment (same shape) from host memory."""
        symbols = np.ascontiguousarray(symbols, dtype=np.uint8)
        weights = _f64(weights, self.pattern_count, "weights")
        if symbols.shape != (self.taxon_count, self.pattern_count):
            raise ValueError("symbols must keep the shape the engine was created with")
        self._check(self._lib.bito_gp_set_site_patterns(self._h, _ptr(symbols), _ptr(weights)))

    # ---- priors ---------------------------------------------------------------------------
    def initialize_priors(self, sbn_prior, unconditional_node_probabilities, inverted_sbn_prior):
        q = _f64(sbn_prior, self.gpcsp_count, "sbn_prior")
        un = _f64(unconditional_node_probabilities, self.node_count, "unconditional_node_probabilities")
        inv = _f64(inverted_sbn_prior, self.gpcsp_count, "inverted_sbn_prior")
        self._check(self._lib.bito_gp_initialize_priors(self._h, _ptr(q), _ptr(un), _ptr(inv)))

    def set_null_prior(self):
        self._check(self._lib.bito_gp_set_null_prior(self._h))

    # ---- the hot call ----------------------------------------------------------------------
    def process_operations(self, ops, vec=None):
        ops, vec = as_arrays(ops, vec)
        self._check(self._lib.bito_gp_process_operations(self._h, _ptr(ops), ops.shape[0],
                                                         _ptr(vec) if vec.size else None, vec.size))

    # ---- branch lengths / optimiser ----------------------------------------------------------
    def set_branch_lengths(self, branch_lengths):
        t = _f64(branch_lengths, self.gpcsp_count, "branch_lengths")
        self._check(self._lib.bito_gp_set_branch_lengths(self._h, _ptr(t)))

    def set_branch_lengths_to_constant(self, branch_length):
        self._check(self._lib.bito_gp_set_branch_lengths_to_constant(self._h, float(branch_length)))

    def set_branch_lengths_to_default(self):
        self._check(self._lib.bito_gp_set_branch_lengths_to_default(self._h))

    def get_branch_lengths(self, start=0, length=None):
        length = self.gpcsp_count - start if length is None else length
        out = np.zeros(length)
        self._check(self._lib.bito_gp_get_branch_lengths(self._h, start, length, _ptr(out)))
        return out

    def get_branch_length_differences(self):
        out = np.zeros(self.gpcsp_count)
        self._check(self._lib.bito_gp_get_branch_length_differences(self._h, _ptr(out)))
        return out

    def set_optimization_method(self, method):
        method = OPTIMIZATION_METHODS[method] if isinstance(method, str) else int(method)
        self._check(self._lib.bito_gp_set_optimization_method(self._h, method))

    def use_gradient_optimization(self, use_gradients):
        self._check(self._lib.bito_gp_use_gradient_optimization(self._h, int(bool(use_gradients))))

    def set_significant_digits_for_optimization(self, significant_digits):
        self._check(self._lib.bito_gp_set_significant_digits_for_optimization(self._h, int(significant_digits)))

    def get_optimization_count(self):
        return int(self._lib.bito_gp_get_optimization_count(self._h))

    def reset_optimization_count(self):
        self._check(self._lib.bito_gp_reset_optimization_count(self._h))

    def increment_optimization_count(self):
        self._check(self._lib.bito_gp_increment_optimization_count(self._h))

    def is_first_optimization(self):
        return self.get_optimization_count() == 0

    def log_likelihood_and_derivative(self, gpcsp, rootward, leafward):
        return self.log_likelihood_and_first_two_derivatives(gpcsp, rootward, leafward)[:2]

    def log_likelihood_and_first_two_derivatives(self, gpcsp, rootward, leafward):
        out = np.zeros(3)
        self._check(self._lib.bito_gp_log_likelihood_and_derivatives(self._h, int(gpcsp), int(rootward),
                                                                     int(leafward), _ptr(out)))
        return tuple(out)

    def set_substitution_model(self, eigenvectors, inverse_eigenvectors, eigenvalues, frequencies):
        """SubstitutionModel::GetEigenvectors / GetInverseEigenvectors (4x4) / GetEigenvalues / GetFrequencies of
        any reversible nucleotide model (substitution_model.hpp:24-30); the engine starts with JC69."""
        v = _f64(eigenvectors, 16, "eigenvectors")
        vinv = _f64(inverse_eigenvectors, 16, "inverse_eigenvectors")
        lam = _f64(eigenvalues, 4, "eigenvalues")
        pi = _f64(frequencies, 4, "frequencies")
        self._check(self._lib.bito_gp_set_substitution_model(self._h, _ptr(v), _ptr(vinv), _ptr(lam), _ptr(pi)))

    def get_transition_matrix(self, branch_length):
        out = np.zeros((4, 4))
        self._check(self._lib.bito_gp_get_transition_matrix(self._h, float(branch_length), _ptr(out)))
        return out

    # ---- read-back ---------------------------------------------------------------------------
    def get_log_marginal_likelihood(self):
        out = C.c_double()
        self._check(self._lib.bito_gp_get_log_marginal_likelihood(self._h, C.byref(out)))
        return out.value

    def get_per_gpcsp_log_likelihoods(self, start=0, length=None):
        length = self.gpcsp_count - start if length is None else length
        out = np.zeros(length)
        self._check(self._lib.bito_gp_get_per_gpcsp_log_likelihoods(self._h, start, length, _ptr(out)))
        return out

    def get_per_gpcsp_components_of_full_log_marginal(self):
        out = np.zeros(self.gpcsp_count)
        self._check(self._lib.bito_gp_get_per_gpcsp_components_of_full_log_marginal(self._h, _ptr(out)))
        return out

    def get_log_likelihood_matrix(self):
        out = np.zeros((self.gpcsp_count, self.pattern_count))
        self._check(self._lib.bito_gp_get_log_likelihood_matrix(self._h, _ptr(out)))
        return out

    def get_per_pattern_log_marginal(self):
        out = np.zeros(self.pattern_count)
        self._check(self._lib.bito_gp_get_per_pattern_log_marginal(self._h, _ptr(out)))
        return out

    def get_sbn_parameters(self):
        out = np.zeros(self.gpcsp_count)
        self._check(self._lib.bito_gp_get_sbn_parameters(self._h, _ptr(out)))
        return out

    def set_sbn_parameters(self, q):
        q = _f64(q, self.gpcsp_count, "q")
        self._check(self._lib.bito_gp_set_sbn_parameters(self._h, _ptr(q)))

    def get_plv(self, plv_id):
        out = np.zeros((self.pattern_count, 4))
        self._check(self._lib.bito_gp_get_plv(self._h, int(plv_id), _ptr(out)))
        return out

    def set_plv(self, plv_id, values, count=0):
        values = _f64(values, 4 * self.pattern_count, "plv")
        self._check(self._lib.bito_gp_set_plv(self._h, int(plv_id), _ptr(values), int(count)))

    def get_rescaling_counts(self):
        out = np.zeros(self.padded_plv_count, dtype=np.int32)
        self._check(self._lib.bito_gp_get_rescaling_counts(self._h, _ptr(out)))
        return out

    # ---- quartet hybrid marginals: gp_engine.cpp:748-816 --------------------------------------------
    def calculate_quartet_hybrid_likelihoods(self, central_gpcsp_idx, tip_counts, tips):
        """GPEngine::CalculateQuartetHybridLikelihoods. tip_counts = (rootward, sister, rotated,
        sorted); tips = rows (tip_node_id, plv_idx, gpcsp_idx) in that order."""
        counts = np.ascontiguousarray(tip_counts, dtype=np.int32).reshape(4)
        tips = np.ascontiguousarray(tips, dtype=np.int64).reshape(-1, 3)
        if tips.shape[0] != int(counts.sum()):
            raise ValueError("tips must hold sum(tip_counts) rows")
        out = np.zeros(int(np.prod(counts.astype(np.int64))))
        self._check(self._lib.bito_gp_calculate_quartet_hybrid_likelihoods(
            self._h, int(central_gpcsp_idx), _ptr(tips), _ptr(counts), _ptr(out)))
        return out

    def process_quartet_hybrid_requests(self, central_gpcsp_idx, tip_counts, tips):
        """GPEngine::ProcessQuartetHybridRequest for a batch of requests in one launch."""
        central = np.ascontiguousarray(central_gpcsp_idx, dtype=np.int64).reshape(-1)
        counts = np.ascontiguousarray(tip_counts, dtype=np.int32).reshape(-1, 4)
        tips = np.ascontiguousarray(tips, dtype=np.int64).reshape(-1, 3)
        if counts.shape[0] != central.size or tips.shape[0] != int(counts.sum()):
            raise ValueError("tip_counts / tips do not match the request count")
        self._check(self._lib.bito_gp_process_quartet_hybrid_requests(
            self._h, central.size, _ptr(central), _ptr(counts), _ptr(tips)))

    def get_hybrid_marginals(self):
        out = np.zeros(self.gpcsp_count)
        self._check(self._lib.bito_gp_get_hybrid_marginals(self._h, _ptr(out)))
        return out

    # ---- resize / copy -------------------------------------------------------------------------
    def grow_plvs(self, node_count, node_reindexer=None, explicit_allocation=None):
        r = None if node_reindexer is None else np.ascontiguousarray(node_reindexer, dtype=np.int64)
        self._check(self._lib.bito_gp_grow_plvs(self._h, int(node_count), _ptr(r),
                                                -1 if explicit_allocation is None else int(explicit_allocation)))

    def grow_gpcsps(self, gpcsp_count, gpcsp_reindexer=None, explicit_allocation=None):
        r = None if gpcsp_reindexer is None else np.ascontiguousarray(gpcsp_reindexer, dtype=np.int64)
        self._check(self._lib.bito_gp_grow_gpcsps(self._h, int(gpcsp_count), _ptr(r),
                                                  -1 if explicit_allocation is None else int(explicit_allocation)))

    def grow_spare_plvs(self, new_node_spare_count):
        self._check(self._lib.bito_gp_grow_spare_plvs(self._h, int(new_node_spare_count)))

    def grow_spare_gpcsps(self, new_gpcsp_spare_count):
        self._check(self._lib.bito_gp_grow_spare_gpcsps(self._h, int(new_gpcsp_spare_count)))

    def copy_node_data(self, src_node_idx, dest_node_idx):
        self._check(self._lib.bito_gp_copy_node_data(self._h, int(src_node_idx), int(dest_node_idx)))

    def copy_plv_data(self, src_plv_idx, dest_plv_idx):
        self._check(self._lib.bito_gp_copy_plv_data(self._h, int(src_plv_idx), int(dest_plv_idx)))

    def copy_gpcsp_data(self, src_gpcsp_idx, dest_gpcsp_idx):
        self._check(self._lib.bito_gp_copy_gpcsp_data(self._h, int(src_gpcsp_idx), int(dest_gpcsp_idx)))

    # ---- multi-GPU / streams / stats ---------------------------------------------------------------
    @staticmethod
    def make_unique_id() -> bytes:
        lib = _lib.load()
        buf = (C.c_uint8 * 128)()
        if lib.bito_gp_comm_make_unique_id(buf) != 0:
            raise RuntimeError(lib.bito_gp_last_error().decode())
        return bytes(buf)

    def comm_init(self, n_ranks, rank, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self._lib.bito_gp_comm_init(self._h, int(n_ranks), int(rank), buf))

    def set_stream(self, cuda_stream):
        self._check(self._lib.bito_gp_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self._lib.bito_gp_synchronize(self._h))

    def set_profiling(self, on: bool):
        self._check(self._lib.bito_gp_set_profiling(self._h, int(bool(on))))

    def reset_kernel_profile(self):
        self._check(self._lib.bito_gp_reset_kernel_profile(self._h))

    def kernel_profile(self) -> list:
        """Per-kernel device time (CUDA events around each launch while profiling is on)."""
        buf = (_lib.KernelProfile * 16)()
        n = C.c_int()
        self._check(self._lib.bito_gp_get_kernel_profile(self._h, buf, 16, C.byref(n)))
        return [dict(name=buf[i].name.decode(), launches=buf[i].launches, total_ms=buf[i].total_ms,
                     algorithmic_bytes=buf[i].algorithmic_bytes) for i in range(n.value)]

    def stats(self) -> dict:
        st = _lib.Stats()
        self._check(self._lib.bito_gp_get_stats(self._h, C.byref(st)))
        return {name: getattr(st, name) for name, _ in st._fields_}

    # ---- short aliases (the names the parity-test harness uses for every engine it drives) ----
    log_marginal_likelihood = get_log_marginal_likelihood
    per_gpcsp_log_likelihoods = get_per_gpcsp_log_likelihoods
    per_gpcsp_components_of_full_log_marginal = get_per_gpcsp_components_of_full_log_marginal
    log_likelihood_matrix = get_log_likelihood_matrix
    per_pattern_log_marginal = get_per_pattern_log_marginal
    sbn_parameters = get_sbn_parameters
    branch_lengths = get_branch_lengths
    branch_length_differences = get_branch_length_differences
    rescaling_counts = get_rescaling_counts
    hybrid_marginals = get_hybrid_marginals
    optimization_count = get_optimization_count
    transition_matrix = get_transition_matrix

    def log_likelihood_and_derivatives(self, gpcsp, rootward, leafward, two=False):
        out = self.log_likelihood_and_first_two_derivatives(gpcsp, rootward, leafward)
        return out if two else out[:2]
