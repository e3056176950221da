"""Loads bito_b200/libbito_gp_b200.so and declares the C-ABI of include/bito_gp.h for ctypes.

There is deliberately no fallback: if the CUDA library is missing or does not export every
symbol the header declares, importing the engine fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbito_gp_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "bito_gp.h")

ABI_VERSION = 1

FLAG_NO_CUDA_GRAPHS = 1
FLAG_NO_FUSION = 2
FLAG_NO_LOGLIK_MATRIX = 4
FLAG_STRICT_ASSERTS = 8
FLAG_NO_ONCHIP_OPTIMIZER = 16


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("device", C.c_int32),
        ("taxon_count", C.c_int64),
        ("pattern_count", C.c_int64),
        ("site_count", C.c_int64),
        ("node_count", C.c_int64),
        ("gpcsp_count", C.c_int64),
        ("rescaling_threshold", C.c_double),
        ("use_gradients", C.c_int32),
        ("spare_node_count", C.c_int32),
        ("spare_gpcsp_count", C.c_int32),
        ("flags", C.c_int32),
        ("max_device_bytes", C.c_int64),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_int64),
        ("graph_launches", C.c_int64),
        ("process_calls", C.c_int64),
        ("programs_compiled", C.c_int64),
        ("levels_last", C.c_int64),
        ("fused_ops_last", C.c_int64),
        ("objective_evaluations", C.c_int64),
        ("collective_calls", C.c_int64),
        ("device_bytes_in_use", C.c_int64),
        ("plvs_resident", C.c_int64),
        ("algorithmic_bytes_last", C.c_double),
        ("last_process_ms", C.c_double),
        ("device_status_bits", C.c_int64),
        ("optimizer_scheme", C.c_int64),
        ("optimizer_cluster_size", C.c_int64),
        ("optimizer_cluster_threads", C.c_int64),
        ("optimizer_edges_in_flight", C.c_int64),
        ("peer_collective_calls", C.c_int64),
        ("programs_evicted", C.c_int64),
        ("programs_cached", C.c_int64),
        ("objective_passes", C.c_int64),
    ]


class KernelProfile(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int64), ("total_ms", C.c_double),
                ("algorithmic_bytes", C.c_double)]


_vp, _i64, _i32, _f64 = C.c_void_p, C.c_int64, C.c_int32, C.c_double
_int = C.c_int

# name -> (restype, argtypes). Must list every BITO_GP_API function of include/bito_gp.h
# (tests/test_abi.py checks this table against the header and the .so).
SIGNATURES = {
    "bito_gp_last_error": (C.c_char_p, []),
    "bito_gp_abi_version": (_int, []),
    "bito_gp_create": (_int, [C.POINTER(Config), C.POINTER(_vp)]),
    "bito_gp_destroy": (None, [_vp]),
    "bito_gp_set_site_patterns": (_int, [_vp, _vp, _vp]),
    "bito_gp_set_site_patterns_device": (_int, [_vp, _vp, _vp]),
    "bito_gp_initialize_priors": (_int, [_vp, _vp, _vp, _vp]),
    "bito_gp_set_null_prior": (_int, [_vp]),
    "bito_gp_process_operations": (_int, [_vp, _vp, _i64, _vp, _i64]),
    "bito_gp_set_branch_lengths": (_int, [_vp, _vp]),
    "bito_gp_set_branch_lengths_range": (_int, [_vp, _i64, _i64, _vp]),
    "bito_gp_set_branch_lengths_to_constant": (_int, [_vp, _f64]),
    "bito_gp_set_branch_lengths_to_default": (_int, [_vp]),
    "bito_gp_get_branch_lengths": (_int, [_vp, _i64, _i64, _vp]),
    "bito_gp_get_branch_length_differences": (_int, [_vp, _vp]),
    "bito_gp_set_optimization_method": (_int, [_vp, _int]),
    "bito_gp_use_gradient_optimization": (_int, [_vp, _int]),
    "bito_gp_set_significant_digits_for_optimization": (_int, [_vp, _int]),
    "bito_gp_get_optimization_count": (_i64, [_vp]),
    "bito_gp_reset_optimization_count": (_int, [_vp]),
    "bito_gp_increment_optimization_count": (_int, [_vp]),
    "bito_gp_log_likelihood_and_derivatives": (_int, [_vp, _i64, _i64, _i64, _vp]),
    "bito_gp_get_transition_matrix": (_int, [_vp, _f64, _vp]),
    "bito_gp_set_substitution_model": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "bito_gp_get_log_marginal_likelihood": (_int, [_vp, C.POINTER(_f64)]),
    "bito_gp_get_per_gpcsp_log_likelihoods": (_int, [_vp, _i64, _i64, _vp]),
    "bito_gp_get_per_gpcsp_components_of_full_log_marginal": (_int, [_vp, _vp]),
    "bito_gp_get_log_likelihood_matrix": (_int, [_vp, _vp]),
    "bito_gp_get_per_pattern_log_marginal": (_int, [_vp, _vp]),
    "bito_gp_get_sbn_parameters": (_int, [_vp, _vp]),
    "bito_gp_set_sbn_parameters": (_int, [_vp, _vp]),
    "bito_gp_get_plv": (_int, [_vp, _i64, _vp]),
    "bito_gp_set_plv": (_int, [_vp, _i64, _vp, _i32]),
    "bito_gp_calculate_quartet_hybrid_likelihoods": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "bito_gp_process_quartet_hybrid_requests": (_int, [_vp, _i64, _vp, _vp, _vp]),
    "bito_gp_get_hybrid_marginals": (_int, [_vp, _vp]),
    "bito_gp_get_rescaling_counts": (_int, [_vp, _vp]),
    "bito_gp_get_node_count": (_i64, [_vp]),
    "bito_gp_get_plv_count": (_i64, [_vp]),
    "bito_gp_get_padded_plv_count": (_i64, [_vp]),
    "bito_gp_get_gpcsp_count": (_i64, [_vp]),
    "bito_gp_get_padded_gpcsp_count": (_i64, [_vp]),
    "bito_gp_get_site_pattern_count": (_i64, [_vp]),
    "bito_gp_grow_plvs": (_int, [_vp, _i64, _vp, _i64]),
    "bito_gp_grow_gpcsps": (_int, [_vp, _i64, _vp, _i64]),
    "bito_gp_grow_spare_plvs": (_int, [_vp, _i64]),
    "bito_gp_grow_spare_gpcsps": (_int, [_vp, _i64]),
    "bito_gp_copy_node_data": (_int, [_vp, _i64, _i64]),
    "bito_gp_copy_plv_data": (_int, [_vp, _i64, _i64]),
    "bito_gp_copy_gpcsp_data": (_int, [_vp, _i64, _i64]),
    "bito_gp_comm_make_unique_id": (_int, [_vp]),
    "bito_gp_comm_init": (_int, [_vp, _i32, _i32, _vp]),
    "bito_gp_set_stream": (_int, [_vp, _vp]),
    "bito_gp_synchronize": (_int, [_vp]),
    "bito_gp_get_stats": (_int, [_vp, C.POINTER(Stats)]),
    "bito_gp_set_profiling": (_int, [_vp, _int]),
    "bito_gp_reset_kernel_profile": (_int, [_vp]),
    "bito_gp_get_kernel_profile": (_int, [_vp, C.POINTER(KernelProfile), _int, C.POINTER(_int)]),
}

_lib = None


def load():
    """Returns the loaded C-ABI library; raises if it is missing or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python -m bito_b200.build` (needs nvcc); "
            "bito_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise ImportError(f"{LIB_PATH} does not export {name}; rebuild it") from exc
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.bito_gp_abi_version() != ABI_VERSION:
        raise ImportError("libbito_gp_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib
