"""CPU tests (no GPU): the C-ABI shared library builds for sm_100a, loads, and exports exactly
the symbols include/bito_gp.h declares. No compute entry point is called."""
import ctypes
import os
import re
import subprocess

import pytest

from bito_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "bito_gp.h")).read()
    return re.findall(r"^BITO_GP_API\s+[\w\s\*]+?\b(bito_gp_\w+)\s*\(", text, flags=re.M)


@pytest.fixture(scope="module")
def library():
    build.build()
    return _lib.load()


def test_header_declares_every_bound_function(library):
    declared = header_functions()
    assert len(declared) == len(set(declared)) >= 45
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(library):
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (bito_gp_\w+)", out))
    assert set(header_functions()) <= exported
    # nothing but the C-ABI leaks out of the library
    leaked = [ln for ln in out.splitlines() if " T " in ln and "bito_gp_" not in ln]
    assert not leaked, leaked[:5]


def test_abi_version_and_struct_sizes(library):
    assert library.bito_gp_abi_version() == _lib.ABI_VERSION == 1
    # bito_gp_op is six int64 (include/bito_gp.h); the Python side passes int64[n][6]
    assert ctypes.sizeof(_lib.Config) == 80
    assert ctypes.sizeof(_lib.Stats) == 168  # 21 x 8: objective_passes was appended in round 2


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_create_without_gpu_fails_loudly(library):
    """No CPU fallback: with no CUDA device the constructor must raise, not degrade."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    import numpy as np
    from bito_b200.gp_engine import GPEngine
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        GPEngine(np.zeros((3, 4), dtype=np.uint8), np.ones(4), 4, 5, 5)


def test_product_path_never_touches_the_oracle():
    """bito_b200/ must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "bito_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/_ref", ""), os.path.join(dirpath, f)
