"""GPOperations, mirroring /root/reference/src/gp_operation.hpp:24-170.

An operation list is carried as an int64 array of shape (n, 6) = (kind, a, b, c, vec_off,
vec_len) — `bito_gp_op` in include/bito_gp.h, same field order as the reference structs —
plus one shared int64 pool holding the PrepForMarginalization source vectors.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np

ZERO_PLV = 0
SET_TO_STATIONARY_DISTRIBUTION = 1
INCREMENT_WITH_WEIGHTED_EVOLVED_PLV = 2
MULTIPLY = 3
LIKELIHOOD = 4
OPTIMIZE_BRANCH_LENGTH = 5
UPDATE_SBN_PROBABILITIES = 6
RESET_MARGINAL_LIKELIHOOD = 7
INCREMENT_MARGINAL_LIKELIHOOD = 8
PREP_FOR_MARGINALIZATION = 9

KIND_NAMES = (
    "ZeroPLV", "SetToStationaryDistribution", "IncrementWithWeightedEvolvedPLV", "Multiply",
    "Likelihood", "OptimizeBranchLength", "UpdateSBNProbabilities", "ResetMarginalLikelihood",
    "IncrementMarginalLikelihood", "PrepForMarginalization")


class GPOperationVector:
    """Builder for a flattened GPOperationVector (gp_operation.hpp:170)."""

    def __init__(self):
        self._rows: List[Tuple[int, int, int, int, int, int]] = []
        self._vec: List[int] = []

    def __len__(self):
        return len(self._rows)

    # constructors, argument order as in the reference structs
    def zero_plv(self, dest):
        self._rows.append((ZERO_PLV, dest, 0, 0, 0, 0))

    def set_to_stationary_distribution(self, dest, root_gpcsp_idx):
        self._rows.append((SET_TO_STATIONARY_DISTRIBUTION, dest, root_gpcsp_idx, 0, 0, 0))

    def increment_with_weighted_evolved_plv(self, dest, gpcsp, src):
        self._rows.append((INCREMENT_WITH_WEIGHTED_EVOLVED_PLV, dest, gpcsp, src, 0, 0))

    def multiply(self, dest, src1, src2):
        self._rows.append((MULTIPLY, dest, src1, src2, 0, 0))

    def likelihood(self, dest, child, parent):
        self._rows.append((LIKELIHOOD, dest, child, parent, 0, 0))

    def optimize_branch_length(self, leafward, rootward, gpcsp):
        self._rows.append((OPTIMIZE_BRANCH_LENGTH, leafward, rootward, gpcsp, 0, 0))

    def update_sbn_probabilities(self, start, stop):
        self._rows.append((UPDATE_SBN_PROBABILITIES, start, stop, 0, 0, 0))

    def reset_marginal_likelihood(self):
        self._rows.append((RESET_MARGINAL_LIKELIHOOD, 0, 0, 0, 0, 0))

    def increment_marginal_likelihood(self, stationary_times_prior, rootsplit, p):
        self._rows.append((INCREMENT_MARGINAL_LIKELIHOOD, stationary_times_prior, rootsplit, p, 0, 0))

    def prep_for_marginalization(self, dest, src_vector: Sequence[int]):
        off = len(self._vec)
        self._vec.extend(int(s) for s in src_vector)
        self._rows.append((PREP_FOR_MARGINALIZATION, dest, 0, 0, off, len(src_vector)))

    def append_after_prep_for_marginalization(self, increments: Iterable[Tuple[int, int, int]]):
        """AppendOperationsAfterPrepForMarginalization, gp_dag.cpp:309-315: `increments` are
        (dest, gpcsp, src) triples into ONE dest; a Prep over their sources goes first."""
        increments = list(increments)
        if not increments:
            return
        dest = increments[0][0]
        assert all(d == dest for d, _, _ in increments), "dest_ mismatch in PrepForMarginalizationVisitor"
        self.prep_for_marginalization(dest, [s for _, _, s in increments])
        for d, g, s in increments:
            self.increment_with_weighted_evolved_plv(d, g, s)

    def extend(self, other: "GPOperationVector"):
        shift = len(self._vec)
        self._vec.extend(other._vec)
        for k, a, b, c, off, ln in other._rows:
            self._rows.append((k, a, b, c, off + shift if k == PREP_FOR_MARGINALIZATION else off, ln))

    def arrays(self):
        ops = np.asarray(self._rows, dtype=np.int64).reshape(-1, 6)
        vec = np.asarray(self._vec, dtype=np.int64)
        return ops, vec


def as_arrays(ops, vec=None):
    """Normalises (ops, vec) to contiguous int64 arrays."""
    if isinstance(ops, GPOperationVector):
        return ops.arrays()
    ops = np.ascontiguousarray(ops, dtype=np.int64).reshape(-1, 6)
    vec = np.ascontiguousarray(vec if vec is not None else np.zeros(0), dtype=np.int64)
    return ops, vec


def to_strings(ops, vec=None):
    """Human-readable dump in the spirit of GPOperationOstream (gp_operation.hpp:218-254)."""
    ops, vec = as_arrays(ops, vec)
    out = []
    for k, a, b, c, off, ln in ops.tolist():
        if k == PREP_FOR_MARGINALIZATION:
            out.append(f"PrepForMarginalization[dest_={a}, src_vector_={vec[off:off + ln].tolist()}]")
        else:
            out.append(f"{KIND_NAMES[k]}[{a}, {b}, {c}]")
    return out
