"""Host-side planner: a subsplit DAG built from rooted trees and the GPOperation lists that
drive the engine over it. Mirrors the role of the reference's GPDAG
(/root/reference/src/gp_dag.{hpp,cpp}, subsplit_dag.cpp) for callers that do not have bito
itself in the loop (synthetic benchmarks, tests, the Python GPInstance).

What is the same as the reference: the PLV id convention (type * N + node, types P,
PHatRight, PHatLeft, RHat, RRight, RLeft; pv_handler.hpp:26-33), leaves first, children
before parents, rootsplits last, rootsplit edges first and the children of one
(parent, clade) contiguous, the "left" clade = the one holding the lowest
taxon (bitset.cpp:268-272, 326-331), and the op sequences of every list (gp_dag.cpp:30-411).
What differs: the order WITHIN those groups. The reference inherits it from libstdc++
unordered_map iteration; here it is sorted, hence reproducible anywhere. Results per
subsplit / PCSP are identical (tests/test_gp_dag.py maps them through the bitsets).
"""
from __future__ import annotations

import sys
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .gp_operation import GPOperationVector

# PLV types, pv_handler.hpp:26-33
P, PHAT_RIGHT, PHAT_LEFT, RHAT, RRIGHT, RLEFT = range(6)

Clade = int  # bit i set <=> taxon i in the clade


def _lowest_bit(x: int) -> int:
    return (x & -x).bit_length() - 1


def make_subsplit(a: Clade, b: Clade) -> Tuple[Clade, Clade]:
    """(left, right): left holds the lowest taxon id (Bitset::SubsplitFromUnorderedClades)."""
    assert a & b == 0 and a and b
    return (a, b) if _lowest_bit(a) < _lowest_bit(b) else (b, a)


class RootedTree:
    """A rooted bifurcating topology over taxa 0..n-1: `children[v] = (a, b)` for internal v;
    leaves are 0..n-1; `root` is the root's id. Branch lengths (optional) are per node."""

    def __init__(self, taxon_count: int, children: Dict[int, Tuple[int, int]], root: int,
                 branch_lengths: Optional[Dict[int, float]] = None):
        self.taxon_count = taxon_count
        self.children = dict(children)
        self.root = root
        self.branch_lengths = dict(branch_lengths or {})

    def postorder(self) -> List[int]:
        out, stack = [], [(self.root, False)]
        while stack:
            v, done = stack.pop()
            if done or v < self.taxon_count:
                out.append(v)
            else:
                stack.append((v, True))
                a, b = self.children[v]
                stack.append((b, False))
                stack.append((a, False))
        return out

    def clades(self) -> Dict[int, Clade]:
        cl: Dict[int, Clade] = {}
        for v in self.postorder():
            cl[v] = (1 << v) if v < self.taxon_count else cl[self.children[v][0]] | cl[self.children[v][1]]
        return cl

    def pcsps(self) -> List[Tuple[Tuple[Clade, Clade], Tuple[Clade, Clade]]]:
        """(parent subsplit, child subsplit) pairs, leaves as (clade, 0); the root's subsplit is
        reported with parent None."""
        cl = self.clades()
        sub = {v: (make_subsplit(cl[c[0]], cl[c[1]])) for v, c in self.children.items()}
        for v in range(self.taxon_count):
            sub[v] = (1 << v, 0)
        out = [(None, sub[self.root])]
        for v, (a, b) in self.children.items():
            out.append((sub[v], sub[a]))
            out.append((sub[v], sub[b]))
        return out


def parse_newick(text: str, taxon_names: Optional[Sequence[str]] = None) -> Tuple[List[RootedTree], List[str]]:
    """Parses rooted bifurcating Newick trees (one per ';'). Taxon ids follow `taxon_names`
    if given, else order of first appearance."""
    names: List[str] = list(taxon_names) if taxon_names is not None else []
    index = {n: i for i, n in enumerate(names)}
    fixed = taxon_names is not None
    parsed = []
    for chunk in text.split(";"):
        s = chunk.strip()
        if not s:
            continue
        if s.startswith("[&R]") or s.startswith("[&U]"):
            s = s[4:].strip()
        pos = 0
        nodes: List[dict] = []

        def parse_node() -> int:
            nonlocal pos
            node = {"children": [], "name": None, "bl": None}
            if s[pos] == "(":
                pos += 1
                while True:
                    node["children"].append(parse_node())
                    if s[pos] == ",":
                        pos += 1
                        continue
                    if s[pos] == ")":
                        pos += 1
                        break
                    raise ValueError(f"bad Newick near {s[pos:pos + 20]!r}")
            start = pos
            while pos < len(s) and s[pos] not in ",():;":
                pos += 1
            label = s[start:pos].strip().strip("'\"")
            if label:
                node["name"] = label
            if pos < len(s) and s[pos] == ":":
                pos += 1
                start = pos
                while pos < len(s) and s[pos] not in ",();":
                    pos += 1
                node["bl"] = float(s[start:pos])
            nodes.append(node)
            return len(nodes) - 1

        root = parse_node()
        parsed.append((nodes, root))
        for nd in nodes:
            if not nd["children"]:
                if nd["name"] is None:
                    raise ValueError("unnamed leaf in Newick")
                if nd["name"] not in index:
                    if fixed:
                        raise ValueError(f"unknown taxon {nd['name']!r}")
                    index[nd["name"]] = len(names)
                    names.append(nd["name"])
    n = len(names)
    trees = []
    for nodes, root in parsed:
        ids: Dict[int, int] = {}
        nxt = n
        children: Dict[int, Tuple[int, int]] = {}
        bls: Dict[int, float] = {}
        for k, nd in enumerate(nodes):  # post-order: children appear before parents
            if not nd["children"]:
                ids[k] = index[nd["name"]]
            else:
                if len(nd["children"]) != 2:
                    raise ValueError("Tree is not bifurcating (the GP path needs rooted bifurcating trees; "
                                     "cf. rooted_tree.cpp:151-152)")
                ids[k] = nxt
                nxt += 1
                children[ids[k]] = (ids[nd["children"][0]], ids[nd["children"][1]])
            if nd["bl"] is not None:
                bls[ids[k]] = nd["bl"]
        trees.append(RootedTree(n, children, ids[root], bls))
    return trees, names


class GPDAG:
    """Subsplit DAG + GP op-list planner."""

    def __init__(self, taxon_count: int, pcsps: Iterable[Tuple[Optional[Tuple[Clade, Clade]], Tuple[Clade, Clade]]]):
        n = taxon_count
        self.taxon_count = n
        subsplits = set()
        rootsplits = set()
        pairs = set()
        for parent, child in pcsps:
            if child[1] != 0:
                subsplits.add(child)
            if parent is None:
                rootsplits.add(child)
            else:
                subsplits.add(parent)
                pairs.add((parent, child))
        full = (1 << n) - 1
        for r in rootsplits:
            assert r[0] | r[1] == full, "rootsplit does not cover all taxa"
        internal = sorted(subsplits, key=lambda s: (bin(s[0] | s[1]).count("1"), s[0] | s[1], s[0]))
        self.subsplits: List[Tuple[Clade, Clade]] = [(1 << t, 0) for t in range(n)] + internal
        self.node_id: Dict[Tuple[Clade, Clade], int] = {s: i for i, s in enumerate(self.subsplits)}
        N = len(self.subsplits)
        self.node_count = N  # without the DAG root (which owns no PLVs, gp_instance.cpp:158-160)
        self.rootsplit_ids: List[int] = sorted(self.node_id[r] for r in rootsplits)
        # leafward[v][side], rootward[v][side]; side 1 = left ("rotated"), 0 = right ("sorted")
        self.leafward: List[List[List[int]]] = [[[], []] for _ in range(N)]
        self.rootward: List[List[List[int]]] = [[[], []] for _ in range(N)]
        for parent, child in pairs:
            p, c = self.node_id[parent], self.node_id[child]
            union = child[0] | child[1]
            if union == parent[0]:
                side = 1
            elif union == parent[1]:
                side = 0
            else:
                raise ValueError("child subsplit does not split a clade of its parent")
            self.leafward[p][side].append(c)
            self.rootward[c][side].append(p)
        for v in range(N):
            for side in (0, 1):
                self.leafward[v][side].sort()
                self.rootward[v][side].sort()
        for v in range(n, N):
            if not self.leafward[v][0] or not self.leafward[v][1]:
                raise ValueError("DAG is not tree-complete: a subsplit lacks children on one clade")
        # ---- edge ids: rootsplit edges, then (parent, clade) ranges of non-leaf children, then leaf edges
        self.edge_id: Dict[Tuple[int, int], int] = {}
        self.edge_parent: List[int] = []
        self.edge_child: List[int] = []
        self.edge_on_left: List[int] = []
        self.parent_to_child_range: Dict[Tuple[int, int], Tuple[int, int]] = {}
        ROOT = N  # the DAG root's node id
        self.dag_root_id = ROOT

        def add_edge(p, c, side):
            self.edge_id[(p, c)] = len(self.edge_parent)
            self.edge_parent.append(p)
            self.edge_child.append(c)
            self.edge_on_left.append(side)

        for r in self.rootsplit_ids:
            add_edge(ROOT, r, 1)
        self.rootsplit_count = len(self.rootsplit_ids)
        for v in range(n, N):
            for side in (1, 0):
                kids = [c for c in self.leafward[v][side] if c >= n]
                if kids:
                    start = len(self.edge_parent)
                    for c in kids:
                        add_edge(v, c, side)
                    self.parent_to_child_range[(v, side)] = (start, len(self.edge_parent))
        for leaf in range(n):
            for side in (0, 1):
                for p in self.rootward[leaf][side]:
                    start = len(self.edge_parent)
                    add_edge(p, leaf, side)
                    self.parent_to_child_range[(p, side)] = (start, start + 1)
        self.edge_count = len(self.edge_parent)
        self._topology_count_below: Optional[List[int]] = None

    # ---- construction helpers ------------------------------------------------------------
    @classmethod
    def from_trees(cls, trees: Sequence[RootedTree]) -> "GPDAG":
        pcsps = set()
        for t in trees:
            pcsps.update(t.pcsps())
        return cls(trees[0].taxon_count, pcsps)

    @classmethod
    def from_newick(cls, text: str, taxon_names: Optional[Sequence[str]] = None):
        trees, names = parse_newick(text, taxon_names)
        dag = cls.from_trees(trees)
        dag.taxon_names = names
        dag.trees = trees
        return dag

    # ---- indices ----------------------------------------------------------------------------
    def plv(self, plv_type: int, node: int) -> int:
        """PLVNodeHandler::GetPVIndex (pv_handler.hpp:487-490)."""
        return plv_type * self.node_count + node

    def r_plv(self, on_left: int, node: int) -> int:
        """The r-PLV facing a child on the given clade (pv_handler.hpp:41-49)."""
        return self.plv(RLEFT if on_left else RRIGHT, node)

    def node_bitset(self, v: int) -> str:
        """Reference-style '0'/'1' string, left clade then right clade."""
        n = self.taxon_count
        if v == self.dag_root_id:
            return "1" * n + "0" * n
        left, right = self.subsplits[v]
        fmt = lambda c: "".join("1" if c >> t & 1 else "0" for t in range(n))  # noqa: E731
        return fmt(left) + fmt(right)

    def pcsp_key(self, edge: int) -> Tuple[str, str]:
        return self.node_bitset(self.edge_parent[edge]), self.node_bitset(self.edge_child[edge])

    # ---- priors: subsplit_dag.cpp:644-664, 987-1007, 1025-1046 ---------------------------------
    def topology_count_below(self) -> List[int]:
        if self._topology_count_below is None:
            cnt = [1] * self.node_count
            for v in range(self.taxon_count, self.node_count):
                cnt[v] = sum(cnt[c] for c in self.leafward[v][1]) * sum(cnt[c] for c in self.leafward[v][0])
            self._topology_count_below = cnt
        return self._topology_count_below

    def topology_count(self) -> int:
        cnt = self.topology_count_below()
        return sum(cnt[r] for r in self.rootsplit_ids)

    def build_uniform_on_topological_support_prior(self) -> np.ndarray:
        cnt = self.topology_count_below()
        q = np.ones(self.edge_count)
        total = float(sum(cnt[r] for r in self.rootsplit_ids))
        for r in self.rootsplit_ids:
            q[self.edge_id[(self.dag_root_id, r)]] = cnt[r] / total
        for v in range(self.taxon_count, self.node_count):
            for side in (0, 1):
                kids = self.leafward[v][side]
                tot = float(sum(cnt[c] for c in kids))
                for c in kids:
                    q[self.edge_id[(v, c)]] = cnt[c] / tot
        return q

    def unconditional_node_probabilities(self, q: np.ndarray) -> np.ndarray:
        """Indexed by node id without the DAG root (the slice MakeGPEngine passes on)."""
        prob = np.zeros(self.node_count)
        for r in self.rootsplit_ids:
            prob[r] += q[self.edge_id[(self.dag_root_id, r)]]
        for v in range(self.node_count - 1, self.taxon_count - 1, -1):  # parents before children
            for side in (0, 1):
                for c in self.leafward[v][side]:
                    prob[c] += prob[v] * q[self.edge_id[(v, c)]]
        return prob

    def inverted_gpcsp_probabilities(self, q: np.ndarray, node_prob: np.ndarray) -> np.ndarray:
        inv = np.ones(self.edge_count)
        for (p, c), e in self.edge_id.items():
            if p != self.dag_root_id:
                inv[e] = node_prob[p] * q[e] / node_prob[c]
        return inv

    # ---- op lists: gp_dag.cpp -----------------------------------------------------------------
    def rootward_order(self) -> List[int]:
        return list(range(self.taxon_count, self.node_count))

    def leafward_order(self) -> List[int]:
        return list(range(self.node_count - 1, self.taxon_count - 1, -1)) + list(range(self.taxon_count))

    def set_rootward_zero(self, ops: GPOperationVector):  # :249-258
        for v in range(self.taxon_count, self.node_count):
            ops.zero_plv(self.plv(P, v))
            ops.zero_plv(self.plv(PHAT_RIGHT, v))
            ops.zero_plv(self.plv(PHAT_LEFT, v))

    def set_leafward_zero(self, ops: GPOperationVector):  # :229-237
        for v in range(self.node_count):
            ops.zero_plv(self.plv(RHAT, v))
            ops.zero_plv(self.plv(RRIGHT, v))
            ops.zero_plv(self.plv(RLEFT, v))

    def set_rhat_to_stationary(self, ops: GPOperationVector):  # :239-247
        for r in self.rootsplit_ids:
            ops.set_to_stationary_distribution(self.plv(RHAT, r), self.edge_id[(self.dag_root_id, r)])

    def _add_phat(self, ops, v, on_left):  # :317-330
        dest = self.plv(PHAT_LEFT if on_left else PHAT_RIGHT, v)
        ops.append_after_prep_for_marginalization(
            [(dest, self.edge_id[(v, c)], self.plv(P, c)) for c in self.leafward[v][on_left]])

    def _add_rhat(self, ops, v):  # :332-344
        if v in self._rootsplit_set():
            return
        dest = self.plv(RHAT, v)
        ops.append_after_prep_for_marginalization(
            [(dest, self.edge_id[(p, v)], self.r_plv(side, p)) for side in (0, 1) for p in self.rootward[v][side]])

    def _rootsplit_set(self):
        if not hasattr(self, "_rs"):
            self._rs = set(self.rootsplit_ids)
        return self._rs

    def rootward_pass(self, ops: Optional[GPOperationVector] = None) -> GPOperationVector:  # :278-294
        ops = ops if ops is not None else GPOperationVector()
        for v in self.rootward_order():
            self._add_phat(ops, v, 0)
            self._add_phat(ops, v, 1)
            ops.multiply(self.plv(P, v), self.plv(PHAT_RIGHT, v), self.plv(PHAT_LEFT, v))
        return ops

    def leafward_pass(self, ops: Optional[GPOperationVector] = None) -> GPOperationVector:  # :260-276
        ops = ops if ops is not None else GPOperationVector()
        for v in self.leafward_order():
            self._add_rhat(ops, v)
            ops.multiply(self.plv(RRIGHT, v), self.plv(RHAT, v), self.plv(PHAT_LEFT, v))
            ops.multiply(self.plv(RLEFT, v), self.plv(RHAT, v), self.plv(PHAT_RIGHT, v))
        return ops

    def populate_plvs(self) -> GPOperationVector:  # :296-304
        ops = GPOperationVector()
        self.set_rootward_zero(ops)
        self.set_leafward_zero(ops)
        self.set_rhat_to_stationary(ops)
        self.rootward_pass(ops)
        self.leafward_pass(ops)
        return ops

    def marginal_likelihood(self, ops: Optional[GPOperationVector] = None) -> GPOperationVector:  # :202-211
        ops = ops if ops is not None else GPOperationVector()
        ops.reset_marginal_likelihood()
        for r in self.rootsplit_ids:
            ops.increment_marginal_likelihood(self.plv(RHAT, r), self.edge_id[(self.dag_root_id, r)],
                                              self.plv(P, r))
        return ops

    def compute_likelihoods(self) -> GPOperationVector:  # :177-196
        ops = GPOperationVector()
        for v in range(self.taxon_count, self.node_count):
            for side in (0, 1):
                for c in self.leafward[v][side]:
                    ops.likelihood(self.edge_id[(v, c)], self.plv(P, c), self.r_plv(side, v))
        return self.marginal_likelihood(ops)

    def optimize_sbn_parameters(self) -> GPOperationVector:  # :217-227, 346-354
        ops = GPOperationVector()
        for v in self.leafward_order():
            # Both orientations of the subsplit; every non-empty range gets an op, length-1 ranges
            # included (the reference's unsigned wrap-around at gp_dag.cpp:350).
            for side in (0, 1):
                rng = self.parent_to_child_range.get((v, side))
                if rng is not None:
                    ops.update_sbn_probabilities(*rng)
        ops.update_sbn_probabilities(0, self.rootsplit_count)
        return ops

    def batched_branch_length_optimization(self) -> GPOperationVector:
        """Jacobi schedule: every non-rootsplit edge optimised against the CURRENT PLVs in one
        dependency level. A valid GPOperationVector, so the reference engine can run it too
        (hard part 1b). Follow with populate_plvs()."""
        ops = GPOperationVector()
        for v in range(self.taxon_count, self.node_count):
            for side in (0, 1):
                for c in self.leafward[v][side]:
                    ops.optimize_branch_length(self.plv(P, c), self.r_plv(side, v), self.edge_id[(v, c)])
        return ops

    # ---- the reference's Gauss-Seidel sweep: gp_dag.cpp:78-121 over tidy_subsplit_dag.hpp:81-172 ----
    def branch_length_optimization(self) -> GPOperationVector:
        ops = GPOperationVector()
        N = self.node_count
        # above[v]: bitmask over (node, side) pairs strictly above v (bit 2*node + side)
        above = [0] * N
        for v in range(N - 1, -1, -1):
            m = 0
            for side in (0, 1):
                for p in self.rootward[v][side]:
                    m |= (1 << (2 * p + side)) | above[p]
            above[v] = m
        state = {"dirty": 0, "updating": None}
        visited = set()
        is_leaf = lambda v: v < self.taxon_count  # noqa: E731
        rootsplits = self._rootsplit_set()
        old_limit = sys.getrecursionlimit()
        sys.setrecursionlimit(max(old_limit, 20 * self.taxon_count + 1000))

        def dirty(v, side):
            return state["dirty"] >> (2 * v + side) & 1

        def set_clean(v, side):
            state["dirty"] &= ~(1 << (2 * v + side))

        def before_node(v):
            if v not in rootsplits:
                self._update_rhat(ops, v)

        def after_node(v):
            ops.multiply(self.plv(P, v), self.plv(PHAT_RIGHT, v), self.plv(PHAT_LEFT, v))

        def before_node_clade(v, side):
            ops.multiply(self.r_plv(side, v), self.plv(RHAT, v), self.plv(PHAT_RIGHT if side else PHAT_LEFT, v))
            ops.zero_plv(self.plv(PHAT_LEFT if side else PHAT_RIGHT, v))

        def modify_edge(v, c, side):  # OptimizeBranchLengthUpdatePHat, :392-411
            e = self.edge_id[(v, c)]
            ops.optimize_branch_length(self.plv(P, c), self.r_plv(side, v), e)
            ops.append_after_prep_for_marginalization(
                [(self.plv(PHAT_LEFT if side else PHAT_RIGHT, v), e, self.plv(P, c))])

        def update_edge(v, c, side):  # UpdatePHatComputeLikelihood, :375-390
            e = self.edge_id[(v, c)]
            dest = self.plv(PHAT_LEFT if side else PHAT_RIGHT, v)
            ops.prep_for_marginalization(dest, [self.plv(P, c)])
            ops.increment_with_weighted_evolved_plv(dest, e, self.plv(P, c))
            ops.likelihood(e, self.plv(P, c), self.r_plv(side, v))

        def for_node(v):
            before_node(v)
            for_node_clade(v, 1)
            for_node_clade(v, 0)
            after_node(v)

        def for_node_clade(v, side):
            if state["updating"] is not None:
                update_clade(v, side)
            else:
                modify_clade(v, side)

        def update_clade(v, side):
            if dirty(v, side):
                for c in self.leafward[v][side]:
                    if not is_leaf(c):
                        for_node_clade(c, 1)
                        for_node_clade(c, 0)
                        after_node(c)
                    update_edge(v, c, side)
                    set_clean(v, side)
            if state["updating"] == (v, side):
                state["updating"] = None

        def modify_clade(v, side):
            if dirty(v, 1 - side):
                state["updating"] = (v, 1 - side)
                update_clade(v, 1 - side)
            before_node_clade(v, side)
            for c in self.leafward[v][side]:
                if c not in visited:
                    visited.add(c)
                    if not is_leaf(c):
                        for_node(c)
                modify_edge(v, c, side)
                state["dirty"] |= above[v]
                set_clean(v, side)

        try:
            for r in self.rootsplit_ids:
                for_node(r)
        finally:
            sys.setrecursionlimit(old_limit)
        return ops

    def _update_rhat(self, ops, v):  # :356-369
        ops.zero_plv(self.plv(RHAT, v))
        dest = self.plv(RHAT, v)
        ops.append_after_prep_for_marginalization(
            [(dest, self.edge_id[(p, v)], self.r_plv(side, p)) for side in (0, 1) for p in self.rootward[v][side]])

    # ---- summary ---------------------------------------------------------------------------------
    def summary(self) -> dict:
        return dict(taxa=self.taxon_count, nodes=self.node_count, edges=self.edge_count,
                    rootsplits=self.rootsplit_count, topologies=float(self.topology_count()))
