"""Synthetic GP workloads of the shapes BASELINE.json names:
a seed tree, JC69-simulated site columns, and a subsplit DAG from trees that are a few NNI
moves away from the seed (heavy subsplit sharing, so the DAG stays HBM-sized)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .gp_dag import GPDAG, RootedTree, make_subsplit


def random_tree(taxon_count: int, rng: np.random.Generator, mean_branch_length: float) -> RootedTree:
    """Random rooted bifurcating topology (uniformly random joins), Exp(mean) branch lengths."""
    avail = list(range(taxon_count))
    children: Dict[int, Tuple[int, int]] = {}
    nxt = taxon_count
    while len(avail) > 1:
        i, j = rng.choice(len(avail), size=2, replace=False)
        a, b = avail[i], avail[j]
        for k in sorted((int(i), int(j)), reverse=True):
            avail.pop(k)
        children[nxt] = (a, b)
        avail.append(nxt)
        nxt += 1
    root = nxt - 1
    bl = {v: float(rng.exponential(mean_branch_length)) for v in range(nxt) if v != root}
    return RootedTree(taxon_count, children, root, bl)


def simulate_alignment(tree: RootedTree, site_count: int, rng: np.random.Generator,
                       gap_rate: float = 0.01) -> np.ndarray:
    """JC69 evolution down `tree`; returns symbols taxa x sites (0..3, 4 = gap)."""
    n = tree.taxon_count
    state: Dict[int, np.ndarray] = {tree.root: rng.integers(0, 4, size=site_count, dtype=np.uint8)}
    out = np.zeros((n, site_count), dtype=np.uint8)
    for v in reversed(tree.postorder()):  # parents before children
        if v >= n:
            for c in tree.children[v]:
                t = tree.branch_lengths[c]
                p_change = 0.75 - 0.75 * np.exp(-4.0 * t / 3.0)
                change = rng.random(site_count) < p_change
                shift = rng.integers(1, 4, size=site_count, dtype=np.uint8)
                state[c] = np.where(change, (state[v] + shift) % 4, state[v]).astype(np.uint8)
            del state[v]
        else:
            out[v] = state.pop(v)
    if gap_rate > 0:
        out[rng.random(out.shape) < gap_rate] = 4
    return out


def distinct_patterns(symbols: np.ndarray, weights: np.ndarray, refill) -> Tuple[np.ndarray, np.ndarray]:
    """Site-pattern compression as SitePattern::Compress does it (site_pattern.cpp:67-115): identical
    columns become ONE pattern whose weight is the sum of theirs. `refill(n)` draws n more columns (and
    weights) so that the caller still gets the pattern count it asked for; the result has no duplicate
    column. The first occurrence keeps its place, so a shard without duplicates is returned unchanged."""
    want = symbols.shape[1]
    for _ in range(64):
        cols = np.ascontiguousarray(symbols.T)
        key = cols.view(np.dtype((np.void, cols.shape[1]))).ravel()
        _, first, inverse = np.unique(key, return_index=True, return_inverse=True)
        if first.size == symbols.shape[1] and symbols.shape[1] == want:
            return symbols, weights
        merged = np.zeros(first.size)
        np.add.at(merged, inverse.ravel(), weights)
        order = np.argsort(first)  # keep first-occurrence order
        symbols, weights = symbols[:, first[order]], merged[order]
        if symbols.shape[1] < want:
            more_s, more_w = refill(want - symbols.shape[1])
            symbols, weights = np.concatenate([symbols, more_s], axis=1), np.concatenate([weights, more_w])
    raise RuntimeError("distinct_patterns: could not reach the requested number of distinct columns")


class _MutableTree:
    def __init__(self, tree: RootedTree):
        self.n = tree.taxon_count
        self.children = {v: tuple(c) for v, c in tree.children.items()}
        self.root = tree.root
        self.parent = {c: v for v, cs in self.children.items() for c in cs}
        self.clade = tree.clades()

    def subsplit(self, v):
        if v < self.n:
            return (1 << v, 0)
        a, b = self.children[v]
        return make_subsplit(self.clade[a], self.clade[b])

    def nni(self, v: int, which: int):
        """Swap child `which` of internal non-root v with v's sibling. Returns the undo token."""
        u = self.parent[v]
        s = self.children[u][0] if self.children[u][1] == v else self.children[u][1]
        c = self.children[v][which]
        other = self.children[v][1 - which]
        self.children[v] = (other, s)
        self.children[u] = (v, c)
        self.parent[s], self.parent[c] = v, u
        self.clade[v] = self.clade[other] | self.clade[s]
        return (v, s)

    def undo(self, token):
        v, s = token  # swap the former sibling back out of v
        self.nni(v, self.children[v].index(s))

    def local_pcsps(self, v: int):
        u = self.parent[v]
        out = []
        pu = self.parent.get(u)
        out.append((None if pu is None else self.subsplit(pu), self.subsplit(u)))
        for c in self.children[u]:
            out.append((self.subsplit(u), self.subsplit(c)))
        for c in self.children[v]:
            out.append((self.subsplit(v), self.subsplit(c)))
        return out


def nni_neighbourhood_dag(seed: RootedTree, tree_count: int, moves_per_tree: int,
                          rng: np.random.Generator, walk: bool = False) -> GPDAG:
    """DAG of the seed tree plus `tree_count - 1` more trees. walk=False: each is
    `moves_per_tree` random NNI moves from the SEED (a star around it); walk=True: each is
    `moves_per_tree` moves from the PREVIOUS tree (a random walk, so the DAG keeps growing).
    Every intermediate tree is included, so the DAG stays tree-complete."""
    pcsps = set(seed.pcsps())
    mt = _MutableTree(seed)
    internal = [v for v in seed.children if v != seed.root]
    for _ in range(max(0, tree_count - 1)):
        undo = []
        for _m in range(moves_per_tree):
            v = int(rng.choice(internal))
            undo.append(mt.nni(v, int(rng.integers(0, 2))))
            pcsps.update(mt.local_pcsps(v))
        if not walk:
            for token in reversed(undo):
                mt.undo(token)
    return GPDAG(seed.taxon_count, pcsps)


def nni_walk_trees(seed: RootedTree, tree_count: int, moves_per_tree: int,
                   rng: np.random.Generator) -> List[RootedTree]:
    """The seed tree plus `tree_count - 1` trees of an NNI random walk from it, as explicit
    trees (for callers that need a tree file, e.g. the reference's own newick parser)."""
    mt = _MutableTree(seed)
    internal = [v for v in seed.children if v != seed.root]
    trees = [seed]
    for _ in range(max(0, tree_count - 1)):
        for _m in range(moves_per_tree):
            mt.nni(int(rng.choice(internal)), int(rng.integers(0, 2)))
        bl = {v: float(rng.exponential(0.08)) + 1e-3 for v in list(mt.children) + list(range(mt.n))
              if v != mt.root}
        trees.append(RootedTree(mt.n, dict(mt.children), mt.root, bl))
    return trees


def tree_to_newick(tree: RootedTree, taxon_names: Sequence[str]) -> str:
    def rec(v):
        label = taxon_names[v] if v < tree.taxon_count else "(" + ",".join(rec(c) for c in tree.children[v]) + ")"
        return label + (f":{tree.branch_lengths[v]:.6f}" if v in tree.branch_lengths else "")
    return rec(tree.root) + ";"


def write_fasta(path: str, symbols: np.ndarray, taxon_names: Sequence[str]) -> None:
    """symbols: taxa x sites of 0..3 (ACGT) / 4 (gap), the engine's symbol table
    (site_pattern.cpp: SitePattern::GetSymbolTable)."""
    alphabet = np.frombuffer(b"ACGT-", dtype=np.uint8)
    with open(path, "w") as f:
        for name, row in zip(taxon_names, symbols):
            f.write(f">{name}\n{alphabet[row].tobytes().decode()}\n")


class Workload:
    """Everything an engine needs: symbols/weights (this rank's shard), DAG, priors, op lists."""

    def __init__(self, name: str, dag: GPDAG, symbols: np.ndarray, weights: np.ndarray, site_count: int,
                 seed_tree: Optional[RootedTree] = None):
        self.name = name
        self.dag = dag
        self.symbols = symbols
        self.weights = weights
        self.site_count = site_count
        self.seed_tree = seed_tree
        self.sbn_prior = dag.build_uniform_on_topological_support_prior()
        self.unconditional = dag.unconditional_node_probabilities(self.sbn_prior)
        self.inverted = dag.inverted_gpcsp_probabilities(self.sbn_prior, self.unconditional)
        self._lists: Dict[str, Tuple[np.ndarray, np.ndarray]] = {}

    def ops(self, which: str) -> Tuple[np.ndarray, np.ndarray]:
        if which not in self._lists:
            self._lists[which] = getattr(self.dag, which)().arrays()
        return self._lists[which]

    @property
    def pattern_count(self) -> int:
        return int(self.symbols.shape[1])

    def updates_per_pass(self) -> int:
        """one unit = one IncrementWithWeightedEvolvedPLV on one pattern; a full
        pass performs 2 * (E - R) of them per pattern."""
        return 2 * (self.dag.edge_count - self.dag.rootsplit_count)

    def subsample(self, pattern_count: int) -> "Workload":
        w = Workload.__new__(Workload)
        w.__dict__.update(self.__dict__)
        w.symbols = np.ascontiguousarray(self.symbols[:, :pattern_count])
        w.weights = np.ascontiguousarray(self.weights[:pattern_count])
        w.site_count = int(w.weights.sum())
        return w


def make_workload(name: str, taxon_count: int, pattern_count: int, tree_count: int, moves_per_tree: int,
                  seed: int, total_tree_length: float = 20.0, rank: int = 0, gap_rate: float = 0.01,
                  walk: bool = False) -> Workload:
    """Deterministic in (arguments, rank): every rank builds the same seed tree and DAG and
    simulates its own `pattern_count` columns (its shard of the alignment)."""
    topo_rng = np.random.default_rng(seed)
    mean_bl = total_tree_length / (2 * taxon_count - 2)
    tree = random_tree(taxon_count, topo_rng, mean_bl)
    dag = nni_neighbourhood_dag(tree, tree_count, moves_per_tree, np.random.default_rng(seed + 1), walk)
    col_rng = np.random.default_rng([seed + 2, rank])
    def draw(n):
        s = simulate_alignment(tree, n, col_rng, gap_rate)
        # multiplicities as site-pattern compression would give them: mostly 1
        w = np.where(col_rng.random(n) < 0.85, 1.0, col_rng.integers(2, 6, size=n).astype(np.float64))
        return s, w

    # `pattern_count` DISTINCT columns, weights = multiplicities (duplicates of the draw are merged)
    symbols, weights = distinct_patterns(*draw(pattern_count), draw)
    return Workload(name, dag, np.ascontiguousarray(symbols), weights, int(weights.sum()), tree)


# The synthetic configurations of BASELINE.json (configs[3] and configs[4]).
CONFIGS = {
    # 200 taxa x 100k site patterns, DAG from 1000 trees, single B200
    "synthetic-200taxa-100kpat-1000trees": dict(taxon_count=200, pattern_count=100_000, tree_count=1000,
                                                moves_per_tree=2, seed=1, walk=True),
    # 1000 taxa x 1M site patterns, DAG from 5000 trees, 125k patterns per GPU over 8 GPUs
    "synthetic-1000taxa-1Mpat-5000trees": dict(taxon_count=1000, pattern_count=125_000, tree_count=5000,
                                               moves_per_tree=1, seed=3),
    # small shapes for tests / smoke
    "synthetic-tiny": dict(taxon_count=12, pattern_count=700, tree_count=8, moves_per_tree=2, seed=5),
    "synthetic-small": dict(taxon_count=40, pattern_count=5000, tree_count=60, moves_per_tree=2, seed=7),
}


def make_named_workload(name: str, rank: int = 0, pattern_count: Optional[int] = None) -> Workload:
    cfg = dict(CONFIGS[name])
    if pattern_count is not None:
        cfg["pattern_count"] = pattern_count
    return make_workload(name, rank=rank, **cfg)
