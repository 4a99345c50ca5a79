"""Attractor sets: the reference's pickle format, device tables, and a host-side finder.

Reference format (data/attractors_Bittner-7.pkl, bns_attractors/*.pkl):
``list`` of attractors, each a ``list`` of states, each state a ``tuple`` of length N whose
elements are ``0/1`` (python ``int`` or ``numpy.int64``) or the wildcard string ``'*'``.
``env.all_attractors`` serves exactly this structure to the agents
(model_tester.py:564,599-604; bdq_model/__init__.py:182).

On the device an attractor set is a CSR table of ``(care, value)`` word pairs: state ``s``
matches entry ``e`` iff ``(s & care[e]) == value[e]`` -- wildcards clear bits in ``care``.
"""
from __future__ import annotations

import pickle
import warnings
from dataclasses import dataclass
from pathlib import Path
from typing import Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from .network import PBNNetwork, words_for

__all__ = [
    "AttractorSet",
    "load_attractor_pickle",
    "save_attractor_pickle",
    "find_attractors_stg",
    "sorted_id_permutation",
]

State = Tuple[Union[int, str], ...]


def _norm_state(state: Sequence) -> State:
    out = []
    for v in state:
        if isinstance(v, str):
            if v != "*":
                raise ValueError("bad attractor element %r" % (v,))
            out.append("*")
        else:
            iv = int(v)
            if iv not in (0, 1):
                raise ValueError("bad attractor element %r" % (v,))
            out.append(iv)
    return tuple(out)


@dataclass
class AttractorSet:
    """``attractors[a]`` = list of state tuples (0/1/'*') in the env's gene order."""

    attractors: List[List[State]]
    n_genes: int

    def __post_init__(self):
        self.attractors = [[_norm_state(s) for s in attr] for attr in self.attractors]
        for attr in self.attractors:
            if not attr:
                raise ValueError("empty attractor")
            for s in attr:
                if len(s) != self.n_genes:
                    raise ValueError("attractor state of length %d for a %d-gene network" % (len(s), self.n_genes))

    def __len__(self) -> int:
        return len(self.attractors)

    @property
    def n_words(self) -> int:
        return words_for(self.n_genes)

    def permuted(self, perm: Sequence[int]) -> "AttractorSet":
        """Re-order gene positions: element ``k`` of every stored tuple moves to position ``perm[k]``.
        Needed for data/attractors_Bittner-28.pkl, which is in ascending gene-ID order while
        kaban/pbn28.ispl is in dataset order."""
        out = []
        for attr in self.attractors:
            row = []
            for s in attr:
                t: List[Union[int, str]] = [0] * self.n_genes
                for k, v in enumerate(s):
                    t[perm[k]] = v
                row.append(tuple(t))
            out.append(row)
        return AttractorSet(out, self.n_genes)

    def tables(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """``(offset[A+1] int32, care[S,W] uint64, value[S,W] uint64)``."""
        w = self.n_words
        offs = [0]
        care, val = [], []
        for attr in self.attractors:
            for s in attr:
                c = [0] * w
                v = [0] * w
                for i, b in enumerate(s):
                    if b == "*":
                        continue
                    c[i >> 6] |= 1 << (i & 63)
                    if b:
                        v[i >> 6] |= 1 << (i & 63)
                care.append(c)
                val.append(v)
            offs.append(len(care))
        return (np.asarray(offs, dtype=np.int32),
                np.asarray(care, dtype=np.uint64).reshape(-1, w),
                np.asarray(val, dtype=np.uint64).reshape(-1, w))

    def representative_words(self) -> np.ndarray:
        """``[A, W]`` uint64: first state of each attractor with ``'*' -> 0`` -- the start/target
        state convention of the evaluator (model_tester.py:604-609)."""
        _, _, val = self.tables()
        offs, _, _ = self.tables()
        return val[offs[:-1]]

    def contains(self, attractor_id: int, state_bits: Sequence[int]) -> bool:
        for s in self.attractors[attractor_id]:
            if all(b == "*" or int(b) == int(x) for b, x in zip(s, state_bits)):
                return True
        return False

    def attractor_of(self, state_bits: Sequence[int]) -> int:
        """Index of the first attractor containing the state, or -1."""
        for a in range(len(self.attractors)):
            if self.contains(a, state_bits):
                return a
        return -1


def load_attractor_pickle(path: Union[str, Path], n_genes: Optional[int] = None) -> AttractorSet:
    """Load an attractor pickle written by the reference's env (mixed ``int``/``np.int64``/``'*'``)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # numpy.core.multiarray.scalar deprecation under numpy>=2
        with open(path, "rb") as f:
            raw = pickle.load(f)
    n = n_genes if n_genes is not None else len(raw[0][0])
    return AttractorSet([list(attr) for attr in raw], n)


def save_attractor_pickle(path: Union[str, Path], attractors: AttractorSet) -> None:
    with open(path, "wb") as f:
        pickle.dump([list(attr) for attr in attractors.attractors], f)


def sorted_id_permutation(genes: Sequence[str]) -> List[int]:
    """``perm[k]`` = position in ``genes`` of the gene with the k-th smallest numeric ID
    (gene names ``x<ID>``).  Maps the sorted-ID layout of data/attractors_Bittner-28.pkl onto a
    network's own gene order."""
    ids = [int(g[1:]) if g[:1] == "x" and g[1:].isdigit() else None for g in genes]
    if any(i is None for i in ids):
        raise ValueError("gene names are not of the form x<ID>")
    return sorted(range(len(genes)), key=lambda i: ids[i])


# --------------------------------------------------------------------------------------
# Host-side attractor finder for small networks (N <= ~20): sink SCCs of the
# perturbation-free state-transition graph (what print_graph.py:15-34 does with networkx).
# --------------------------------------------------------------------------------------

def _successor_sets(net: PBNNetwork) -> Tuple[np.ndarray, np.ndarray]:
    """Per state: bit-mask of genes that *can* become 1 and that *can* become 0."""
    n = net.n_genes
    states = np.arange(1 << n, dtype=np.int64)
    can1 = np.zeros(1 << n, dtype=np.int64)
    can0 = np.zeros(1 << n, dtype=np.int64)
    for i, fs in enumerate(net.functions):
        for f in fs:
            a = np.zeros(1 << n, dtype=np.int64)
            for j, g in enumerate(f.inputs):
                a |= ((states >> g) & 1) << j
            lut = np.array([(f.lut >> k) & 1 for k in range(1 << f.arity)], dtype=np.int64)
            v = lut[a]
            can1 |= v << i
            can0 |= (1 - v) << i
    return can1, can0


def _successors(s: int, can1: int, can0: int, n: int) -> List[int]:
    fixed = 0
    free = []
    for i in range(n):
        one, zero = (can1 >> i) & 1, (can0 >> i) & 1
        if one and zero:
            free.append(i)
        elif one:
            fixed |= 1 << i
    out = [fixed]
    for i in free:
        out += [t | (1 << i) for t in out]
    return out


def find_attractors_stg(net: PBNNetwork, max_genes: int = 20) -> Tuple[AttractorSet, dict]:
    """Exhaustive attractor search: build the STG (edge s->t iff some per-gene predictor choice
    yields t), return its sink strongly-connected components as an :class:`AttractorSet`
    (states sorted ascending inside each attractor, attractors sorted by smallest state) and
    a dict with ``n_edges`` / ``n_sccs`` (fixture K5).  Iterative Tarjan; host only."""
    n = net.n_genes
    if n > max_genes:
        raise ValueError("exhaustive STG search is limited to %d genes (got %d)" % (max_genes, n))
    can1, can0 = _successor_sets(net)
    n_states = 1 << n
    succ = [_successors(s, int(can1[s]), int(can0[s]), n) for s in range(n_states)]
    n_edges = sum(len(x) for x in succ)

    index = [-1] * n_states
    low = [0] * n_states
    on_stack = [False] * n_states
    comp = [-1] * n_states
    stack: List[int] = []
    counter = 0
    n_comp = 0
    for root in range(n_states):
        if index[root] != -1:
            continue
        work = [(root, 0)]
        while work:
            v, pi = work.pop()
            if pi == 0:
                index[v] = low[v] = counter
                counter += 1
                stack.append(v)
                on_stack[v] = True
            recurse = False
            sv = succ[v]
            for k in range(pi, len(sv)):
                w = sv[k]
                if index[w] == -1:
                    work.append((v, k + 1))
                    work.append((w, 0))
                    recurse = True
                    break
                if on_stack[w]:
                    low[v] = min(low[v], index[w])
            if recurse:
                continue
            if low[v] == index[v]:
                while True:
                    w = stack.pop()
                    on_stack[w] = False
                    comp[w] = n_comp
                    if w == v:
                        break
                n_comp += 1
            if work:
                parent = work[-1][0]
                low[parent] = min(low[parent], low[v])
    is_sink = [True] * n_comp
    for s in range(n_states):
        for t in succ[s]:
            if comp[t] != comp[s]:
                is_sink[comp[s]] = False
    members: dict = {}
    for s in range(n_states):
        if is_sink[comp[s]]:
            members.setdefault(comp[s], []).append(s)
    sinks = sorted(sorted(m) for m in members.values())
    attrs = [[tuple((s >> i) & 1 for i in range(n)) for s in m] for m in sinks]
    return AttractorSet(attrs, n), {"n_states": n_states, "n_edges": n_edges, "n_sccs": n_comp,
                                   "sink_sccs": sinks}
