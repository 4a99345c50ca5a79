"""Host-side PBN description: genes, predictor functions as truth tables, packing helpers.

A :class:`PBNNetwork` is what the loader produces (from ISPL text, from the JSON fixtures, or
from the ``genes=/logic_functions=`` kwargs the reference passes to ``gym.make``,
train_assa_BQN.py:121-124) and what the C-ABI consumes (``pbn_net_desc`` in include/pbn_b200.h).

State packing convention: bit ``i`` of a state is gene ``i`` in the
network's gene order (``Vars:`` order of the ISPL file); genes 0..63 live in 64-bit word 0,
genes 64..127 in word 1.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Iterable, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np

from .ispl import BoolFunction, IsplError, compile_expression, parse_ispl, render_ispl

__all__ = ["PBNNetwork", "MAX_GENES", "MAX_ARITY", "pack_states", "unpack_states", "words_for"]

MAX_GENES = 128  # two 64-bit words per state
MAX_ARITY = 16  # PBN_MAX_WIDE_ARITY: up to 6 inputs use a 64-bit truth table, 7..16 a multi-word one (scalar kernel)
NARROW_ARITY = 6  # PBN_MAX_ARITY


def words_for(n_genes: int) -> int:
    """Number of 64-bit words per packed state."""
    return 1 if n_genes <= 64 else 2


def pack_states(bits: np.ndarray) -> np.ndarray:
    """``[..., N]`` 0/1 array -> ``[..., W]`` uint64 (bit i of the state = gene i)."""
    bits = np.asarray(bits)
    n = bits.shape[-1]
    w = words_for(n)
    out = np.zeros(bits.shape[:-1] + (w,), dtype=np.uint64)
    for i in range(n):
        out[..., i >> 6] |= (bits[..., i].astype(np.uint64) & np.uint64(1)) << np.uint64(i & 63)
    return out


def unpack_states(words: np.ndarray, n_genes: int) -> np.ndarray:
    """``[..., W]`` uint64 -> ``[..., N]`` uint8."""
    words = np.asarray(words, dtype=np.uint64)
    out = np.zeros(words.shape[:-1] + (n_genes,), dtype=np.uint8)
    for i in range(n_genes):
        out[..., i] = ((words[..., i >> 6] >> np.uint64(i & 63)) & np.uint64(1)).astype(np.uint8)
    return out


@dataclass
class PBNNetwork:
    """Genes + per-gene predictor functions (truth tables) + selection probabilities."""

    genes: List[str]
    functions: List[List[BoolFunction]]
    probabilities: List[List[float]] = field(default_factory=list)
    name: str = ""

    def __post_init__(self):
        n = len(self.genes)
        if not 1 <= n <= MAX_GENES:
            raise IsplError("network has %d genes; supported range is 1..%d" % (n, MAX_GENES))
        if len(self.functions) != n:
            raise IsplError("need one function list per gene")
        if not self.probabilities:
            self.probabilities = [[1.0 / len(fs)] * len(fs) if fs else [] for fs in self.functions]
        for i, (fs, ps) in enumerate(zip(self.functions, self.probabilities)):
            if not fs:
                raise IsplError("gene %r has no predictor function" % self.genes[i])
            if len(fs) != len(ps):
                raise IsplError("gene %r: %d functions but %d probabilities" % (self.genes[i], len(fs), len(ps)))
            tot = float(sum(ps))
            if tot <= 0 or any(p < 0 for p in ps):
                raise IsplError("gene %r: bad selection probabilities %r" % (self.genes[i], ps))
            self.probabilities[i] = [float(p) / tot for p in ps]
            for f in fs:
                if f.arity > MAX_ARITY:
                    raise IsplError("gene %r: predictor arity %d exceeds %d" % (self.genes[i], f.arity, MAX_ARITY))

    # ---------------------------------------------------------------- constructors
    @classmethod
    def from_expressions(cls, genes: Sequence[str], exprs: Sequence[Sequence[Union[str, Tuple[str, float]]]],
                         name: str = "") -> "PBNNetwork":
        """``exprs[i]`` = list of expression strings or ``(expr, prob)`` tuples for gene ``i``."""
        genes = [str(g) for g in genes]
        index = {g: i for i, g in enumerate(genes)}
        if len(index) != len(genes):
            raise IsplError("duplicate gene names")
        funcs: List[List[BoolFunction]] = []
        probs: List[List[float]] = []
        for row in exprs:
            fs, ps = [], []
            for item in row:
                if isinstance(item, (tuple, list)):
                    e, p = item[0], float(item[1])
                else:
                    e, p = item, None
                fs.append(compile_expression(str(e), index))
                ps.append(p)
            if any(p is None for p in ps):
                ps = [1.0 / len(fs)] * len(fs)
            funcs.append(fs)
            probs.append(ps)
        return cls(genes, funcs, probs, name)

    @classmethod
    def from_ispl(cls, text: str, name: str = "") -> "PBNNetwork":
        genes, funcs = parse_ispl(text)
        return cls.from_expressions(genes, [funcs[g] for g in genes], name)

    @classmethod
    def from_ispl_file(cls, path: Union[str, Path]) -> "PBNNetwork":
        p = Path(path)
        return cls.from_ispl(p.read_text(), name=p.stem)

    @classmethod
    def from_json(cls, path: Union[str, Path]) -> "PBNNetwork":
        """Load a ``{"genes": [...], "functions": [[expr,...],...]}`` file (tests/golden/pbn*.json)."""
        d = json.loads(Path(path).read_text())
        return cls.from_expressions(d["genes"], d["functions"], d.get("name", Path(path).stem))

    @classmethod
    def from_logic_functions(cls, genes: Sequence[str],
                             logic_functions: Union[Sequence, Mapping], name: str = "") -> "PBNNetwork":
        """The ``genes=/logic_functions=`` kwargs of ``gym.make("gym-PBN/PBNEnv", ...)``
        (train_assa_BQN.py:121-124 passes a list; train_assa_matlab_BQN.py:171 a dict keyed by
        gene index)."""
        if isinstance(logic_functions, Mapping):
            keys = list(logic_functions.keys())
            if all(k in logic_functions for k in range(len(genes))):
                rows = [logic_functions[i] for i in range(len(genes))]
            else:
                rows = [logic_functions[g] for g in genes]
            del keys
        else:
            rows = list(logic_functions)
        return cls.from_expressions(genes, rows, name)

    # ---------------------------------------------------------------- properties
    @property
    def n_genes(self) -> int:
        return len(self.genes)

    @property
    def n_words(self) -> int:
        return words_for(self.n_genes)

    @property
    def n_functions(self) -> int:
        return sum(len(fs) for fs in self.functions)

    @property
    def max_arity(self) -> int:
        return max(f.arity for fs in self.functions for f in fs)

    @property
    def is_uniform(self) -> bool:
        """True when every gene selects uniformly among its predictors."""
        return all(max(ps) - min(ps) < 1e-12 for ps in self.probabilities)

    def state_mask(self) -> Tuple[int, ...]:
        n = self.n_genes
        if n <= 64:
            return ((1 << n) - 1,)
        return ((1 << 64) - 1, (1 << (n - 64)) - 1)

    # ---------------------------------------------------------------- host evaluation (tiny; not the product path)
    def next_state_int(self, state: int, sel: Sequence[int]) -> int:
        """Deterministic synchronous update of one state given per-gene function choices.
        Host helper for loaders/attractor search on small networks; the kernels do the real work."""
        out = 0
        for i, fs in enumerate(self.functions):
            f = fs[sel[i]]
            a = 0
            for j, g in enumerate(f.inputs):
                a |= ((state >> g) & 1) << j
            out |= ((f.lut >> a) & 1) << i
        return out

    # ---------------------------------------------------------------- device descriptor arrays
    def descriptor_arrays(self) -> Dict[str, np.ndarray]:
        """Flat arrays for ``pbn_net_desc`` (include/pbn_b200.h)."""
        n = self.n_genes
        offs = np.zeros(n + 1, dtype=np.int32)
        arity, inputs, luts, cum = [], [], [], []
        wide_inputs, wide_offs, wide_words = [], [0], []
        for i, (fs, ps) in enumerate(zip(self.functions, self.probabilities)):
            offs[i + 1] = offs[i] + len(fs)
            acc = 0.0
            for k, (f, p) in enumerate(zip(fs, ps)):
                arity.append(f.arity)
                if f.arity > NARROW_ARITY:      # wide predictor: multi-word truth table, func_lut = its index
                    inputs.append([0] * 8)
                    luts.append(len(wide_inputs))
                    wide_inputs.append(list(f.inputs) + [0] * (16 - f.arity))
                    nw = 1 << (f.arity - 6)
                    wide_words += [(f.lut >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(nw)]
                    wide_offs.append(wide_offs[-1] + nw)
                else:
                    inputs.append(list(f.inputs) + [0] * (8 - f.arity))
                    luts.append(f.lut)
                acc += p
                # cumulative threshold in 2^-32 units; the last one is pinned to 2^32-1 (inclusive top)
                thr = 0xFFFFFFFF if k == len(fs) - 1 else min(int(round(acc * 4294967296.0)), 0xFFFFFFFF)
                cum.append(thr)
        return {
            "func_offset": offs,
            "func_arity": np.asarray(arity, dtype=np.uint8),
            "func_inputs": np.asarray(inputs, dtype=np.uint8).reshape(-1, 8),
            "func_lut": np.asarray(luts, dtype=np.uint64),
            "func_cum": np.asarray(cum, dtype=np.uint32),
            "wide_inputs": np.asarray(wide_inputs, dtype=np.uint8).reshape(-1, 16),
            "wide_lut_offset": np.asarray(wide_offs, dtype=np.int32),
            "wide_lut": np.asarray(wide_words, dtype=np.uint64),
        }

    # ---------------------------------------------------------------- writers
    def to_ispl(self, template: Optional[str] = None) -> str:
        """Render back to ISPL through the reference template layout (model_template.jj2).
        Gene names of the form ``x<ID>`` are written as ``<ID>`` keys, as the template prefixes ``x``."""
        log_funcs: Dict[str, List[str]] = {}
        for g, fs in zip(self.genes, self.functions):
            key = g[1:] if g.startswith("x") else g
            log_funcs[key] = [f.expr for f in fs]
        return render_ispl(log_funcs, template)

    # ---------------------------------------------------------------- graph introspection (reference a-9)
    def adjacency(self) -> List[List[int]]:
        """``adj[i]`` = sorted gene indices feeding any predictor of gene ``i``
        (what ``env.graph.get_adj_list()`` serves to the GNN agents, gbdq_model/__init__.py:259-277)."""
        return [sorted({g for f in fs for g in f.inputs}) for fs in self.functions]
