"""Other on-disk network formats the reference's scripts feed to the env.

* ASSA-PBN "matlab" text format -- truth tables + predictor sets + selection probabilities +
  perturbation rate.  Reference parser: train_assa_matlab_BQN.py:72-160 (it turns the truth tables
  into sum-of-products expression strings with sympy and hands ``(expr, prob)`` lists to
  ``gym.make("gym-PBN/PBNEnv", genes=, logic_functions=)``); here the truth tables go straight into
  the device descriptors.
* ``.bnet`` (boolnet/biodivine): ``target, factors`` lines with ``& | !`` -- models/bb33/bb33.bnet is the
  same network as models/bb33/bb33.ispl.
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Tuple, Union

from .ispl import BoolFunction, IsplError, compile_expression
from .network import PBNNetwork

__all__ = ["parse_assa_matlab", "network_from_assa_matlab", "parse_bnet", "network_from_bnet"]


def parse_assa_matlab(text: str, index_base: Union[int, str] = 0):
    """Returns ``(genes, functions, probabilities, perturbation_rate)``.

    Layout (train_assa_matlab_BQN.py:72-137): two header lines; number of genes; functions per gene;
    number of predictors of every function (flat); one truth-table line per function (2^k values,
    column j = assignment ``itertools.product([0,1], repeat=k)[j]``, i.e. the FIRST predictor is the
    most significant bit of j, :109-113); one predictor-set line per function; one probability line
    per gene; the perturbation rate; one trailing line.  Genes are named ``x0 .. x{n-1}`` and a
    predictor token ``t`` names gene ``x{t}`` (:126,160) -- ``index_base=1`` (or ``"auto"``) shifts
    1-based files as written by ASSA-PBN's own matlab tooling."""
    lines = [ln for ln in text.splitlines()]
    it = iter(lines)

    def nxt():
        try:
            return next(it)
        except StopIteration:
            raise IsplError("ASSA file ends early") from None

    nxt(), nxt()
    n = int(nxt().split()[0])
    n_funcs = [int(t) for t in nxt().split()]
    if len(n_funcs) != n:
        raise IsplError("expected %d function counts, got %d" % (n, len(n_funcs)))
    n_pred = [int(t) for t in nxt().split()]
    if len(n_pred) != sum(n_funcs):
        raise IsplError("expected %d predictor counts, got %d" % (sum(n_funcs), len(n_pred)))
    tables: List[List[int]] = []
    for k in n_pred:
        vals = [float(t) for t in nxt().split()]
        if len(vals) != 1 << k:
            raise IsplError("truth table with %d entries for %d predictors" % (len(vals), k))
        tables.append([1 if v else 0 for v in vals])
    preds: List[List[int]] = []
    for k in n_pred:
        toks = [int(t) for t in nxt().split()]
        if len(toks) != k:
            raise IsplError("predictor set with %d entries, expected %d" % (len(toks), k))
        preds.append(toks)
    probs = [[float(t) for t in nxt().split()] for _ in range(n)]
    rate = float(nxt().split()[0])
    flat = [t for row in preds for t in row]
    if index_base == "auto":
        index_base = 1 if flat and min(flat) >= 1 and max(flat) == n else 0
    genes = ["x%d" % i for i in range(n)]
    functions: List[List[BoolFunction]] = []
    f = 0
    for g in range(n):
        row = []
        if len(probs[g]) != n_funcs[g]:
            raise IsplError("gene %d: %d functions but %d probabilities" % (g, n_funcs[g], len(probs[g])))
        for _ in range(n_funcs[g]):
            k = n_pred[f]
            ins = [t - int(index_base) for t in preds[f]]
            if any(not 0 <= t < n for t in ins):
                raise IsplError("function %d: predictor index outside 0..%d" % (f, n - 1))
            lut = 0
            for a in range(1 << k):      # bit i of a = value of predictor i; the file's column index is its bit reversal
                j = 0
                for i in range(k):
                    j |= ((a >> i) & 1) << (k - 1 - i)
                lut |= tables[f][j] << a
            row.append(_reduced(ins, lut, "assa_tt_%d" % f))
            f += 1
        functions.append(row)
    return genes, functions, probs, rate


def _reduced(inputs: List[int], lut: int, label: str) -> BoolFunction:
    """Sort inputs by gene index, merge duplicates, drop inessential ones (as compile_expression does)."""
    k = len(inputs)
    uniq = sorted(set(inputs))
    m = len(uniq)
    table = 0
    for a in range(1 << m):
        full = 0
        for i, g in enumerate(inputs):
            full |= ((a >> uniq.index(g)) & 1) << i
        table |= ((lut >> full) & 1) << a
    essential = []
    for j in range(m):
        if any(((table >> a) & 1) != ((table >> (a ^ (1 << j))) & 1) for a in range(1 << m)):
            essential.append(j)
    if len(essential) != m:
        t2 = 0
        for a in range(1 << len(essential)):
            full = 0
            for nj, j in enumerate(essential):
                full |= ((a >> nj) & 1) << j
            t2 |= ((table >> full) & 1) << a
        table, uniq = t2, [uniq[j] for j in essential]
    del k
    return BoolFunction(tuple(uniq), table, label)


def network_from_assa_matlab(path_or_text: Union[str, Path], index_base: Union[int, str] = 0) -> Tuple[PBNNetwork, float]:
    """``(network, perturbation_rate)`` from an ASSA-PBN matlab-format file (or its text)."""
    text, name = _read(path_or_text)
    genes, functions, probs, rate = parse_assa_matlab(text, index_base)
    return PBNNetwork(genes, functions, probs, name), rate


def parse_bnet(text: str) -> Tuple[List[str], List[str]]:
    """``(genes, expressions)`` in file order; the ``targets,factors`` header and ``#`` comments are skipped."""
    genes, exprs = [], []
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if "," not in line:
            raise IsplError("bnet line without ',': %r" % raw)
        name, expr = line.split(",", 1)
        name, expr = name.strip(), expr.strip()
        if name.lower() == "targets" and expr.lower().replace(" ", "") == "factors":
            continue
        genes.append(name)
        exprs.append(expr)
    if len(set(genes)) != len(genes):
        raise IsplError("duplicate target in bnet file")
    return genes, exprs


def network_from_bnet(path_or_text: Union[str, Path]) -> PBNNetwork:
    """A Boolean network (one predictor per gene, probability 1) from ``.bnet`` text or file."""
    text, name = _read(path_or_text)
    genes, exprs = parse_bnet(text)
    index = {g: i for i, g in enumerate(genes)}
    return PBNNetwork(genes, [[compile_expression(e, index)] for e in exprs], [[1.0]] * len(genes), name)


def _read(path_or_text: Union[str, Path]) -> Tuple[str, str]:
    if isinstance(path_or_text, Path) or ("\n" not in str(path_or_text) and Path(str(path_or_text)).exists()):
        p = Path(path_or_text)
        return p.read_text(), p.stem
    return str(path_or_text), ""
