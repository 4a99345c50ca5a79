"""ISPL network files: reader, writer and Boolean-expression compiler.

The reference keeps its benchmark networks as ISPL text (``kaban/pbn{7,10,28,70}.ispl``),
rendered from ``model_template.jj2`` (reference: model_template.jj2:1-23), and its scripts
re-parse ISPL into ``genes`` + ``logic_functions`` for ``gym.make("gym-PBN/PBNEnv", ...)``
(reference: train_assa_BQN.py:51-124, model_tester.py:344-413).  This module accepts both
dialects found in the reference tree:

* the *kaban* dialect: tab-indented, blank lines between ``Vars:`` entries, several
  ``x<ID>=true if (<expr>)=true;`` / ``x<ID>=false if (<expr>)=false;`` pairs per gene;
* the *compact* dialect (models/bb33/bb33.ispl:41-49): ``v_AP=true  if ((a&b)|~c)=true;``,
  one function per gene, bare-identifier right-hand sides (``v_DCII=true  if v_DCI=true;``).

Expressions are compiled by a small recursive-descent parser (no ``eval``) into truth
tables over their *essential* inputs; the truth tables are what the CUDA kernels consume.
Operators accepted: ``~ ! not``, ``& && and``, ``| || or``, parentheses, the constants
``true/false/True/False/1/0`` and ISPL atoms ``name=true`` / ``name=false``.
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

__all__ = [
    "IsplError",
    "BoolFunction",
    "compile_expression",
    "parse_ispl",
    "render_ispl",
    "logic_functions_from_ispl",
]


class IsplError(ValueError):
    """Raised for malformed ISPL text or Boolean expressions."""


# --------------------------------------------------------------------------------------
# Boolean expressions -> truth tables
# --------------------------------------------------------------------------------------

_TOKEN = re.compile(r"\s*(?:(\|\||&&|[()~!&|])|([A-Za-z_][A-Za-z_0-9.]*|[01])(?:\s*=\s*(true|false))?)")

_TRUE = {"true", "True", "TRUE", "1"}
_FALSE = {"false", "False", "FALSE", "0"}
_OR = {"|", "||", "or"}
_AND = {"&", "&&", "and"}
_NOT = {"~", "!", "not"}


def _tokenize(text: str) -> List[Tuple[str, str]]:
    """Split into (kind, value) tokens; kind in {'op', 'id', 'const', 'idneg'}."""
    out: List[Tuple[str, str]] = []
    pos = 0
    n = len(text)
    while pos < n:
        if text[pos].isspace():
            pos += 1
            continue
        m = _TOKEN.match(text, pos)
        if not m:
            raise IsplError("cannot tokenize %r at offset %d" % (text, pos))
        pos = m.end()
        op, ident, eq = m.groups()
        if op:
            out.append(("op", op))
        elif ident in _OR | _AND | _NOT and eq is None:
            out.append(("op", ident))
        elif ident in _TRUE | _FALSE and eq is None:
            out.append(("const", "1" if ident in _TRUE else "0"))
        elif eq == "false":
            out.append(("idneg", ident))
        else:
            out.append(("id", ident))
    return out


class _Parser:
    """expr := and ('|' and)* ; and := not ('&' not)* ; not := '~' not | atom."""

    def __init__(self, tokens, columns: Mapping[str, int], ones: int):
        self.toks = tokens
        self.i = 0
        self.columns = columns
        self.ones = ones

    def _peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else (None, None)

    def parse(self) -> int:
        v = self._or()
        if self.i != len(self.toks):
            raise IsplError("unexpected token %r" % (self.toks[self.i][1],))
        return v

    def _or(self) -> int:
        v = self._and()
        while self._peek()[0] == "op" and self._peek()[1] in _OR:
            self.i += 1
            v |= self._and()
        return v

    def _and(self) -> int:
        v = self._not()
        while self._peek()[0] == "op" and self._peek()[1] in _AND:
            self.i += 1
            v &= self._not()
        return v

    def _not(self) -> int:
        kind, val = self._peek()
        if kind == "op" and val in _NOT:
            self.i += 1
            return self.ones ^ self._not()
        return self._atom()

    def _atom(self) -> int:
        kind, val = self._peek()
        if kind is None:
            raise IsplError("unexpected end of expression")
        self.i += 1
        if kind == "op":
            if val != "(":
                raise IsplError("unexpected operator %r" % val)
            v = self._or()
            if self._peek() != ("op", ")"):
                raise IsplError("missing ')'")
            self.i += 1
            return v
        if kind == "const":
            return self.ones if val == "1" else 0
        col = self.columns[val]
        return (self.ones ^ col) if kind == "idneg" else col


@dataclass(frozen=True)
class BoolFunction:
    """A predictor function reduced to its essential inputs.

    ``inputs[j]`` is the gene index feeding bit ``j`` of the truth-table index (LSB first);
    bit ``a`` of ``lut`` is the function value when the inputs spell the integer ``a``.
    """

    inputs: Tuple[int, ...]
    lut: int
    expr: str = ""

    @property
    def arity(self) -> int:
        return len(self.inputs)

    def __call__(self, state_bits: Sequence[int]) -> int:
        a = 0
        for j, g in enumerate(self.inputs):
            a |= (int(state_bits[g]) & 1) << j
        return (self.lut >> a) & 1


def _var_column(j: int, k: int) -> int:
    """Truth-table column of variable j among k variables, as a 2^k-bit integer."""
    period = 1 << (j + 1)
    half = 1 << j
    block = ((1 << half) - 1) << half  # 'half' zeros then 'half' ones
    col = 0
    for start in range(0, 1 << k, period):
        col |= block << start
    return col


def compile_expression(expr: str, gene_index: Mapping[str, int], max_syntactic: int = 20) -> BoolFunction:
    """Compile one Boolean expression into a :class:`BoolFunction` over gene indices.

    Variables the function does not actually depend on (e.g. ``x & ~x`` minterms,
    kaban/pbn7.ispl:48) are dropped from the support, so the arity is the essential arity.
    """
    tokens = _tokenize(expr)
    names: List[str] = []
    for kind, val in tokens:
        if kind in ("id", "idneg") and val not in names:
            if val not in gene_index:
                raise IsplError("unknown gene %r in expression %r" % (val, expr))
            names.append(val)
    names.sort(key=lambda nm: gene_index[nm])
    k = len(names)
    if k > max_syntactic:
        raise IsplError("expression mentions %d genes (limit %d): %r" % (k, max_syntactic, expr))
    ones = (1 << (1 << k)) - 1
    columns = {nm: _var_column(j, k) for j, nm in enumerate(names)}
    table = _Parser(tokens, columns, ones).parse()

    # essential support: variable j matters iff cofactors differ somewhere
    essential = []
    for j in range(k):
        col = columns[names[j]]
        hi = (table & col) >> (1 << j)
        lo = table & (ones ^ col)
        if hi != lo:
            essential.append(j)
    if len(essential) != k:
        new_k = len(essential)
        new_table = 0
        for a in range(1 << new_k):
            full = 0
            for nj, j in enumerate(essential):
                if (a >> nj) & 1:
                    full |= 1 << j
            if (table >> full) & 1:
                new_table |= 1 << a
        table = new_table
        names = [names[j] for j in essential]
    return BoolFunction(tuple(gene_index[nm] for nm in names), table, expr)


# --------------------------------------------------------------------------------------
# ISPL text
# --------------------------------------------------------------------------------------

_VAR_LINE = re.compile(r"^([A-Za-z_][A-Za-z_0-9.]*)\s*:\s*boolean\s*;?$")
_EVO_LINE = re.compile(r"^([A-Za-z_][A-Za-z_0-9.]*)\s*=\s*(true|false)\s+if\s+(.*?)\s*=\s*(true|false)\s*;?$")


def parse_ispl(text: str) -> Tuple[List[str], Dict[str, List[str]]]:
    """Parse ISPL text into ``(genes, {gene: [expr, ...]})``.

    ``genes`` keeps the ``Vars:`` order (this *is* the env's gene order: bit ``i`` of a packed
    state is the ``i``-th ``Vars:`` entry).  Only the ``=true if ...=true`` line of each pair is
    used, as in the reference loader (train_assa_BQN.py:87-89); expressions are returned with
    the outer parentheses of ``(<expr>)=true`` removed when they wrap the whole expression.
    Blank lines inside ``Vars:`` (the kaban dialect) are skipped; the reference loader crashes
    on them (train_assa_BQN.py:68, IndexError) -- accepting them is a deliberate extension.
    """
    genes: List[str] = []
    funcs: Dict[str, List[str]] = {}
    section: Optional[str] = None
    for raw in text.splitlines():
        line = raw.strip()
        if not line:
            continue
        head = line.split()[0]
        if head == "Vars:":
            section = "vars"
            continue
        if head == "Evolution:":
            section = "evo"
            continue
        if head == "end":
            section = None
            continue
        if section == "vars":
            m = _VAR_LINE.match(line)
            if not m:
                raise IsplError("bad Vars line: %r" % raw)
            if m.group(1) in funcs:
                raise IsplError("duplicate variable %r" % m.group(1))
            genes.append(m.group(1))
            funcs[m.group(1)] = []
        elif section == "evo":
            m = _EVO_LINE.match(line)
            if not m:
                raise IsplError("bad Evolution line: %r" % raw)
            tgt, val, expr, cond = m.groups()
            if tgt not in funcs:
                raise IsplError("Evolution line for undeclared variable %r" % tgt)
            if val == "false":
                continue  # the '=false if (...)=false' twin carries no extra information
            if cond != "true":
                expr = "~(%s)" % expr
            funcs[tgt].append(_strip_outer_parens(expr))
    if not genes:
        raise IsplError("no Vars: section found")
    return genes, funcs


def _strip_outer_parens(expr: str) -> str:
    expr = expr.strip()
    if not (expr.startswith("(") and expr.endswith(")")):
        return expr
    depth = 0
    for i, ch in enumerate(expr):
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0 and i != len(expr) - 1:
                return expr  # the first '(' closes early: not a wrapping pair
    return expr[1:-1]


def render_ispl(log_funcs: Mapping[str, Sequence[str]], template: Optional[str] = None) -> str:
    """Render ``{gene_id_without_x: [expr, ...]}`` as ISPL text.

    With ``template=None`` the built-in writer reproduces the output of the reference's
    ``model_template.jj2`` (reference: model_template.jj2:1-23) byte for byte -- fixture K1
    checks the sha256 of all four kaban files.  ``template`` may be the *text* of a jinja2
    template (e.g. the user's own copy of model_template.jj2); it is rendered with the same
    single variable ``log_funcs``.
    """
    if template is not None:
        import jinja2  # optional dependency, only for user-supplied templates

        return jinja2.Template(template).render(log_funcs=log_funcs)
    out = ["Agent M\n\tVars:\n\t\t"]
    for gene in log_funcs:
        out.append("\n\t\tx%s: boolean;\n\t\t" % gene)
    out.append("\n\tend Vars\n\tActions = {none};\n\tProtocol:\n\t\tOther: {none};\n\tend Protocol\n\tEvolution:\n\t\t")
    for key in log_funcs:
        out.append("\n\t\t")
        for fun in log_funcs[key]:
            out.append("\n\t\tx%s=true if (%s)=true;\n\t\tx%s=false if (%s)=false;\n\t\t" % (key, fun, key, fun))
        out.append("\n\t\t")
    out.append("\n\tend Evolution\nend Agent\n\nInitStates\n\t\tM.x234237=true or M.x234237=false;\nend InitStates\n")
    return "".join(out)


def logic_functions_from_ispl(text: str, prob: Optional[float] = None):
    """ISPL text -> ``(genes, logic_functions)`` in the shape the reference passes to
    ``gym.make("gym-PBN/PBNEnv", genes=..., logic_functions=...)`` (train_assa_BQN.py:109,121-124):
    ``logic_functions[i]`` is a list of ``(python_bool_expr, probability)``.

    The reference hard-codes probability 1.0 per (single) function; with several functions per
    gene we default to the uniform ``1/k`` (the only distribution evidenced in the reference,
    train_pbn_28.py:139-151), unless ``prob`` is given.
    """
    genes, funcs = parse_ispl(text)
    out = []
    for g in genes:
        k = max(len(funcs[g]), 1)
        p = prob if prob is not None else 1.0 / k
        row = []
        for e in funcs[g]:
            py = e.replace("(", " ( ").replace(")", " ) ").replace("|", " or ").replace("&", " and ").replace("~", " not ")
            row.append((py, p))
        out.append(row)
    return genes, out
