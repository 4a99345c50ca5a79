#!/usr/bin/env python
"""Generates the golden fixtures for the other network formats by RUNNING the
reference's own parsers in this container (they are plain Python; /root/reference is read, never
copied):

  assa_example.txt          a small synthetic network in the ASSA-PBN "matlab" text format (written here)
  assa_example_expected.json  what train_assa_matlab_BQN.py:50-160 makes of it: per function the python
                            expression string + probability it would pass to gym.make, the perturbation
                            rate, and the full truth table of every expression over all 2^n states
                            (evaluated with python's eval: an evaluator independent of the product)
  bb33.bnet                 models/bb33/bb33.bnet (a reference DATA file)
  bb33_expected.json        genes + python expressions the reference's ISPL parser
                            (train_assa_BQN.py:51-109) extracts from models/bb33/bb33.ispl, and the next
                            states they give for 4096 fixed inputs (sha256 + first rows)
  control14.json            the 14-gene network + control_nodes of train_control_gbdq.py:45-72 (the
                            literals are read out of the script with ast), next states for 4096 inputs

    python tests/golden/make_golden_formats.py
"""
import ast
import hashlib
import itertools
import json
import random
import types
from collections import defaultdict
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def slice_source(path, start_marker, end_marker):
    src = path.read_text().splitlines()
    a = next(i for i, ln in enumerate(src) if ln.startswith(start_marker))
    b = next(i for i, ln in enumerate(src) if i > a and ln.startswith(end_marker))
    return "\n".join(src[a:b + 1])


def write_assa_example(path):
    rng = random.Random(20241018)
    n = 5
    n_funcs = [2, 1, 3, 1, 2]
    arities = [3, 2, 1, 4, 2, 3, 2, 1, 4]
    lines = ["% synthetic PBN in the ASSA-PBN matlab text format", "% fixture of pbn_rl_b200 (tests/golden)", str(n),
             " ".join(map(str, n_funcs)), " ".join(map(str, arities))]
    tables = []
    for f, k in enumerate(arities):
        tt = [rng.randint(0, 1) for _ in range(1 << k)]
        if f == 2:
            tt = [1, 1]            # constant True  -> the reference's translate() special case
        if f == 7:
            tt = [0, 0]            # constant False
        tables.append(tt)
        lines.append(" ".join(map(str, tt)))
    for k in arities:
        lines.append(" ".join(map(str, rng.sample(range(n), k))))
    for g, nf in enumerate(n_funcs):
        w = [rng.randint(1, 5) for _ in range(nf)]
        lines.append(" ".join("%.4f" % (x / sum(w)) for x in w))
    lines += ["0.0025", "1"]
    path.write_text("\n".join(lines) + "\n")


def full_table(expr, genes):
    code = compile(expr.strip(), "<golden>", "eval")
    n = len(genes)
    t = 0
    for s in range(1 << n):
        env = {g: bool((s >> i) & 1) for i, g in enumerate(genes)}
        if eval(code, {}, env):
            t |= 1 << s
    return t


def fixed_inputs(n, count=4096):
    mask = (1 << n) - 1
    return [(((j + 1) * 0x9E3779B97F4A7C15) ^ ((j + 1) * 0xBF58476D1CE4E5B9 << 17)) & mask for j in range(count)]


def next_states(genes, exprs, inputs):
    codes = [compile(e.strip(), "<golden>", "eval") for e in exprs]   # eval() itself ignores leading blanks
    out = []
    for s in inputs:
        env = {g: bool((s >> i) & 1) for i, g in enumerate(genes)}
        t = 0
        for i, c in enumerate(codes):
            if eval(c, {}, env):
                t |= 1 << i
        out.append(t)
    return out


def digest(values, n):
    nb = (n + 7) // 8
    h = hashlib.sha256()
    for v in values:
        h.update(int(v).to_bytes(nb, "little"))
    return h.hexdigest()


def main():
    from sympy import symbols
    from sympy.logic import SOPform

    # ---- ASSA matlab format through the reference's parser
    assa = OUT / "assa_example.txt"
    write_assa_example(assa)
    code = slice_source(REF / "train_assa_matlab_BQN.py", "def translate", "genes = [f")
    ns = {"args": types.SimpleNamespace(assa_file=str(assa)), "np": np, "itertools": itertools, "defaultdict": defaultdict,
          "symbols": symbols, "SOPform": SOPform}
    exec(compile(code, "train_assa_matlab_BQN.py[sliced]", "exec"), ns)
    genes = ns["genes"]
    log_funcs = ns["log_funcs"]
    expected = {"genes": genes, "perturbation_rate": ns["perturbation_rate"],
                "logic_functions": {str(g): [[e, p] for e, p in log_funcs[g]] for g in sorted(log_funcs)},
                "full_tables": {str(g): [str(full_table(e, genes)) for e, _ in log_funcs[g]] for g in sorted(log_funcs)}}
    (OUT / "assa_example_expected.json").write_text(json.dumps(expected, indent=1) + "\n")

    # ---- bb33: the reference's ISPL parser on models/bb33/bb33.ispl, and the .bnet twin
    code = slice_source(REF / "train_assa_BQN.py", "with open(args.assa_file", "print(list(logic_funcs.keys()))")
    ns = {"args": types.SimpleNamespace(assa_file=str(REF / "models/bb33/bb33.ispl")), "defaultdict": defaultdict,
          "print": lambda *a, **k: None}
    exec(compile(code, "train_assa_BQN.py[sliced]", "exec"), ns)
    lf = ns["logic_funcs"]
    g33 = list(lf.keys())
    exprs = [lf[g][0][0] for g in g33]
    inputs = fixed_inputs(len(g33))
    nxt = next_states(g33, exprs, inputs)
    (OUT / "bb33.bnet").write_text((REF / "models/bb33/bb33.bnet").read_text())
    (OUT / "bb33_expected.json").write_text(json.dumps({
        "genes": g33, "python_exprs": exprs, "n_inputs": len(inputs), "sha256": digest(nxt, len(g33)),
        "first_rows": [[str(inputs[j]), str(nxt[j])] for j in range(8)],
        "max_arity": max(len({t for t in e.replace("(", " ").replace(")", " ").split() if t not in ("and", "or", "not")})
                         for e in exprs)}, indent=1) + "\n")

    # ---- the 14-gene control network of train_control_gbdq.py (wide predictor: MyoD1 has 8 inputs)
    tree = ast.parse((REF / "train_control_gbdq.py").read_text())
    call = next(nd for nd in ast.walk(tree) if isinstance(nd, ast.Call) and getattr(nd.func, "attr", "") == "make"
                and any(k.arg == "control_nodes" for k in nd.keywords))
    kw = {k.arg: ast.literal_eval(k.value) for k in call.keywords if k.arg in ("genes", "control_nodes", "logic_functions")}
    cg = kw["genes"]
    cexprs = [row[0][0] for row in kw["logic_functions"]]
    cin = fixed_inputs(len(cg))
    cn = next_states(cg, cexprs, cin)
    (OUT / "control14.json").write_text(json.dumps({
        "genes": cg, "control_nodes": kw["control_nodes"], "logic_functions": kw["logic_functions"],
        "n_inputs": len(cin), "sha256": digest(cn, len(cg)), "first_rows": [[str(cin[j]), str(cn[j])] for j in range(8)]},
        indent=1) + "\n")
    print("wrote", [p.name for p in (assa, OUT / "assa_example_expected.json", OUT / "bb33.bnet", OUT / "bb33_expected.json",
                                     OUT / "control14.json")])


if __name__ == "__main__":
    main()
