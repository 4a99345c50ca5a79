"""Host-side loaders against the golden fixtures derived from the reference tree."""
import hashlib
import itertools
import pickle

import numpy as np
import pytest

from helpers import GOLD, NETS, attractor_set, golden, k4_inputs, k4_selections, product_net
from pbn_rl_b200 import (AttractorSet, IsplError, PBNNetwork, compile_expression, find_attractors_stg,
                         load_attractor_pickle, logic_functions_from_ispl, pack_states, parse_ispl, render_ispl,
                         save_attractor_pickle, sorted_id_permutation, unpack_states)


@pytest.mark.parametrize("name", NETS)
def test_k1_writer_reproduces_reference_ispl_bytes(name):
    """Rendering the network back through the model_template.jj2 layout gives the reference file."""
    net = product_net(name)
    text = net.to_ispl()
    assert hashlib.sha256(text.encode()).hexdigest() == golden(f"{name}.json")["ispl_sha256"]
    again = PBNNetwork.from_ispl(text)
    assert again.genes == net.genes
    assert [[(f.inputs, f.lut) for f in fs] for fs in again.functions] == \
           [[(f.inputs, f.lut) for f in fs] for fs in net.functions]


def test_k1_user_supplied_jinja_template_matches_builtin_writer():
    jinja2 = pytest.importorskip("jinja2")
    del jinja2
    template = ("Agent M\n\tVars:\n\t\t{% for gene in log_funcs %}\n\t\tx{{ gene }}: boolean;\n\t\t{% endfor %}\n"
                "\tend Vars\n\tActions = {none};\n\tProtocol:\n\t\tOther: {none};\n\tend Protocol\n\tEvolution:\n"
                "\t\t{% for key in log_funcs %}\n\t\t{% for fun in log_funcs[key] %}\n"
                "\t\tx{{ key }}=true if ({{ fun }})=true;\n\t\tx{{ key }}=false if ({{ fun }})=false;\n"
                "\t\t{% endfor %}\n\t\t{% endfor %}\n\tend Evolution\nend Agent\n\nInitStates\n"
                "\t\tM.x234237=true or M.x234237=false;\nend InitStates\n\n")
    net = product_net("pbn7")
    assert net.to_ispl(template) == net.to_ispl()


@pytest.mark.parametrize("name", NETS)
def test_k4_truth_tables(name):
    net = product_net(name)
    n = net.n_genes
    x = k4_inputs(n)
    h = hashlib.sha256()
    for sel in k4_selections(n):
        for j in range(4096):
            s = int(x[j, 0]) | (int(x[j, 1]) << 64 if n > 64 else 0)
            t = net.next_state_int(s, sel[j])
            h.update(int(t & 0xFFFFFFFFFFFFFFFF).to_bytes(8, "little"))
            if n > 64:
                h.update(int(t >> 64).to_bytes(8, "little"))
    assert h.hexdigest() == golden("k4_transitions.json")[name]["sha256"]


@pytest.mark.parametrize("name", ["pbn7", "pbn10"])
def test_k5_stg_attractor_finder(name):
    attrs, info = find_attractors_stg(product_net(name))
    want = golden("k5_stg.json")[name]
    assert info["n_edges"] == want["n_edges"] and info["n_sccs"] == want["n_sccs"]
    assert info["sink_sccs"] == want["sink_sccs"]
    assert len(attrs) == len(want["sink_sccs"])


def test_k2_bittner7_pickle_equals_sink_sccs():
    attrs = attractor_set("pbn7")
    found, _ = find_attractors_stg(product_net("pbn7"))
    expanded = []
    for attr in attrs.attractors:
        states = set()
        for pat in attr:
            stars = [i for i, b in enumerate(pat) if b == "*"]
            for fill in itertools.product((0, 1), repeat=len(stars)):
                bits = [0 if b == "*" else int(b) for b in pat]
                for i, v in zip(stars, fill):
                    bits[i] = v
                states.add(tuple(bits))
        expanded.append(sorted(states))
    assert sorted(expanded) == sorted(sorted(tuple(s) for s in a) for a in found.attractors)


def test_k3_bittner28_permutation_and_packed_targets():
    net = product_net("pbn28")
    k3 = golden("k3_bittner28.json")
    assert sorted_id_permutation(net.genes) == k3["sorted_to_file_order"]
    _, _, val = attractor_set("pbn28").tables()
    assert [int(v) for v in val[:, 0]] == k3["targets_file_order"]
    assert hex(int(val[0, 0])) == "0xeddf7d7"


def test_attractor_pickle_roundtrip_reference_format(tmp_path):
    """list[list[tuple]] with python ints, numpy ints and '*' (SURVEY.md Appendix A)."""
    raw = [[(np.int64(1), 0, "*", 1)], [(0, 0, np.int64(1), 1), (1, 1, 1, 1)]]
    p = tmp_path / "a.pkl"
    with open(p, "wb") as f:
        pickle.dump(raw, f)
    attrs = load_attractor_pickle(p)
    assert attrs.attractors == [[(1, 0, "*", 1)], [(0, 0, 1, 1), (1, 1, 1, 1)]]
    offs, care, val = attrs.tables()
    assert offs.tolist() == [0, 1, 3] and care[:, 0].tolist() == [0b1011, 0b1111, 0b1111]
    assert val[:, 0].tolist() == [0b1001, 0b1100, 0b1111]
    assert attrs.contains(0, [1, 0, 1, 1]) and not attrs.contains(0, [0, 0, 1, 1])
    assert attrs.attractor_of([1, 1, 1, 1]) == 1 and attrs.attractor_of([0, 1, 0, 0]) == -1
    save_attractor_pickle(tmp_path / "b.pkl", attrs)
    assert load_attractor_pickle(tmp_path / "b.pkl").attractors == attrs.attractors
    with pytest.raises(ValueError):
        AttractorSet([[(1, 0)]], 4)
    with pytest.raises(ValueError):
        AttractorSet([[]], 4)


def test_pack_unpack_two_words():
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 2, size=(50, 70)).astype(np.uint8)
    w = pack_states(bits)
    assert w.shape == (50, 2) and (w[:, 1] >> np.uint64(6)).max() == 0
    assert np.array_equal(unpack_states(w, 70), bits)


def test_both_ispl_dialects_and_reference_parser_shape():
    compact = """Agent M
	Vars:
		v_A: boolean;
		v_B: boolean;
		v_C: boolean;
	end Vars
	Evolution:
		v_A=true  if ((v_A&v_B)|~v_C)=true;
		v_A=false if ((v_A&v_B)|~v_C)=false;
		v_B=true  if v_A=true;
		v_B=false if v_A=false;
		v_C=true  if (v_A=false & v_B)=true;
		v_C=false if (v_A=false & v_B)=false;
	end Evolution
end Agent
"""
    genes, funcs = parse_ispl(compact)
    assert genes == ["v_A", "v_B", "v_C"] and funcs["v_B"] == ["v_A"]
    net = PBNNetwork.from_ispl(compact)
    assert [f.arity for fs in net.functions for f in fs] == [3, 1, 2]
    # v_C = (not A) and B
    assert [net.functions[2][0]([a, b, 0]) for a in (0, 1) for b in (0, 1)] == [0, 1, 0, 0]
    g2, lf = logic_functions_from_ispl(compact)
    assert g2 == genes and lf[0][0][1] == 1.0 and " and " in lf[0][0][0] and " not " in lf[0][0][0]
    # the gym.make(genes=, logic_functions=) path accepts python-style expressions (train_assa_BQN.py:109)
    net2 = PBNNetwork.from_logic_functions(genes, [[(lf[0][0][0], 1.0)], [("v_A", 1.0)], [("( not v_A ) and v_B", 1.0)]])
    assert [[(f.inputs, f.lut) for f in fs] for fs in net2.functions] == \
           [[(f.inputs, f.lut) for f in fs] for fs in net.functions]
    # dict keyed by gene index (train_assa_matlab_BQN.py:171)
    net3 = PBNNetwork.from_logic_functions(genes, {0: [(lf[0][0][0], 1.0)], 1: [("v_A", 1.0)], 2: [("not v_A and v_B", 1)]})
    assert net3.functions[2][0].lut == net.functions[2][0].lut


def test_expression_compiler_edges():
    idx = {"a": 0, "b": 1, "c": 2, "d": 3, "e": 4}
    f = compile_expression("a & ~a | b", idx)
    assert f.inputs == (1,) and f.lut == 0b10          # support reduced to b
    assert compile_expression("true", idx).arity == 0 and compile_expression("a | ~a", idx).lut == 1
    g = compile_expression("(a and not b) or (c && !d) || e", idx)
    assert g.arity == 5
    for bits in itertools.product((0, 1), repeat=5):
        a, b, c, d, e = bits
        assert g(bits) == int((a and not b) or (c and not d) or e)
    for bad in ("a &", "(a | b", "a b", "a & zz", ")"):
        with pytest.raises(IsplError):
            compile_expression(bad, idx)
    with pytest.raises(IsplError):
        PBNNetwork.from_expressions(["a", "a"], [["a"], ["a"]])
    with pytest.raises(IsplError):
        PBNNetwork.from_expressions(["a"], [[]])
    with pytest.raises(IsplError):
        parse_ispl("Agent M\nend Agent\n")


def test_descriptor_arrays_and_probabilities():
    net = PBNNetwork.from_expressions(["a", "b"], [[("a | b", 0.2), ("a & b", 0.3), ("b", 0.5)], ["a"]])
    arr = net.descriptor_arrays()
    assert arr["func_offset"].tolist() == [0, 3, 4]
    cum = arr["func_cum"].tolist()
    assert abs(cum[0] / 2**32 - 0.2) < 1e-9 and abs(cum[1] / 2**32 - 0.5) < 1e-9 and cum[2] == 0xFFFFFFFF
    assert not net.is_uniform and product_net("pbn28").is_uniform
    assert net.adjacency() == [[0, 1], [0]]
