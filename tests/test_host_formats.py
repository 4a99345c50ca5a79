"""CPU: the other network formats (pbn_rl_b200/formats.py, wide predictors) against fixtures produced by
RUNNING the reference's own parsers (tests/golden/make_golden_formats.py)."""
import hashlib
import json

import numpy as np
import pytest

from helpers import GOLD


def _golden(name):
    return json.loads((GOLD / name).read_text())


def _full_table(f, n):
    t = 0
    for s in range(1 << n):
        if f([(s >> i) & 1 for i in range(n)]):
            t |= 1 << s
    return t


def test_assa_matlab_format_matches_the_reference_parser():
    from pbn_rl_b200.formats import network_from_assa_matlab
    want = _golden("assa_example_expected.json")
    net, rate = network_from_assa_matlab(GOLD / "assa_example.txt")
    assert net.genes == want["genes"] and rate == want["perturbation_rate"]
    for g in range(net.n_genes):
        ref = want["logic_functions"][str(g)]
        assert len(net.functions[g]) == len(ref)
        assert np.allclose(net.probabilities[g], [p for _, p in ref], atol=1e-4)
        for f, tab in zip(net.functions[g], want["full_tables"][str(g)]):
            assert _full_table(f, net.n_genes) == int(tab)
    assert not net.is_uniform          # real selection probabilities: the scalar kernel's thresholds carry them
    auto, _ = network_from_assa_matlab((GOLD / "assa_example.txt").read_text(), index_base="auto")
    assert [[f.lut for f in fs] for fs in auto.functions] == [[f.lut for f in fs] for fs in net.functions]


def test_assa_matlab_errors():
    from pbn_rl_b200 import IsplError
    from pbn_rl_b200.formats import parse_assa_matlab
    text = (GOLD / "assa_example.txt").read_text().splitlines()
    with pytest.raises(IsplError):
        parse_assa_matlab("\n".join(text[:9]))                       # truncated
    bad = list(text)
    bad[5] = "1 0 0"                                                  # 3 entries for 3 predictors
    with pytest.raises(IsplError):
        parse_assa_matlab("\n".join(bad))
    with pytest.raises(IsplError):
        parse_assa_matlab("\n".join(text), index_base=1)              # 0-based file read as 1-based: index -1


def _digest(values, n):
    nb = (n + 7) // 8
    h = hashlib.sha256()
    for v in values:
        h.update(int(v).to_bytes(nb, "little"))
    return h.hexdigest()


def _inputs(n, count=4096):
    mask = (1 << n) - 1
    return [(((j + 1) * 0x9E3779B97F4A7C15) ^ ((j + 1) * 0xBF58476D1CE4E5B9 << 17)) & mask for j in range(count)]


def test_bnet_and_ispl_of_bb33_agree_with_the_reference_parser():
    """models/bb33/bb33.bnet and bb33.ispl are the same network; the reference's ISPL parser output
    (python expressions, train_assa_BQN.py:51-109) pins the next states."""
    from pbn_rl_b200 import PBNNetwork
    from pbn_rl_b200.formats import network_from_bnet
    want = _golden("bb33_expected.json")
    net = network_from_bnet(GOLD / "bb33.bnet")
    assert net.genes == want["genes"] and net.n_genes == 33 and net.max_arity == want["max_arity"]
    ins = _inputs(33)
    nxt = [net.next_state_int(s, [0] * 33) for s in ins]
    assert _digest(nxt, 33) == want["sha256"]
    assert [[str(a), str(b)] for a, b in zip(ins[:8], nxt[:8])] == want["first_rows"]
    # the reference's python-expression form (what it hands to gym.make) loads to the same truth tables
    net2 = PBNNetwork.from_logic_functions(want["genes"], [[(e, 1.0)] for e in want["python_exprs"]])
    assert [f[0].lut for f in net2.functions] == [f[0].lut for f in net.functions]


def test_control_network_with_wide_predictor():
    """train_control_gbdq.py:45-72: MyoD1 has 8 inputs -> a wide predictor (multi-word truth table)."""
    from pbn_rl_b200 import PBNNetwork
    want = _golden("control14.json")
    net = PBNNetwork.from_logic_functions(want["genes"], want["logic_functions"])
    assert net.max_arity == 8
    ins = _inputs(14)
    nxt = [net.next_state_int(s, [0] * 14) for s in ins]
    assert _digest(nxt, 14) == want["sha256"]
    arr = net.descriptor_arrays()
    assert arr["wide_inputs"].shape == (1, 16) and arr["wide_lut_offset"].tolist() == [0, 4]
    assert int(arr["func_arity"].max()) == 8 and arr["wide_lut"].shape == (4,)


def test_wide_networks_specialise_up_to_12_inputs():
    """Predictors of 7..12 inputs get Shannon-expanded LOP3 trees in the generated source (bit-sliced kernels);
    beyond 12 inputs the library refuses to specialise (thread-per-env kernel only)."""
    from pbn_rl_b200 import PBNNetwork, _cabi
    from pbn_rl_b200.vec_env import jit_source
    want = _golden("control14.json")
    net = PBNNetwork.from_logic_functions(want["genes"], want["logic_functions"])
    src = jit_source(net)
    assert "pbn_update_part" in src and "bmux(" in src
    genes = ["v%d" % i for i in range(14)]
    big = PBNNetwork.from_expressions(genes, [[" & ".join(genes[:13])]] + [[g] for g in genes[1:]])
    assert big.max_arity == 13
    with pytest.raises(_cabi.PbnError) as ei:
        jit_source(big)
    assert "12 inputs" in str(ei.value)
