"""GPU: networks loaded from the other formats (ASSA matlab text, .bnet, inline logic functions with a wide
predictor, ControlPBNEnv) stepped by the kernels, against the reference-parser fixtures and the oracle."""
import hashlib
import json

import numpy as np
import pytest

from helpers import GOLD

pytestmark = pytest.mark.gpu


def _golden(name):
    return json.loads((GOLD / name).read_text())


def _inputs(n, count=4096):
    mask = (1 << n) - 1
    return [(((j + 1) * 0x9E3779B97F4A7C15) ^ ((j + 1) * 0xBF58476D1CE4E5B9 << 17)) & mask for j in range(count)]


def _digest(values, n):
    nb = (n + 7) // 8
    h = hashlib.sha256()
    for v in values:
        h.update(int(v).to_bytes(nb, "little"))
    return h.hexdigest()


def _step_fixed_inputs(net, kernel):
    import torch
    from pbn_rl_b200 import VecPBNEnv
    ins = _inputs(net.n_genes)
    env = VecPBNEnv(net, len(ins), None, device="cuda:0", kernel=kernel)
    env.set_state(torch.tensor(ins, dtype=torch.int64).reshape(-1, 1), packed=True)
    env.step(None)
    torch.cuda.synchronize()
    out = [int(v) for v in env.state.cpu().numpy()[:, 0]]
    k = env.kernel
    env.close()
    return out, k


@pytest.mark.parametrize("kernel", ["auto", "scalar"])
def test_bb33_bnet_known_answer(kernel):
    from pbn_rl_b200.formats import network_from_bnet
    want = _golden("bb33_expected.json")
    out, k = _step_fixed_inputs(network_from_bnet(GOLD / "bb33.bnet"), kernel)
    assert k == ("sliced" if kernel == "auto" else "scalar")
    assert _digest(out, 33) == want["sha256"]


def test_wide_predictor_known_answer():
    from pbn_rl_b200 import PBNNetwork
    want = _golden("control14.json")
    net = PBNNetwork.from_logic_functions(want["genes"], want["logic_functions"])
    out, k = _step_fixed_inputs(net, "auto")
    assert k == "scalar"                      # wide predictors run on the general kernel
    assert _digest(out, 14) == want["sha256"]
    assert [[str(a), str(b)] for a, b in zip(_inputs(14)[:8], out[:8])] == want["first_rows"]


def test_assa_network_rollout_bit_exact_with_selection_probabilities():
    """Product: truth tables from the ASSA file.  Oracle: the python expressions + probabilities the REFERENCE
    parser made of the same file.  Own-RNG rollout (non-uniform predictor selection) must agree bit for bit."""
    import torch
    from oracle import pbn_oracle as O
    from pbn_rl_b200 import VecPBNEnv
    from pbn_rl_b200.formats import network_from_assa_matlab
    want = _golden("assa_example_expected.json")
    net, rate = network_from_assa_matlab(GOLD / "assa_example.txt")
    onet = O.OracleNetwork(want["genes"], [[(e, p) for e, p in want["logic_functions"][str(g)]] for g in range(5)])
    e, seed, p = 4096, 1234, 0.05
    env = VecPBNEnv(net, e, None, device="cuda:0", perturb_p=p, perturb_mode="A", seed=seed)
    assert env.kernel == "scalar"
    rng = np.random.default_rng(0)
    state = rng.integers(0, 32, size=(e, 1)).astype(np.uint64)
    env.set_state(torch.from_numpy(state.astype(np.int64)), packed=True)
    ids = np.arange(e, dtype=np.uint64)
    tables = (np.zeros(1, np.int32), np.zeros((0, 1), np.uint64), np.zeros((0, 1), np.uint64))
    counts = np.zeros(3)
    for step in range(6):
        act = rng.integers(0, 6, size=(e, 3), dtype=np.uint8)
        env.step(torch.from_numpy(act).cuda())
        sel = O.scalar_stream_selection(onet, ids, step, seed)
        pert = O.scalar_stream_perturbation(5, p, ids, step, seed)
        state, *_ = O.batched_step(onet, tables, state, act, np.full(e, -1), np.zeros(e, np.uint16), horizon=0,
                                   mode=O.PERT_A, sel=sel, pert=pert, r_success=5.0, r_step=0.0, r_action=-1.0)
        assert np.array_equal(env.state.cpu().numpy().astype(np.uint64), state), step
        counts += np.bincount(sel[:, 2], minlength=3)
    freq = counts / counts.sum()                 # gene 2: probabilities 0.3 / 0.2 / 0.5
    assert np.allclose(freq, [0.3, 0.2, 0.5], atol=0.02)
    assert abs(rate - 0.0025) < 1e-12
    env.close()


def test_control_env_protocol():
    """train_control_gbdq.py:45-72 + control_gbdq_model/__init__.py:35,66-86,169: binary action per control node."""
    import torch
    from pbn_rl_b200 import make
    want = _golden("control14.json")
    env = make("gym-PBN/ControlPBNEnv", N=14, genes=want["genes"], control_nodes=want["control_nodes"],
               logic_functions=want["logic_functions"])
    assert len(env.control_nodes) == 8 and env.observation_space.shape[0] == 14
    (state, target), _ = env.reset()
    assert len(state) == 14 and len(env.all_attractors) >= 2
    inputs_only = [6, 7, 8, 10, 11, 12, 13]                 # FGF8, SHH, Pax3, Mef2c, Mef2a, ID3, WNT keep their value
    env.graph.setState([0] * 14)
    s, r, term, trunc, _ = env.step(torch.tensor([1, 0, 0, 0, 0, 0, 0, 0]))      # flip control node 6 = FGF8
    assert s[6] == 1 and all(s[g] == 0 for g in inputs_only if g != 6) and r in (-1.0, 4.0)
    s2, r2, *_ = env.step(torch.zeros(8, dtype=torch.int64))                      # no intervention
    assert s2[6] == 1 and r2 in (0.0, 5.0)
    s3, *_ = env.step([0, 0, 0, 0, 0, 0, 0, 1])                                   # node 14 does not exist: inert
    assert s3[6] == 1
    with pytest.raises(ValueError):
        env.step([1, 0, 0])
    env.close()


def test_pbnenv_from_reference_parser_output_finds_attractors_on_the_gpu():
    """train_assa_BQN.py:121-124: gym.make("gym-PBN/PBNEnv", N=, genes=, logic_functions=) with what the
    reference's ISPL parser produced for models/bb33/bb33.ispl.  N = 33: the attractor table comes from
    GPU rollouts + the device closure search; every attractor must be closed under the network's update."""
    import torch
    from pbn_rl_b200 import make
    want = _golden("bb33_expected.json")
    env = make("gym-PBN/PBNEnv", N=33, genes=want["genes"], logic_functions=[[(e, 1.0)] for e in want["python_exprs"]],
               min_attractors=1)
    assert len(env.all_attractors) >= 1 and sum(env.attractor_search["basin_fraction"]) > 0.9
    net = env.network
    for members in env.attractor_search["states"]:
        mset = set(members)
        for s in members[:64]:             # a Boolean network: exactly one successor per state
            assert net.next_state_int(s, [0] * 33) in mset
    (state, target), _ = env.reset()
    assert env.is_attracting_state(state) and len(target) == 33
    s, r, term, trunc, _ = env.step(torch.tensor([0, 0, 0]))
    assert env.is_attracting_state(s)      # no intervention: the state stays inside its attractor
    env.close()
