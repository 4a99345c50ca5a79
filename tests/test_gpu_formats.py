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


@pytest.mark.parametrize("kernel", ["auto", "scalar"])
def test_wide_predictor_known_answer(kernel):
    from pbn_rl_b200 import PBNNetwork
    want = _golden("control14.json")
    net = PBNNetwork.from_logic_functions(want["genes"], want["logic_functions"])
    out, k = _step_fixed_inputs(net, kernel)
    assert k == ("sliced" if kernel == "auto" else "scalar")   # 8-input predictor: Shannon-expanded LOP3 tree
    assert _digest(out, 14) == want["sha256"]
    assert [[str(a), str(b)] for a, b in zip(_inputs(14)[:8], out[:8])] == want["first_rows"]


@pytest.mark.parametrize("kernel", ["auto", "scalar"])
def test_assa_network_rollout_bit_exact_with_selection_probabilities(kernel):
    """Product: truth tables from the ASSA file.  Oracle: the python expressions + probabilities the REFERENCE
    parser made of the same file.  Own-RNG rollout (non-uniform predictor selection) must agree bit for bit -- on the
    bit-sliced kernel (threshold comparison of a bit-sliced 32-bit uniform, train_assa_matlab_BQN.py:72-171) and on
    the thread-per-env kernel -- and the selection frequencies must follow the file's probabilities (chi-square)."""
    import torch
    from oracle import pbn_oracle as O
    from pbn_rl_b200 import VecPBNEnv
    from pbn_rl_b200.formats import network_from_assa_matlab
    want = _golden("assa_example_expected.json")
    net, rate = network_from_assa_matlab(GOLD / "assa_example.txt")
    onet = O.OracleNetwork(want["genes"], [[(e, p) for e, p in want["logic_functions"][str(g)]] for g in range(5)])
    e, seed, p = 8192, 1234, 0.05
    env = VecPBNEnv(net, e, None, device="cuda:0", perturb_p=p, perturb_mode="A", seed=seed, kernel=kernel)
    assert env.kernel == ("sliced" if kernel == "auto" else "scalar")
    rng = np.random.default_rng(0)
    state = rng.integers(0, 32, size=(e, 1)).astype(np.uint64)
    env.set_state(torch.from_numpy(state.astype(np.int64)), packed=True)
    ids = np.arange(e, dtype=np.uint64)
    tables = (np.zeros(1, np.int32), np.zeros((0, 1), np.uint64), np.zeros((0, 1), np.uint64))
    counts = np.zeros(3)
    for step in range(6):
        act = rng.integers(0, 6, size=(e, 3), dtype=np.uint8)
        env.step(torch.from_numpy(act).cuda())
        if env.kernel == "sliced":
            sel, pert = O.sliced_stream(onet, p, ids, step, seed)
        else:
            sel = O.scalar_stream_selection(onet, ids, step, seed)
            pert = O.scalar_stream_perturbation(5, p, ids, step, seed)
        state, *_ = O.batched_step(onet, tables, state, act, np.full(e, -1), np.zeros(e, np.uint16), horizon=0,
                                   mode=O.PERT_A, sel=sel, pert=pert, r_success=5.0, r_step=0.0, r_action=-1.0)
        assert np.array_equal(env.state.cpu().numpy().astype(np.uint64), state), step
        counts += np.bincount(sel[:, 2], minlength=3)
    probs = np.array([0.3, 0.2, 0.5])               # gene 2 of the file
    chi2 = float((((counts - counts.sum() * probs) ** 2) / (counts.sum() * probs)).sum())
    assert chi2 < 18.4, (chi2, counts)              # 2 degrees of freedom, p = 1e-4
    assert abs(rate - 0.0025) < 1e-12
    env.close()


def test_weighted_network_resident_equals_rows():
    """A network mixing K = 1..4 with uniform and weighted selection: plane-resident steps == row-format steps."""
    import torch
    from pbn_rl_b200 import AttractorSet, PBNNetwork, VecPBNEnv
    genes = ["g%d" % i for i in range(6)]
    fs = [[("g1 | g2", 0.9), ("g1 & g2", 0.1)], ["g0"], ["g3 & g4", "g3 | g4", "~g5"],
          [("g0", 0.25), ("g1", 0.25), ("g2 & ~g3", 0.5)], ["g4", "~g4"], [("g0", 0.1), ("g1", 0.2), ("g2", 0.3), ("g3", 0.4)]]
    net = PBNNetwork.from_expressions(genes, fs)
    attrs = AttractorSet([[(1, 1, 1, 1, 1, 1)], [(0, 0, 0, 0, 0, 0)], [(1, 0, 1, 0, 1, 0)]], 6)
    e = 2048 + 100
    envs = [VecPBNEnv(net, e, attrs, device="cuda:0", perturb_p=0.01, perturb_mode="B", seed=5, horizon=4, auto_reset=True,
                      kernel="sliced", resident=r) for r in (False, True)]
    g = torch.Generator(device="cuda").manual_seed(1)
    st = torch.randint(0, 64, (e, 1), generator=g, device="cuda", dtype=torch.int64)
    tg = torch.randint(0, 3, (e,), generator=g, device="cuda", dtype=torch.int32)
    for env in envs:
        env.set_state(st, packed=True)
        env.set_target(tg)
    for step in range(10):
        a = torch.randint(0, 7, (e, 3), generator=g, device="cuda", dtype=torch.uint8)
        for env in envs:
            env.step(a)
        assert torch.equal(envs[0].reward, envs[1].reward) and torch.equal(envs[0].terminated, envs[1].terminated)
    assert torch.equal(envs[0].state, envs[1].state) and envs[0].stats() == envs[1].stats()
    for env in envs:
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
