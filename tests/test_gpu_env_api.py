"""GPU: the per-instance gym-style shim (the reference's call protocol) and distribution tests of the
kernels' own random streams against exact probabilities and against the CPU oracle env."""
import pickle

import numpy as np
import pytest

from helpers import attractor_set, golden, oracle_net, product_net

pytestmark = pytest.mark.gpu


@pytest.fixture()
def pbn7_root(tmp_path):
    """A working directory laid out like the fork's CWD: kaban/pbn7.ispl + data/attractors_Bittner-7.pkl."""
    (tmp_path / "kaban").mkdir()
    (tmp_path / "data").mkdir()
    (tmp_path / "kaban" / "pbn7.ispl").write_text(product_net("pbn7").to_ispl())
    raw = [[tuple(s) for s in a] for a in golden("attractors_bittner7.json")["attractors"]]
    with open(tmp_path / "data" / "attractors_Bittner-7.pkl", "wb") as f:
        pickle.dump(raw, f)
    return tmp_path


def test_bdq_training_protocol(pbn7_root):
    """The loop of bdq_model/__init__.py:161-213 against the shim."""
    import torch
    from pbn_rl_b200 import make
    env = make("gym-PBN/BittnerMultiGeneral", N=7, horizon=20, min_attractors=4, root=pbn7_root)
    assert env.observation_space.shape[0] == 7 and env.action_space.n == 8
    assert len(env.all_attractors) == 4 and len(env.attracting_states) == 4
    assert env.env.env.env is env and env.unwrapped is env
    (state, target), info = env.reset()
    assert len(state) == 7 and len(target) == 7 and env.is_attracting_state(state)
    assert env.state_attractor_id != env.target_attractor_id
    np.stack((state, target))  # what the agent does with them
    episodes = 0
    for frame in range(120):
        action = torch.randint(0, 8, (3,))
        env_action = list(action.unique())            # list of 0-d tensors, as in training
        new_state, reward, terminated, truncated, infos = env.step(env_action)
        assert isinstance(reward, float) and isinstance(terminated, bool) and len(new_state) == 7
        assert terminated == env.in_target(new_state)
        if terminated | truncated:
            env.rework_probas(env.n_steps)
            (new_state, target), _ = env.reset()
            assert env.n_steps == 0
            episodes += 1
    assert episodes >= 5
    env.close()


def test_model_tester_protocol(pbn7_root):
    """model_tester.py:595-648: setState / setTarget / in_target / render, raw action tensors, '*' -> 0."""
    import torch
    from pbn_rl_b200 import make
    env = make("gym-PBN/BittnerMultiGeneral", N=7, min_attractors=4, root=pbn7_root)
    all_attractors = env.all_attractors
    for src, tgt in ((0, 1), (1, 0), (3, 2)):
        initial = [0 if b == "*" else b for b in all_attractors[src][0]]
        env.reset()
        env.graph.setState(initial)
        env.setTarget(all_attractors[tgt])
        assert env.render() == tuple(initial)
        assert env.in_target(env.render()) == (src == tgt)
        state, *_ = env.step(torch.tensor([0, 0, 0]))   # raw tensor with duplicates = no intervention
        assert state == env.render()
    # singleton attractors are fixed points under every predictor choice (fixture K2)
    env.graph.setState(list(all_attractors[1][0]))
    env.setTarget(1)
    for _ in range(5):
        state, reward, term, trunc, _ = env.step([])     # graph_classifier/__init__.py:148
        assert term and state == tuple(all_attractors[1][0]) and reward == 5.0
    s2, r2, *_ = env.step(3)                             # ddqn_per/__init__.py:354: a bare int
    assert r2 in (-1.0, 4.0)
    assert env.graph.getNodeByID("x25485").index == 0 and env.graph.getNodeByID(25485).index == 0
    assert env.graph.get_adj_list()[0] == sorted(set(env.graph.nodes[0].predictors[0][0]) |
                                                 set(env.graph.nodes[0].predictors[2][0]))
    stg = env.graph.genSTG()
    assert len(stg) == 128 and abs(sum(stg[(1, 0, 1, 0, 1, 0, 1)][1].values()) - 1.0) < 1e-12
    with pytest.raises(ValueError):
        env.step([9])
    env.close()


def test_pbnenv_from_logic_functions_and_growing_attractor_table():
    """gym.make("gym-PBN/PBNEnv", genes=, logic_functions=) (train_assa_BQN.py:121-124)."""
    from pbn_rl_b200 import make
    genes = ["a", "b", "c"]
    lf = [[("a or b", 1.0)], [("a and c", 1.0)], [("not a", 1.0)]]
    env = make("gym-PBN/PBNEnv", N=3, genes=genes, logic_functions=lf, horizon=5)
    assert env.vec.kernel in ("sliced", "scalar")
    n0 = len(env.all_attractors)
    assert n0 >= 1
    k = env.add_attractor([(0, 1, 0)])
    assert len(env.all_attractors) == n0 + 1
    env.setTarget(k)
    env.graph.setState([0, 1, 0])
    assert env.in_target(env.render())
    env.close()


def _exact_next_distribution(net, s1):
    """Exact next-state distribution of the perturbation-free update from state s1."""
    dist = {0: 1.0}
    for i, (fs, ps) in enumerate(zip(net.functions, net.probabilities)):
        p1 = sum(p for f, p in zip(fs, ps) if f([(s1 >> g) & 1 for g in range(net.n_genes)]))
        nxt = {}
        for t, pr in dist.items():
            if p1 > 0:
                nxt[t | (1 << i)] = nxt.get(t | (1 << i), 0) + pr * p1
            if p1 < 1:
                nxt[t] = nxt.get(t, 0) + pr * (1 - p1)
        dist = nxt
    return dist


@pytest.mark.parametrize("kernel", ["scalar", "sliced"])
def test_next_state_distribution_chi_square(kernel):
    """Own-RNG mode, perturbation off: the empirical next-state histogram over 2^17 envs that share one
    start state matches the exact PBN transition probabilities (chi-square, df <= 127)."""
    import torch
    from pbn_rl_b200 import VecPBNEnv
    net = product_net("pbn7")
    e = 1 << 17
    for s0, act in ((0b0010101, [0, 0, 0]), (0b1000110, [2, 0, 5])):
        env = VecPBNEnv(net, e, attractor_set("pbn7"), device="cuda:0", kernel=kernel, seed=1234 + s0)
        env.state.fill_(s0)
        env.step(torch.tensor([act] * e, dtype=torch.uint8))
        counts = torch.bincount(env.state[:, 0], minlength=128).cpu().numpy().astype(float)
        s1 = s0
        for a in act:
            if a:
                s1 ^= 1 << (a - 1)
        exact = _exact_next_distribution(net, s1)
        assert counts[[t for t in range(128) if t not in exact]].sum() == 0
        chi2 = sum((counts[t] - e * p) ** 2 / (e * p) for t, p in exact.items())
        df = len(exact) - 1
        assert chi2 < df + 6 * np.sqrt(2 * df) + 10, (kernel, chi2, df)
        env.close()


@pytest.mark.parametrize("kernel", ["scalar", "sliced"])
@pytest.mark.parametrize("mode", ["A", "B", "C"])
def test_perturbation_rate_and_mode_statistics(kernel, mode):
    """Perturbed-gene counts follow Binomial(N*E, p); in mode A a perturbed env keeps s1 XOR pert."""
    import torch
    from pbn_rl_b200 import VecPBNEnv
    net = product_net("pbn28")
    e, p = 1 << 16, 0.01
    env = VecPBNEnv(net, e, attractor_set("pbn28"), device="cuda:0", kernel=kernel, perturb_p=p, perturb_mode=mode)
    fixed = int(attractor_set("pbn28").tables()[2][0, 0])
    env.state.fill_(fixed)
    env.step(None)
    n_pert = env.stats()["perturbed"]
    mean = e * 28 * p
    assert abs(n_pert - mean) < 6 * np.sqrt(mean)
    if mode == "A":
        changed = (env.state[:, 0] != fixed)
        # P(env has >= 1 perturbed gene) = 1 - (1-p)^28; the fixed point only moves when perturbed... or by its free genes
        frac = 1 - (1 - p) ** 28
        assert changed.float().mean().item() >= frac * 0.8
    env.close()


def test_attractor_hit_rate_matches_oracle_env():
    """Random-action rollouts from reset: the fraction of episodes that reach the target within the
    horizon agrees between the GPU env (own Philox streams) and the CPU oracle env (python random)."""
    import torch
    from oracle import pbn_oracle as O
    from pbn_rl_b200 import VecPBNEnv
    name, horizon = "pbn7", 12
    attrs = attractor_set(name)
    e = 1 << 15
    env = VecPBNEnv(product_net(name), e, attrs, device="cuda:0", horizon=horizon, auto_reset=True, bins=1,
                    perturb_p=0.0, seed=99)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    for _ in range(5 * horizon):
        env.step(torch.randint(0, 8, (e, 1), generator=g, device="cuda", dtype=torch.uint8))
    st = env.stats()
    gpu_rate = st["terminated"] / st["episodes"]
    onet = oracle_net(name)
    oenv = O.OraclePBNEnv(onet, attrs.attractors, horizon=horizon, perturb_p=0.0, seed=7)
    rng = np.random.default_rng(11)
    term = eps = 0
    oenv.reset()
    for _ in range(40000):
        _, _, t, tr, _ = oenv.step([int(rng.integers(0, 8))])
        if t or tr:
            term += int(t)
            eps += 1
            oenv.reset()
    cpu_rate = term / eps
    sigma = np.sqrt(cpu_rate * (1 - cpu_rate) / eps + gpu_rate * (1 - gpu_rate) / st["episodes"])
    assert abs(gpu_rate - cpu_rate) < 5 * sigma + 0.005, (gpu_rate, cpu_rate, sigma)
    env.close()


def test_cuda_graph_replay_advances_the_streams():
    """device_counter=True: a captured step replayed twice draws different randomness each time and
    equals two eager steps with host counters 0 and 1."""
    import torch
    from pbn_rl_b200 import VecPBNEnv
    net, attrs = product_net("pbn28"), attractor_set("pbn28")
    e = 4096
    a = VecPBNEnv(net, e, attrs, device="cuda:0", perturb_p=0.01, device_counter=True)
    b = VecPBNEnv(net, e, attrs, device="cuda:0", perturb_p=0.01)
    g = torch.Generator(device="cuda").manual_seed(0)
    s0 = torch.randint(0, 1 << 28, (e, 1), generator=g, device="cuda", dtype=torch.int64)
    acts = torch.randint(0, 29, (e, 3), generator=g, device="cuda", dtype=torch.uint8)
    for env in (a, b):
        env.state.copy_(s0)
        env.set_target(3)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        a.step(acts)                         # warm-up launch (counter 0 -> 1)
        a.state.copy_(s0)
        a.t.zero_()
        a.step_ctr_dev.zero_()
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            a.step(acts)
        a.state.copy_(s0)
        a.t.zero_()
        a.step_ctr_dev.zero_()
        graph.replay()
        stream.synchronize()
        first = a.state.clone()
        graph.replay()
        stream.synchronize()
    b.step(acts)
    assert torch.equal(first, b.state)
    b.step(acts)
    assert torch.equal(a.state, b.state)
    assert int(a.step_ctr_dev.item()) == 2


def test_pdl_graph_sequence_matches_eager_steps():
    """pdl=True: steps launched with programmatic stream serialisation (selection planes drawn before
    griddepcontrol.wait) inside a CUDA graph give exactly the states of plain eager steps."""
    import torch
    from pbn_rl_b200 import VecPBNEnv
    net, attrs = product_net("pbn28"), attractor_set("pbn28")
    e, steps = 8192, 5
    a = VecPBNEnv(net, e, attrs, device="cuda:0", perturb_p=0.01, pdl=True, auto_reset=True)
    b = VecPBNEnv(net, e, attrs, device="cuda:0", perturb_p=0.01, auto_reset=True)
    g = torch.Generator(device="cuda").manual_seed(0)
    s0 = torch.randint(0, 1 << 28, (e, 1), generator=g, device="cuda", dtype=torch.int64)
    acts = [torch.randint(0, 29, (e, 3), generator=g, device="cuda", dtype=torch.uint8) for _ in range(steps)]
    tgt = torch.randint(0, 14, (e,), generator=g, device="cuda", dtype=torch.int32)
    for env in (a, b):
        env.state.copy_(s0)
        env.set_target(tgt)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        a.step(acts[0])                      # warm-up (loads the module before capture)
        a.advance_counter()
        stream.synchronize()
        a.state.copy_(s0)
        a.set_target(tgt)
        a.t.zero_()
        a.step_ctr_dev.zero_()
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for k in range(steps):
                a.step(acts[k])
            a.advance_counter()
        a.state.copy_(s0)
        a.set_target(tgt)
        a.t.zero_()
        a.step_ctr_dev.zero_()
        graph.replay()
        graph.replay()
        stream.synchronize()
    for rep in range(2):
        for k in range(steps):
            b.step(acts[k])
    assert torch.equal(a.state, b.state) and torch.equal(a.target_id, b.target_id) and torch.equal(a.t, b.t)
    assert int(a.step_ctr_dev.item()) == 2 * steps
