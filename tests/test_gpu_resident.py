"""GPU parity of the plane-resident step (pbn_step with args->resident, csrc/step_planes.cuh): against the CPU
oracle with injected randomness, against the oracle's Philox twins in own-RNG mode, and bit for bit against the
row-format kernel (same random streams) over multi-step rollouts with perturbations, auto-reset and statistics."""
import hashlib
import os
import zlib

import numpy as np
import pytest

from helpers import NETS, attractor_set, golden, k4_inputs, k4_selections, oracle_net, product_net, random_case

pytestmark = pytest.mark.gpu

KW = dict(horizon=20, r_success=5.0, r_step=-0.25, r_action=-1.0)
MODES = {"none": 0, "A": 1, "B": 2, "C": 3}


def _env(name, e, mode="A", p=0.0, resident=True, attrs="default", **extra):
    from pbn_rl_b200 import VecPBNEnv
    kw = dict(KW)
    kw.update(extra)
    a = attractor_set(name) if attrs == "default" else attrs
    return VecPBNEnv(product_net(name), e, a, device="cuda:0", perturb_p=p, perturb_mode=mode, seed=0x5EED,
                     kernel="sliced", resident=resident, **kw)


def _load(env, case):
    import torch
    env.set_state(torch.from_numpy(case["state"].astype(np.int64)), packed=True)
    env.set_target(torch.from_numpy(case["target"]))
    env.t.copy_(torch.from_numpy(case["t"].astype(np.int16)))


def _oracle_step(name, case, sel, pert, mode, horizon=20):
    from oracle import pbn_oracle as O
    onet = oracle_net(name)
    tables = O.attractor_tables(attractor_set(name).attractors, onet.n)
    kw = dict(KW)
    kw["horizon"] = horizon
    return O.batched_step(onet, tables, case["state"], case["actions"], case["target"], case["t"],
                          mode=MODES[mode], sel=sel, pert=pert, **kw)


def _compare(env, expect, tag):
    import torch
    torch.cuda.synchronize()
    nxt, t1, rew, term, trunc = expect
    assert np.array_equal(env.reward.cpu().numpy().view(np.uint32), rew.view(np.uint32)), tag + ": reward"
    assert np.array_equal(env.terminated.cpu().numpy(), term), tag + ": terminated"
    assert np.array_equal(env.truncated.cpu().numpy(), trunc), tag + ": truncated"
    assert np.array_equal(env.state.cpu().numpy().astype(np.uint64), nxt), tag + ": state"
    assert np.array_equal(env.t.cpu().numpy().astype(np.uint16), t1), tag + ": t"


@pytest.mark.parametrize("name", NETS)
@pytest.mark.parametrize("e", [1, 1000, 4096, 5000])
def test_import_export_roundtrip(name, e):
    import torch
    case = random_case(name, e, seed=e + len(name))
    case["target"][::7] = -1
    env2 = _env(name, e)
    _load(env2, case)
    env2.set_target(torch.from_numpy(case["target"]))
    env2._planes()             # import only
    env2._rows_fresh = False
    assert np.array_equal(env2.state.cpu().numpy().astype(np.uint64), case["state"])
    assert np.array_equal(env2.target_id.cpu().numpy(), case["target"])
    assert np.array_equal(env2.t.cpu().numpy().astype(np.uint16), case["t"])


@pytest.mark.parametrize("name", NETS)
def test_k4_known_answer_resident(name):
    import torch
    from pbn_rl_b200 import VecPBNEnv
    net = product_net(name)
    x = k4_inputs(net.n_genes)
    h = hashlib.sha256()
    env = VecPBNEnv(net, 4096, None, device="cuda:0", perturb_mode="none", horizon=0, kernel="sliced", resident=True)
    for sel in k4_selections(net.n_genes):
        env.set_state(torch.from_numpy(x.astype(np.int64)), packed=True)
        env.step_injected(None, torch.from_numpy(sel))
        torch.cuda.synchronize()
        h.update(env.state.cpu().numpy().astype("<u8").tobytes())
    assert h.hexdigest() == golden("k4_transitions.json")[name]["sha256"]


@pytest.mark.parametrize("warps", ["4", "8"])
@pytest.mark.parametrize("name", NETS)
@pytest.mark.parametrize("mode", ["none", "A", "B", "C"])
def test_injected_step_bit_exact(name, mode, warps, monkeypatch):
    import torch
    monkeypatch.setenv("PBN_B200_PLANES_WARPS", warps)
    e = 4096
    case = random_case(name, e, seed=zlib.crc32((name + mode).encode()) & 0xFFFF)
    env = _env(name, e, mode=mode)
    _load(env, case)
    env.step_injected(torch.from_numpy(case["actions"]), torch.from_numpy(case["sel"]),
                      torch.from_numpy(case["pert"].astype(np.int64)))
    _compare(env, _oracle_step(name, case, case["sel"], case["pert"], mode), f"{name}/{mode}/w{warps}")


@pytest.mark.parametrize("e", [1, 31, 33, 1025, 3000])
def test_ragged_sizes(e):
    import torch
    from oracle import pbn_oracle as O
    name = "pbn28"
    case = random_case(name, e, seed=e)
    env = _env(name, e, mode="A")
    _load(env, case)
    env.step_injected(torch.from_numpy(case["actions"]), torch.from_numpy(case["sel"]),
                      torch.from_numpy(case["pert"].astype(np.int64)))
    _compare(env, _oracle_step(name, case, case["sel"], case["pert"], "A"), f"E={e}")
    env2 = _env(name, e, mode="A", p=0.02)
    _load(env2, case)
    env2.step(torch.from_numpy(case["actions"]).cuda())
    sel, pert = O.sliced_stream(oracle_net(name), 0.02, np.arange(e, dtype=np.uint64), 0, 0x5EED)
    _compare(env2, _oracle_step(name, case, sel, pert, "A"), f"E={e}/philox")


@pytest.mark.parametrize("name", NETS)
@pytest.mark.parametrize("mode", ["A", "B", "C"])
def test_own_rng_matches_oracle_twin(name, mode):
    import torch
    from oracle import pbn_oracle as O
    e = 2048 + 77
    case = random_case(name, e, seed=11)
    env = _env(name, e, mode=mode, p=0.01)
    _load(env, case)
    cur = dict(case)
    rng = np.random.default_rng(5)
    for step in range(3):
        acts = rng.integers(0, env.n_genes + 1, size=(e, 3), dtype=np.uint8)
        cur["actions"] = acts
        env.step(torch.from_numpy(acts).cuda())
        sel, pert = O.sliced_stream(oracle_net(name), 0.01, np.arange(e, dtype=np.uint64), step, 0x5EED)
        exp = _oracle_step(name, cur, sel, pert, mode)
        _compare(env, exp, f"{name}/{mode}/step{step}")
        cur["state"], cur["t"] = exp[0], exp[1]


@pytest.mark.parametrize("warps", ["4", "8"])
@pytest.mark.parametrize("name", NETS)
def test_resident_equals_row_kernel_over_rollout(name, warps, monkeypatch):
    """Same seeds, perturbations, auto-reset, statistics: every output of every step and the final env state are
    bit-identical between the plane-resident kernel and the row-format kernel."""
    import torch
    monkeypatch.setenv("PBN_B200_PLANES_WARPS", warps)
    e = 3 * 1024 + 500
    case = random_case(name, e, seed=3)
    envs = [_env(name, e, mode="A", p=0.01, resident=r, auto_reset=True, horizon=6) for r in (True, False)]
    for env in envs:
        _load(env, case)
    g = torch.Generator(device="cuda").manual_seed(9)
    for step in range(12):
        acts = torch.randint(0, envs[0].n_genes + 1, (e, 3), generator=g, device="cuda", dtype=torch.uint8)
        outs = []
        for env in envs:
            env.step(acts)
            outs.append((env.reward.clone(), env.terminated.clone(), env.truncated.clone(), env.source_id.clone()))
        for x, y in zip(*outs):
            assert torch.equal(x, y), f"{name}: output differs at step {step}"
    torch.cuda.synchronize()
    assert torch.equal(envs[0].state, envs[1].state)
    assert torch.equal(envs[0].target_id, envs[1].target_id)
    assert torch.equal(envs[0].t, envs[1].t)
    s0, s1 = envs[0].stats(), envs[1].stats()
    assert s0 == s1 and s0["episodes"] > 0 and s0["perturbed"] > 0, (s0, s1)


def test_resident_sharding_invariance():
    """A batch stepped as one resident env and as two shards with env_offset gives the same results."""
    import torch
    name, e = "pbn28", 4096
    case = random_case(name, e, seed=21)
    whole = _env(name, e, mode="A", p=0.01, auto_reset=True, horizon=5)
    _load(whole, case)
    halves = []
    for k in range(2):
        h = _env(name, e // 2, mode="A", p=0.01, auto_reset=True, horizon=5, env_offset=k * (e // 2))
        sub = {key: v[k * (e // 2):(k + 1) * (e // 2)] for key, v in case.items()}
        _load(h, sub)
        halves.append(h)
    g = torch.Generator(device="cuda").manual_seed(2)
    for _ in range(8):
        acts = torch.randint(0, 29, (e, 3), generator=g, device="cuda", dtype=torch.uint8)
        whole.step(acts)
        for k, h in enumerate(halves):
            h.step(acts[k * (e // 2):(k + 1) * (e // 2)].contiguous())
    assert torch.equal(whole.state, torch.cat([h.state for h in halves]))
    assert torch.equal(whole.target_id, torch.cat([h.target_id for h in halves]))
    assert torch.equal(whole.reward, torch.cat([h.reward for h in halves]))


def test_resident_graph_pdl_sequence_matches_eager():
    """CUDA graph of PDL launches on a resident env == eager resident steps with the host counter."""
    import torch
    name, e = "pbn28", 8192
    case = random_case(name, e, seed=33)
    acts = [torch.randint(0, 29, (e, 3), device="cuda", dtype=torch.uint8) for _ in range(6)]
    ref = _env(name, e, mode="A", p=0.01, auto_reset=True)
    _load(ref, case)
    for a in acts:
        ref.step(a)
    from pbn_rl_b200 import VecPBNEnv
    env = VecPBNEnv(product_net(name), e, attractor_set(name), device="cuda:0", perturb_p=0.01, perturb_mode="A",
                    seed=0x5EED, kernel="sliced", resident=True, auto_reset=True, pdl=True, **KW)
    _load(env, case)
    env._planes()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for a in acts[:3]:
                env.step(a)
            env.advance_counter()
        # capture does not execute: run the 3-step graph, then a second graph with the other actions
        graph.replay()
        graph2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph2, stream=stream):
            for a in acts[3:]:
                env.step(a)
            env.advance_counter()
        graph2.replay()
    torch.cuda.synchronize()
    assert torch.equal(env.state, ref.state)
    assert torch.equal(env.t, ref.t)
    assert torch.equal(env.reward, ref.reward)


@pytest.mark.parametrize("name", ["pbn28", "pbn70", "pbn10"])
def test_chained_sequence_matches_eager(name):
    """PBN_STEP_CHAIN: graphs of tile-chained launches on ONE env batch (every step really depends on the previous
    one, tile by tile) == eager resident steps with the host counter, over several replays."""
    import torch
    from pbn_rl_b200 import VecPBNEnv
    e = 40 * 1024 + 300
    case = random_case(name, e, seed=41)
    n = product_net(name).n_genes
    g = torch.Generator(device="cuda").manual_seed(4)
    acts = [torch.randint(0, n + 1, (e, 3), generator=g, device="cuda", dtype=torch.uint8) for _ in range(8)]
    ref = _env(name, e, mode="A", p=0.01, auto_reset=True, horizon=7)
    _load(ref, case)
    env = VecPBNEnv(product_net(name), e, attractor_set(name), device="cuda:0", perturb_p=0.01, perturb_mode="A",
                    seed=0x5EED, kernel="sliced", resident=True, auto_reset=True, chain=True,
                    **dict(KW, horizon=7))
    _load(env, case)
    env._planes()
    stream = torch.cuda.Stream()
    rewards = []
    with torch.cuda.stream(stream):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for a in acts:
                env.step(a)
            env.advance_counter()
        for rep in range(3):
            graph.replay()
    for rep in range(3):
        for a in acts:
            ref.step(a)
    torch.cuda.synchronize()
    assert torch.equal(env.reward, ref.reward)
    assert torch.equal(env.state, ref.state)
    assert torch.equal(env.t, ref.t)
    assert torch.equal(env.target_id, ref.target_id)
    assert env.stats() == ref.stats()
