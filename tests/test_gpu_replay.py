"""GPU: device replay ring + fused observation unpack (pbn_replay_*, pbn_observe) against the CPU
restatement of the reference's python-list memory and update_policy() tensor building
(oracle/replay_oracle.py)."""
import numpy as np
import pytest

from helpers import attractor_set, product_net
from oracle import replay_oracle as R

pytestmark = pytest.mark.gpu


def _env(name, e, **kw):
    import torch
    from pbn_rl_b200 import VecPBNEnv
    env = VecPBNEnv(product_net(name), e, attractor_set(name), device="cuda:0", horizon=5, bins=3, perturb_p=0.02,
                    perturb_mode="A", seed=3, auto_reset=True, **kw)
    env.reset()
    torch.cuda.synchronize()
    return env


@pytest.mark.parametrize("name,e,cap", [("pbn7", 37, 100), ("pbn28", 1500, 4000), ("pbn70", 300, 700), ("pbn10", 64, 64)])
def test_replay_ring_matches_python_list_memory(name, e, cap):
    import torch
    from pbn_rl_b200.replay import DeviceReplay
    net, attrs = product_net(name), attractor_set(name)
    n = net.n_genes
    env = _env(name, e)
    ring = DeviceReplay(env, cap)
    memory = R.OracleReplay(cap)
    rng = np.random.default_rng(5)
    for step in range(7):   # wraps the ring at least once for every case
        state = env.state.cpu().numpy().astype(np.uint64)
        target = env.target_id.cpu().numpy()
        act = rng.integers(0, n + 1, size=(e, 3), dtype=np.uint8)
        if step == 3:
            ring.step(None)                                        # env.step([]) transitions store zero actions
            act[...] = 0
        else:
            ring.step(torch.from_numpy(act).cuda())
        torch.cuda.synchronize()
        nxt = ring._final.cpu().numpy().astype(np.uint64)          # pre-reset next state of every instance
        rew = env.reward.cpu().numpy()
        done = (env.terminated.cpu().numpy() | env.truncated.cpu().numpy()).astype(bool)
        for k in range(e):                                          # what learn() stores, instance by instance
            memory.store(R.Transition(R.words_to_bits(state[k], n), R.target_state(attrs.attractors, int(target[k]), n),
                                      act[k].astype(np.int64), rew[k], R.words_to_bits(nxt[k], n), done[k]))
        assert len(ring) == len(memory) and ring.head == memory.current_index
        index = rng.integers(0, len(memory), size=257)
        want = memory.batch_tensors(index)
        got = ring.sample(257, index=torch.from_numpy(index))
        for key in ("obs", "next_obs", "actions", "reward", "done"):
            g = got[key].cpu().numpy()
            assert g.dtype == want[key].dtype and g.shape == want[key].shape, key
            assert np.array_equal(g, want[key]), (key, step)
    # the whole ring, in slot order
    want = memory.batch_tensors(range(len(memory)))
    got = ring.sample(0, index=torch.arange(len(memory)))
    assert np.array_equal(got["obs"].cpu().numpy(), want["obs"])
    env.close()


@pytest.mark.parametrize("name,e", [("pbn7", 1), ("pbn28", 4097), ("pbn70", 513)])
def test_observe_is_the_agents_input(name, e):
    import torch
    net, attrs = product_net(name), attractor_set(name)
    env = _env(name, e)
    env.step(None)
    obs = env.observe()
    torch.cuda.synchronize()
    want = R.observation(env.state.cpu().numpy().astype(np.uint64), env.target_id.cpu().numpy(), attrs.attractors,
                         net.n_genes)
    assert obs.dtype == torch.float32 and tuple(obs.shape) == (2, e, net.n_genes)
    assert np.array_equal(obs.cpu().numpy(), want)
    # without a target (-1) the target plane is zero
    env.set_target(-1)
    assert float(env.observe()[1].abs().sum().item()) == 0.0
    env.close()


def test_replay_argument_errors():
    import torch
    from pbn_rl_b200 import _cabi
    from pbn_rl_b200.replay import DeviceReplay
    env = _env("pbn10", 128)
    with pytest.raises(ValueError):
        DeviceReplay(env, 64)
    ring = DeviceReplay(env, 256)
    with pytest.raises(RuntimeError):
        ring.commit(None, env.state)
    with pytest.raises(RuntimeError):
        ring.sample(4)
    ring.head = 999   # out of range: the C entry point refuses
    with pytest.raises(_cabi.PbnError):
        ring.observe()
    env.close()


def test_insitu_bdq_loop_runs():
    """BASELINE config 5 in miniature: GPU env -> device replay ring -> policy update, all on the device."""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("insitu_bdq", Path(__file__).resolve().parent.parent / "scripts" / "insitu_bdq.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.run("pbn28", envs=4096, iters=4, warmup=1, batch=64)
    assert out["value"] > 0 and out["episodes"] > 0 and np.isfinite(out["loss_last"])
    assert out["kernel"] == "sliced" and out["env_launches"] >= 5 * 3


def test_device_ring_reproduces_reference_memory_fixture():
    """The scripted store / overwrite / sample sequence of tests/golden/replay_expected.json (produced by running the
    reference's ExperienceReplay and update_policy tensor code) through pbn_replay_observe / commit / sample."""
    import ctypes as C
    import json
    from pathlib import Path

    import torch
    from pbn_rl_b200 import AttractorSet, PBNNetwork, VecPBNEnv, _cabi
    from pbn_rl_b200.replay import DeviceReplay
    fx = json.loads((Path(__file__).resolve().parent / "golden" / "replay_expected.json").read_text())
    n, cap = fx["n"], fx["capacity"]
    net = PBNNetwork.from_expressions(["g%d" % i for i in range(n)], [["g%d" % i] for i in range(n)])
    targets = sorted({tuple(t["target"]) for t in fx["script"]})
    attrs = AttractorSet([[t] for t in targets], n)
    env = VecPBNEnv(net, 1, attrs, device="cuda:0", horizon=0, bins=fx["bins"])
    ring = DeviceReplay(env, cap)
    bits = lambda v: sum(int(b) << i for i, b in enumerate(v))
    lib, dev = env.lib, env.device
    for t in fx["script"]:      # one transition per store, exactly like memory.store(Transition(...))
        env.set_state(torch.tensor([[bits(t["state"])]], dtype=torch.int64), packed=True)
        env.set_target(targets.index(tuple(t["target"])))
        ring.observe()
        env.reward.fill_(t["reward"])
        env.terminated.fill_(1 if t["done"] else 0)
        env.truncated.fill_(0)
        nxt = torch.tensor([[bits(t["next_state"])]], dtype=torch.int64, device=dev)
        ring.commit(torch.tensor([t["action"]], dtype=torch.uint8, device=dev), nxt)
    assert len(ring) == fx["len"] and ring.head == fx["current_index"]
    for smp in fx["samples"]:
        got = ring.sample(smp["batch"], index=torch.tensor(smp["index"]))
        assert got["obs"][0].cpu().tolist() == smp["states"] and got["obs"][1].cpu().tolist() == smp["targets"]
        assert got["next_obs"][0].cpu().tolist() == smp["next_states"] and got["next_obs"][1].cpu().tolist() == smp["targets"]
        assert got["actions"].cpu().tolist() == smp["actions"] and list(got["actions"].shape) == smp["actions_shape"]
        assert got["reward"].cpu().tolist() == smp["rewards"] and got["done"].cpu().tolist() == smp["masks"]
    env.close()
