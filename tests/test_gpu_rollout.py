"""GPU: pbn_rollout (S uncontrolled updates in one launch, states resident on chip as bit-planes) against S
calls of pbn_step with no actions, and against the oracle's deterministic core."""
import numpy as np
import pytest

from helpers import attractor_set, oracle_net, product_net

pytestmark = pytest.mark.gpu


def _pair(name, e, p, mode, **kw):
    import torch
    from pbn_rl_b200 import VecPBNEnv
    net = product_net(name)
    envs = [VecPBNEnv(net, e, attractor_set(name), device="cuda:0", perturb_p=p, perturb_mode=mode, seed=99, horizon=0, **kw)
            for _ in range(2)]
    rng = np.random.default_rng(1)
    masks = np.array(net.state_mask(), dtype=np.uint64)
    st = (rng.integers(0, 2**63, size=(e, net.n_words), dtype=np.int64).astype(np.uint64) * np.uint64(2)
          + rng.integers(0, 2, size=(e, net.n_words)).astype(np.uint64)) & masks
    for env in envs:
        env.set_state(torch.from_numpy(st.astype(np.int64)), packed=True)
        env.set_target(0)
    return envs, st


@pytest.mark.parametrize("name,e", [("pbn28", 4096 + 37), ("pbn70", 2048 + 5), ("pbn7", 33), ("pbn10", 1024)])
@pytest.mark.parametrize("mode,p", [("A", 0.0), ("A", 0.03), ("B", 0.03), ("C", 0.03), ("A", 0.4)])
def test_rollout_equals_repeated_steps(name, e, mode, p):
    import torch
    (ref, dut), _ = _pair(name, e, p, mode)
    t_before = dut.t.clone()
    for chunk in (1, 4, 7):
        for _ in range(chunk):
            ref.step(None)
        dut.rollout(chunk)
        torch.cuda.synchronize()
        assert torch.equal(ref.state, dut.state), (name, mode, p, chunk)
    assert dut.step_ctr == ref.step_ctr == 12
    assert torch.equal(dut.t, t_before)                     # counters are left alone
    if e % 1024 == 0:
        assert ref.stats()["perturbed"] == dut.stats()["perturbed"] and dut.stats()["steps"] == 12 * e
    ref.close()
    dut.close()


def test_rollout_against_the_oracle():
    """Three updates of Bittner-28 with p = 0.02 (model A): the oracle walks the same Philox streams."""
    import torch
    from oracle import pbn_oracle as O
    name, e, p = "pbn28", 3000, 0.02
    (_, dut), st = _pair(name, e, p, "A")
    onet = oracle_net(name)
    ids = np.arange(e, dtype=np.uint64)
    state = st
    zero = np.zeros((e, 1), dtype=np.uint64)
    for step in range(3):
        sel, pert = O.sliced_stream(onet, p, ids, step, 99)
        state = onet.transition_batch(state, zero, sel, pert, O.PERT_A)
    dut.rollout(3)
    torch.cuda.synchronize()
    assert np.array_equal(dut.state.cpu().numpy().astype(np.uint64), state)
    dut.close()


def test_rollout_scalar_fallback_and_counter_modes():
    import torch
    (ref, dut), _ = _pair("pbn10", 500, 0.05, "A", kernel="scalar")
    for _ in range(5):
        ref.step(None)
    dut.rollout(5)
    assert torch.equal(ref.state, dut.state) and int(dut.t.max().item()) == 0
    from pbn_rl_b200 import VecPBNEnv
    env = VecPBNEnv(product_net("pbn10"), 1024, None, device="cuda:0", device_counter=True)
    with pytest.raises(RuntimeError):
        env.rollout(2)
    for x in (ref, dut, env):
        x.close()
