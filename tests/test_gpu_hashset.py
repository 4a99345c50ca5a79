"""GPU: attractor membership through the hash set (large attractors) against the oracle's attractor_contains --
env.in_target / is_attracting_state (pbn_in_target, pbn_attractor_id) and the `terminated` flag of both step kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _identity_net(n):
    from pbn_rl_b200 import PBNNetwork
    genes = ["g%d" % i for i in range(n)]
    return PBNNetwork.from_expressions(genes, [[g] for g in genes])   # every gene keeps its value: next state = state


def _big_table(n, rng, sizes=(5000, 300, 1, 17), wildcards=True):
    """Attractors of `sizes` distinct random states + (wildcards) one attractor made of two wildcard patterns."""
    from pbn_rl_b200 import AttractorSet
    seen, attractors = set(), []
    for size in sizes:
        states = []
        while len(states) < size:
            s = tuple(int(v) for v in rng.integers(0, 2, size=n))
            if s not in seen:
                seen.add(s)
                states.append(s)
        attractors.append(states)
    wild = [tuple([1, 0, "*", 1] + ["*"] * 2 + [0] * (n - 6)), tuple([0, 1, 1, "*"] + [1] * (n - 4))]
    if wildcards:
        attractors.insert(2, wild)
    return AttractorSet(attractors, n)


@pytest.mark.parametrize("n", [28, 40, 70])
@pytest.mark.parametrize("kernel", ["sliced", "scalar"])
def test_membership_with_a_5000_state_attractor(n, kernel):
    import torch
    from oracle import pbn_oracle as O
    from pbn_rl_b200 import VecPBNEnv
    rng = np.random.default_rng(n)
    net, attrs = _identity_net(n), _big_table(n, rng)
    e = 6000
    env = VecPBNEnv(net, e, attrs, device="cuda:0", horizon=0, kernel=kernel)
    assert env.kernel == kernel
    assert env.lib.pbn_attractor_hash_slots(env._h) >= 2 * (5000 + 300 + 1 + 17)
    # states: a third from the big attractor, some from the others / the wildcard patterns, the rest random
    rows = []
    for k in range(e):
        r = k % 6
        if r in (0, 1):
            rows.append(attrs.attractors[0][int(rng.integers(0, 5000))])
        elif r == 2:
            a = int(rng.integers(1, len(attrs.attractors)))
            st = attrs.attractors[a][int(rng.integers(0, len(attrs.attractors[a])))]
            rows.append(tuple(int(rng.integers(0, 2)) if b == "*" else b for b in st))
        else:
            rows.append(tuple(int(v) for v in rng.integers(0, 2, size=n)))
    target = rng.integers(0, len(attrs.attractors), size=e).astype(np.int32)
    target[::3] = 0
    env.set_state(np.array(rows, dtype=np.uint8))
    env.set_target(torch.from_numpy(target))
    want_in = np.array([O.attractor_contains(attrs.attractors[int(t)], r) for r, t in zip(rows, target)])
    want_id = np.array([next((a for a, at in enumerate(attrs.attractors) if O.attractor_contains(at, r)), -1) for r in rows])
    assert want_in.sum() > 1000 and (want_id >= 0).sum() > 2500
    out = torch.empty((e,), dtype=torch.uint8, device="cuda")
    from pbn_rl_b200._cabi import check
    check(env.lib.pbn_in_target(env._h, env.state.data_ptr(), env.target_id.data_ptr(), out.data_ptr(), e, env._stream()))
    assert np.array_equal(out.cpu().numpy().astype(bool), want_in)
    assert np.array_equal(env.attractor_ids().cpu().numpy(), want_id)
    # the step kernel: the network keeps every state, so `terminated` is the membership of the (unchanged) state
    env.step(None)
    torch.cuda.synchronize()
    assert np.array_equal(env.terminated.cpu().numpy().astype(bool), want_in)
    env.close()


@pytest.mark.parametrize("n", [28, 70])
@pytest.mark.parametrize("e", [6000, 3 * 1024])
def test_step_membership_without_wildcards_probes_through_the_warp_queue(n, e):
    """Fully specified tables (no wildcard entry): the row kernel tests single-state targets in shared memory and
    compacts the envs of the larger targets into a per-warp queue for the hash-set probes -- ragged and full tiles,
    one- and two-word states, targets of every size, envs without a target."""
    import torch
    from oracle import pbn_oracle as O
    from pbn_rl_b200 import VecPBNEnv
    rng = np.random.default_rng(100 + n)
    net, attrs = _identity_net(n), _big_table(n, rng, sizes=(5000, 1, 300, 1, 17, 1), wildcards=False)
    env = VecPBNEnv(net, e, attrs, device="cuda:0", horizon=0, kernel="sliced")
    assert env.lib.pbn_attractor_hash_slots(env._h) > 0
    A = len(attrs.attractors)
    rows, target = [], rng.integers(0, A, size=e).astype(np.int32)
    for k in range(e):
        r = k % 4
        if r == 0:      # a state of the env's own target
            at = attrs.attractors[int(target[k])]
            rows.append(at[int(rng.integers(0, len(at)))])
        elif r == 1:    # a state of some attractor, usually not the target
            at = attrs.attractors[int(rng.integers(0, A))]
            rows.append(at[int(rng.integers(0, len(at)))])
        else:
            rows.append(tuple(int(v) for v in rng.integers(0, 2, size=n)))
    target[5::97] = -1   # no target: never terminated
    env.set_state(np.array(rows, dtype=np.uint8))
    env.set_target(torch.from_numpy(target))
    want = np.array([t >= 0 and O.attractor_contains(attrs.attractors[int(t)], r) for r, t in zip(rows, target)])
    assert want.sum() > e // 5 and (~want).sum() > e // 3
    for _ in range(2):   # the network keeps every state: the second step sees the same states
        env.step(None)
        torch.cuda.synchronize()
        assert np.array_equal(env.terminated.cpu().numpy().astype(bool), want)
    env.close()


def test_growing_the_table_switches_paths():
    """Small table (scan) -> large table (hash set) -> small again on the same handle (env.all_attractors may grow)."""
    import torch
    from pbn_rl_b200 import AttractorSet, VecPBNEnv
    n = 12
    net = _identity_net(n)
    small = AttractorSet([[tuple([1] * n)], [tuple([0] * n), tuple([1] + [0] * (n - 1))]], n)
    env = VecPBNEnv(net, 2048, small, device="cuda:0", horizon=0, kernel="sliced")
    assert env.lib.pbn_attractor_hash_slots(env._h) == 0
    rng = np.random.default_rng(0)
    big = _big_table(n, rng, sizes=(600, 3))
    env.set_attractors(big)
    assert env.lib.pbn_attractor_hash_slots(env._h) > 0
    st = np.array([big.attractors[0][k % 600] for k in range(2048)], dtype=np.uint8)
    env.set_state(st)
    env.set_target(0)
    env.step(None)
    assert bool(env.terminated.all().item())
    env.set_attractors(small)
    assert env.lib.pbn_attractor_hash_slots(env._h) == 0
    env.set_state(np.ones((2048, n), dtype=np.uint8))
    env.set_target(0)
    env.step(None)
    assert bool(env.terminated.all().item())
    env.close()
