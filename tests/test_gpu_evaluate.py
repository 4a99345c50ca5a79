"""GPU: batched all-pairs evaluator (pbn_rl_b200/evaluate.py) against the CPU restatement of the
reference's model_tester loop (oracle/evaluate_oracle.py), with identical Philox randomness."""
import pickle

import numpy as np
import pytest

from helpers import attractor_set, oracle_net, product_net
from oracle import evaluate_oracle as EO
from oracle import pbn_oracle as O

pytestmark = pytest.mark.gpu


def hamming_policy_np(bins=3):
    def policy(obs):
        diff = obs[0] != obs[1]
        e, n = diff.shape
        out = np.zeros((e, bins), dtype=np.uint8)
        for k in range(e):
            idx = np.nonzero(diff[k])[0][:bins]
            out[k, : len(idx)] = idx + 1
        return out
    return policy


def noop_policy(bins=3):
    import torch
    return lambda obs: torch.zeros((obs.shape[1], bins), dtype=torch.uint8, device=obs.device)


@pytest.mark.parametrize("name,runs,p,kernel", [("pbn7", 3, 0.0, "auto"), ("pbn10", 2, 0.05, "auto"), ("pbn28", 1, 0.01, "auto"),
                                                ("pbn10", 2, 0.05, "scalar")])
def test_all_pairs_matches_reference_loop(name, runs, p, kernel):
    from pbn_rl_b200.evaluate import evaluate_all_pairs, hamming_policy
    net, onet, attrs = product_net(name), oracle_net(name), attractor_set(name)
    max_steps = 12
    seed = 0xABCDE
    m, data, env, extra = evaluate_all_pairs(net, attrs, hamming_policy(3), runs=runs, max_steps=max_steps, seed=seed,
                                             perturb_p=p, kernel=kernel, poll_every=3, return_env=True)

    def draw(step, e):
        ids = np.arange(e, dtype=np.uint64)
        if env.kernel == "scalar":
            return O.scalar_stream_selection(onet, ids, step, seed), O.scalar_stream_perturbation(onet.n, p, ids, step, seed)
        return O.sliced_stream(onet, p, ids, step, seed)

    want_m, want_data, want_count = EO.all_pairs(onet, attrs.attractors, hamming_policy_np(3), runs, max_steps, draw)
    assert np.array_equal(extra["count"].cpu().numpy(), want_count)
    assert np.array_equal(m, want_m) and dict(data) == dict(want_data)
    a = len(attrs)
    assert np.all(np.diag(m) == 0) and data[0] >= runs * a and sum(data.values()) == runs * a * a
    env.close()


def test_failures_are_booked_as_max_steps_plus_one(tmp_path):
    """A policy that never intervenes cannot leave a fixed-point attractor: every off-diagonal pair of pbn7's
    singleton attractors fails with 101 (model_tester.py:627-636), the diagonal takes 0 steps."""
    from pbn_rl_b200.evaluate import evaluate_all_pairs, load_results, save_results
    net, attrs = product_net("pbn7"), attractor_set("pbn7")
    m, data = evaluate_all_pairs(net, attrs, noop_policy(), runs=2, max_steps=100, n_attractors=4)
    singles = [k for k, at in enumerate(attrs.attractors) if len(at) == 1 and "*" not in at[0]]
    for s in singles:
        for t in range(4):
            assert m[s, t] == (0 if s == t else 2 * 101)
    assert data[101] >= 2 * len(singles) * 3 and data[0] >= 2 * 4
    save_results(tmp_path / "data" / "results" / "pbn_7_4.pkl", m, data)
    with open(tmp_path / "data" / "results" / "pbn_7_4.pkl", "rb") as f:
        raw = pickle.load(f)
    assert isinstance(raw, tuple) and raw[0].dtype == np.float64 and raw[0].shape == (4, 4)
    m2, d2 = load_results(tmp_path / "data" / "results" / "pbn_7_4.pkl")
    assert np.array_equal(m2, m) and dict(d2) == dict(data)
