"""CPU: the evaluator restatement (oracle/evaluate_oracle.py) follows model_tester.py's counting rules."""
import numpy as np

from helpers import attractor_set, oracle_net
from oracle import evaluate_oracle as EO


def test_counting_rules_on_pbn7():
    onet, attrs = oracle_net("pbn7"), attractor_set("pbn7").attractors

    def draw(step, e):  # predictor 0 everywhere, no perturbation
        return np.zeros((e, 7), dtype=np.uint8), np.zeros((e, 1), dtype=np.uint64)

    noop = lambda obs: np.zeros((obs.shape[1], 3), dtype=np.uint8)
    m, data, count = EO.all_pairs(onet, attrs, noop, runs=1, max_steps=5, draw=draw)
    assert np.all(np.diag(m) == 0)
    singles = [k for k, a in enumerate(attrs) if len(a) == 1 and "*" not in a[0]]
    for s in singles:          # fixed points never move: every other target fails with max_steps + 1
        for t in range(len(attrs)):
            if t != s:
                assert m[s, t] == 6
    assert sum(data.values()) == len(attrs) ** 2 and set(data) <= {0, 1, 2, 3, 4, 5, 6}

    def jump(obs):             # flip every differing gene at once (bins = 7): reaches any target in 1 step if it is a fixed point
        diff = obs[0] != obs[1]
        out = np.zeros((diff.shape[0], 7), dtype=np.uint8)
        for k in range(diff.shape[0]):
            idx = np.nonzero(diff[k])[0]
            out[k, : len(idx)] = idx + 1
        return out

    m2, data2, _ = EO.all_pairs(onet, attrs, jump, runs=1, max_steps=5, draw=draw)
    for s in range(len(attrs)):
        for t in singles:
            if s != t:
                assert m2[s, t] == 1
