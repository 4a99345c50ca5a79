"""TEST INFRASTRUCTURE -- CPU restatement of the reference's all-pairs evaluation loop
(model_tester.py:584-658); only tests/ may import this, the product never does.

Per (run, source attractor, target attractor), in the reference:
    state = source[0] with '*' -> 0; env.graph.setState(state); env.setTarget(target); count = 0   :602-614
    while not env.in_target(state):                                                            :616
        count += 1; action = model.predict(state, target[0]); env.step(action); state = env.render()  :617-625
        if count > 100: result_matrix[s, t] += 101; data[101] += 1; break                       :627-636
    else: result_matrix[s, t] += count; data[count] += 1                                       :638-652
The rollouts are independent, so they are restated step-synchronously over all E = runs*A*A rollouts
with the randomness of rollout e at step k handed in by `draw(k)` (the CPU twins of the kernels'
Philox streams, oracle/pbn_oracle.py) -- which is what makes the comparison with the GPU bit-exact.
"""
from collections import defaultdict

import numpy as np

from . import pbn_oracle as O


def representative_words(attractors, n):
    w = 1 if n <= 64 else 2
    out = np.zeros((len(attractors), w), dtype=np.uint64)
    for a, attr in enumerate(attractors):
        for i, b in enumerate(attr[0]):
            if b != "*" and int(b):
                out[a, i >> 6] |= np.uint64(1) << np.uint64(i & 63)
    return out


def words_to_bit_matrix(words, n):
    e = words.shape[0]
    out = np.zeros((e, n), dtype=np.float32)
    for i in range(n):
        out[:, i] = ((words[:, i >> 6] >> np.uint64(i & 63)) & np.uint64(1)).astype(np.float32)
    return out


def all_pairs(onet, attractors, policy, runs, max_steps, draw, mode=O.PERT_A, n_attractors=None):
    """Returns (matrix[A,A] float64 totals, data defaultdict{steps: rollouts}, count[E])."""
    n = onet.n
    a = len(attractors) if n_attractors is None else n_attractors
    tables = O.attractor_tables(attractors, n)
    offs, care, val = tables
    rep = representative_words(attractors, n)[:a]
    pair = np.tile(np.arange(a * a), runs)
    src, tgt = pair // a, pair % a
    e = len(pair)
    state = rep[src].copy()
    target_bits = words_to_bit_matrix(rep[tgt], n)
    count = np.zeros(e, dtype=np.int64)

    def in_target(words, k):
        return any(((words & care[s]) == val[s]).all() for s in range(offs[tgt[k]], offs[tgt[k] + 1]))

    running = np.array([not in_target(state[k], k) for k in range(e)])
    for step in range(max_steps + 1):
        if not running.any():
            break
        sel, pert = draw(step, e)
        actions = policy(np.stack((words_to_bit_matrix(state, n), target_bits)))
        nxt, _, _, hit, _ = O.batched_step(onet, tables, state, actions, tgt, np.zeros(e, np.uint16), horizon=0, mode=mode,
                                           sel=sel, pert=pert, r_success=5.0, r_step=0.0, r_action=-1.0)
        for k in np.nonzero(running)[0]:
            count[k] += 1
            state[k] = nxt[k]
            if count[k] > max_steps:
                running[k] = False          # failed: booked as max_steps + 1
            elif hit[k]:
                running[k] = False
    matrix = np.zeros((a, a), dtype=np.float64)
    data = defaultdict(int)
    for k in range(e):
        matrix[src[k], tgt[k]] += count[k]
        data[int(count[k])] += 1
    return matrix, data, count
