"""Batched all-pairs evaluation of a control policy.

The reference's ``model_tester.py:584-658`` walks ``runs x A x A`` (source attractor, target attractor)
pairs one after the other: start in the source attractor's first state (``'*' -> 0``), call the policy
and ``env.step`` until ``env.in_target(state)`` or more than ``max_steps`` = 100 steps, add the step
count (101 for a failure) to ``result_matrix[source, target]`` and to the histogram ``data``, and
pickle ``(matrix, data)`` to ``data/results/pbn_{n}_{A}.pkl``.  Here every (run, source, target)
triple is one env instance of a :class:`VecPBNEnv`; all rollouts step together and the loop's
bookkeeping runs on the device (``pbn_in_target`` / ``pbn_rollout_track`` / ``pbn_rollout_reduce``).
"""
from __future__ import annotations

import ctypes as C
import pickle
from collections import defaultdict
from pathlib import Path
from typing import Callable, Optional, Tuple, Union

import numpy as np
import torch

from ._cabi import check
from .attractors import AttractorSet
from .network import PBNNetwork
from .vec_env import VecPBNEnv

__all__ = ["evaluate_all_pairs", "save_results", "load_results", "hamming_policy"]

Policy = Callable[[torch.Tensor], torch.Tensor]


def hamming_policy(bins: int = 3) -> Policy:
    """A hand-written baseline policy: flip the first ``bins`` genes in which state and target differ
    (0 = no-op for unused branches).  Deterministic, so evaluator runs are reproducible bit for bit."""

    def policy(obs: torch.Tensor) -> torch.Tensor:
        diff = obs[0] != obs[1]                                            # [E, N]
        rank = torch.cumsum(diff.to(torch.int32), dim=1)                    # 1-based rank of every differing gene
        n = diff.shape[1]
        gene = torch.arange(1, n + 1, device=obs.device, dtype=torch.int32).unsqueeze(0)
        cols = [torch.where(diff & (rank == k + 1), gene, torch.zeros_like(gene)).sum(dim=1) for k in range(bins)]
        return torch.stack(cols, dim=1).to(torch.uint8)

    return policy


def evaluate_all_pairs(network: PBNNetwork, attractors: AttractorSet, policy: Policy, runs: int = 10,
                       max_steps: int = 100, n_attractors: Optional[int] = None,
                       device: Union[str, torch.device] = "cuda:0", seed: int = 0x5EED, bins: int = 3,
                       perturb_p: float = 0.0, perturb_mode: str = "A", kernel: str = "auto",
                       poll_every: int = 8, return_env: bool = False):
    """Run the all-pairs test.  ``policy(obs)`` maps the float32 ``[2, E, N]`` observation
    (:meth:`VecPBNEnv.observe`) to uint8 actions ``[E, bins]`` with values in ``[0, N]`` (duplicates
    allowed, as the raw action tensor of model_tester.py:622-624).  Returns ``(matrix, data)`` exactly
    as the reference pickles them: ``matrix[s, t]`` = total steps over the runs (``max_steps + 1`` per
    failure) as float64 ``[A, A]``, ``data`` = ``defaultdict(int)`` {steps: number of rollouts}."""
    a = len(attractors) if n_attractors is None else int(n_attractors)
    if a < 1 or a > len(attractors):
        raise ValueError("n_attractors=%d outside 1..%d" % (a, len(attractors)))
    e = int(runs) * a * a
    env = VecPBNEnv(network, e, attractors, device=device, seed=seed, horizon=0, bins=bins, perturb_p=perturb_p,
                    perturb_mode=perturb_mode, kernel=kernel)
    dev = env.device
    rep = torch.from_numpy(attractors.representative_words()[:a].astype(np.int64)).to(dev)   # [A, W]
    pair = torch.arange(a * a, device=dev, dtype=torch.int32).repeat(int(runs))               # e -> src * A + tgt
    src, tgt = (pair // a).to(torch.int64), (pair % a).to(torch.int32)
    env.set_state(rep[src], packed=True)
    env.set_target(tgt)
    active = torch.empty((e,), dtype=torch.uint8, device=dev)
    count = torch.zeros((e,), dtype=torch.int32, device=dev)
    n_active = torch.zeros((1,), dtype=torch.int32, device=dev)
    lib, h = env.lib, env._h
    check(lib.pbn_in_target(h, env.state.data_ptr(), env.target_id.data_ptr(), active.data_ptr(), e, env._stream()))
    active.logical_not_()                       # rollouts that start inside their target take 0 steps
    obs = torch.empty((2, e, env.n_genes), dtype=torch.float32, device=dev)
    steps_taken = 0
    for it in range(max_steps + 1):
        env.observe(obs)
        actions = policy(obs)
        env.step(actions, stats=False)
        steps_taken += 1
        poll = (it + 1) % poll_every == 0 or it == max_steps
        if poll:
            n_active.zero_()
        check(lib.pbn_rollout_track(h, env.terminated.data_ptr(), active.data_ptr(), count.data_ptr(), max_steps, e,
                                    n_active.data_ptr() if poll else None, env._stream()))
        if poll and int(n_active.item()) == 0:
            break
    matrix = torch.zeros((a * a,), dtype=torch.int64, device=dev)
    hist = torch.zeros((max_steps + 2,), dtype=torch.int64, device=dev)
    check(lib.pbn_rollout_reduce(h, count.data_ptr(), pair.data_ptr(), e, a * a, max_steps, matrix.data_ptr(),
                                 hist.data_ptr(), env._stream()))
    m = matrix.cpu().numpy().astype(np.float64).reshape(a, a)
    data = defaultdict(int)
    for steps, n in enumerate(hist.cpu().tolist()):
        if n:
            data[int(steps)] = int(n)
    if return_env:
        return m, data, env, {"count": count, "steps_taken": steps_taken}
    env.close()
    return m, data


def save_results(path: Union[str, Path], matrix: np.ndarray, data) -> None:
    """``pkl.dump((save_matrix, data), f)`` of model_tester.py:656-658 (totals, not per-run means)."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump((np.asarray(matrix, dtype=np.float64), defaultdict(int, data)), f)


def load_results(path: Union[str, Path]) -> Tuple[np.ndarray, defaultdict]:
    with open(path, "rb") as f:
        matrix, data = pickle.load(f)
    return np.asarray(matrix, dtype=np.float64), data
