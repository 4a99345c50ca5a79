"""TEST INFRASTRUCTURE -- CPU restatement of the reference's replay memory and of the tensors its
agent builds from it (only tests/ may import this; the product never does).

Follows, in the reference tree:
  ExperienceReplay.store / __len__ (python list, overwrite at current_index)   bdq_model/memory.py:40-51,64-70
  Transition(state, target, action, reward, next_state, done)                  bdq_model/memory.py:17-19
  what learn() stores per env step                                             bdq_model/__init__.py:186-195
  tensors of update_policy(): states/targets/next_states float, actions long
  [B,bins,1], rewards [B,1], masks (= the stored `done`) [B,1], stacked as
  (states, targets) and (next_states, targets)                                 bdq_model/__init__.py:100-118
  network input of predict(): np.stack((state, target)) as float               bdq_model/__init__.py:92-93
The reference's tests (ddqn_per/test_memory.py) are stale and pin nothing here.
"""
from collections import namedtuple
from typing import List, Sequence

import numpy as np

Transition = namedtuple("Transition", ("state", "target", "action", "reward", "next_state", "done"))


def words_to_bits(words: Sequence[int], n: int) -> tuple:
    """Packed state words (bit i of the state = gene i, 64 genes per word) -> tuple of N ints."""
    return tuple((int(words[i >> 6]) >> (i & 63)) & 1 for i in range(n))


def target_state(attractors, target_id: int, n: int) -> tuple:
    """`target` as env.reset() hands it to the agent: first state of the attractor, '*' -> 0
    (model_tester.py:609 convention); all zeros when there is no target."""
    if target_id < 0 or target_id >= len(attractors):
        return tuple([0] * n)
    return tuple(0 if b == "*" else int(b) for b in attractors[target_id][0])


class OracleReplay:
    """bdq_model/memory.py:22-51 restated: a python list that grows to `capacity`, then overwrites."""

    def __init__(self, capacity: int):
        self.capacity = capacity
        self.buffer: List[Transition] = []
        self.current_index = 0

    def store(self, transition: Transition) -> None:
        if len(self.buffer) < self.capacity:
            self.buffer.append(transition)
        else:
            self.buffer[self.current_index] = transition
        self.current_index = (self.current_index + 1) % self.capacity

    def __len__(self) -> int:
        return len(self.buffer)

    def batch_tensors(self, index: Sequence[int]):
        """The arrays update_policy() builds from memory.sample() when the sample is buffer[index]."""
        x = [self.buffer[i] for i in index]
        b_states, b_targets, b_actions, b_rewards, b_next_states, b_masks = zip(*x)
        states = np.stack(b_states).astype(np.float32)
        targets = np.stack(b_targets).astype(np.float32)
        actions = np.stack(b_actions).astype(np.int64).reshape(states.shape[0], -1, 1)
        rewards = np.stack(b_rewards).astype(np.float32).reshape(-1, 1)
        next_states = np.stack(b_next_states).astype(np.float32)
        masks = np.stack(b_masks).astype(np.float32).reshape(-1, 1)
        return {"obs": np.stack((states, targets)), "next_obs": np.stack((next_states, targets)),
                "actions": actions, "reward": rewards, "done": masks}


def observation(words: np.ndarray, target_ids: np.ndarray, attractors, n: int) -> np.ndarray:
    """predict()'s input for E instances: float32 [2, E, N] = (state bits, target-state bits)."""
    e = words.shape[0]
    out = np.zeros((2, e, n), dtype=np.float32)
    for k in range(e):
        out[0, k] = words_to_bits(words[k], n)
        out[1, k] = target_state(attractors, int(target_ids[k]), n)
    return out
