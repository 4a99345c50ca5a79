"""Device-resident replay ring for the batched env.

The reference stores one ``Transition(state, target, action, reward, next_state, done)`` per env
step in a python list and samples with ``random.sample`` (bdq_model/memory.py:22-70); every policy
update then rebuilds float tensors from the tuples (bdq_model/__init__.py:100-109).  With 2^20 env
instances per step that path is the bottleneck, so here the transitions stay *packed* in HBM
(28 B per transition at N <= 64, bins = 3) and one gather+unpack kernel produces exactly the tensors
``update_policy`` consumes.  All device work goes through the C-ABI (``pbn_replay_*``); torch only
lends memory and draws the sample indices.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _cabi
from ._cabi import check

__all__ = ["DeviceReplay"]


class DeviceReplay:
    """Ring of ``capacity`` packed transitions next to a :class:`VecPBNEnv`.

    Protocol per env step (the two halves of ``memory.store``)::

        replay.observe()                 # before env.step: (state, target) of every instance
        env.step(actions, final_state=nxt)
        replay.commit(actions, nxt)      # after: (action, reward, done, next_state)

    or simply ``replay.step(actions)`` which does the three calls.
    """

    def __init__(self, env, capacity: int):
        if capacity < env.num_envs:
            raise ValueError("capacity %d < num_envs %d: one step would overwrite itself" % (capacity, env.num_envs))
        self.env = env
        self.capacity = int(capacity)
        dev, w, b = env.device, env.n_words, env.bins
        c = self.capacity
        self.state = torch.zeros((c, w), dtype=torch.int64, device=dev)
        self.next_state = torch.zeros((c, w), dtype=torch.int64, device=dev)
        self.target_id = torch.full((c,), -1, dtype=torch.int32, device=dev)
        self.actions = torch.zeros((c, b), dtype=torch.uint8, device=dev)
        self.reward = torch.zeros((c,), dtype=torch.float32, device=dev)
        self.done = torch.zeros((c,), dtype=torch.uint8, device=dev)
        self._r = _cabi.Replay()
        self._r.state, self._r.next_state = self.state.data_ptr(), self.next_state.data_ptr()
        self._r.target_id, self._r.actions = self.target_id.data_ptr(), self.actions.data_ptr()
        self._r.reward, self._r.done = self.reward.data_ptr(), self.done.data_ptr()
        self._r.capacity = c
        self.head = 0          # next slot to write
        self.size = 0          # transitions stored (<= capacity)
        self._pending = False
        self._final = torch.zeros((env.num_envs, w), dtype=torch.int64, device=dev)

    def __len__(self) -> int:
        return self.size

    @property
    def bytes_per_transition(self) -> int:
        return 16 * self.env.n_words + 4 + self.env.bins + 4 + 1

    def observe(self) -> None:
        env = self.env
        check(env.lib.pbn_replay_observe(env._h, C.byref(self._r), self.head, env.state.data_ptr(),
                                         env.target_id.data_ptr(), env.num_envs, env._stream()))
        self._pending = True

    def commit(self, actions: Optional[torch.Tensor], next_state: torch.Tensor) -> None:
        if not self._pending:
            raise RuntimeError("commit() without observe()")
        env = self.env
        actions = env._check_actions(actions)
        check(env.lib.pbn_replay_commit(env._h, C.byref(self._r), self.head,
                                        None if actions is None else actions.data_ptr(), env.reward.data_ptr(),
                                        env.terminated.data_ptr(), env.truncated.data_ptr(), next_state.data_ptr(),
                                        env.num_envs, env._stream()))
        self.head = (self.head + env.num_envs) % self.capacity
        self.size = min(self.capacity, self.size + env.num_envs)
        self._pending = False

    def step(self, actions: Optional[torch.Tensor]):
        """observe -> env.step -> commit.  Returns what ``env.step`` returns."""
        self.observe()
        actions = self.env._check_actions(actions)
        out = self.env.step(actions, final_state=self._final)
        self.commit(actions, self._final)
        return out

    def sample(self, batch_size: int, generator: Optional[torch.Generator] = None,
               index: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """``memory.sample(batch_size)`` + the tensor building of ``update_policy``: returns
        ``obs`` [2,B,N] (states, targets), ``next_obs`` [2,B,N] (next states, targets) float32,
        ``actions`` [B,bins,1] int64, ``reward`` [B,1], ``done`` [B,1] float32 and the ``index`` used."""
        env = self.env
        if index is None:
            if self.size == 0:
                raise RuntimeError("the replay ring is empty")
            index = torch.randint(0, self.size, (batch_size,), device=env.device, generator=generator)
        index = index.to(device=env.device, dtype=torch.int64).contiguous()
        b, n = int(index.numel()), env.n_genes
        obs = torch.empty((2, b, n), dtype=torch.float32, device=env.device)
        nxt = torch.empty((2, b, n), dtype=torch.float32, device=env.device)
        act = torch.empty((b, env.bins, 1), dtype=torch.int64, device=env.device)
        rew = torch.empty((b, 1), dtype=torch.float32, device=env.device)
        done = torch.empty((b, 1), dtype=torch.float32, device=env.device)
        check(env.lib.pbn_replay_sample(env._h, C.byref(self._r), index.data_ptr(), b, obs.data_ptr(), nxt.data_ptr(),
                                        act.data_ptr(), rew.data_ptr(), done.data_ptr(), env._stream()))
        return {"obs": obs, "next_obs": nxt, "actions": act, "reward": rew, "done": done, "index": index}
