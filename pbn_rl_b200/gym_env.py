"""Per-instance gym-style view of the GPU env: the drop-in for ``gym.make("gym-PBN/...")``.

The reference's agents and scripts drive ONE env instance through gymnasium's API plus ~25
extra attributes of the gym-PBN fork.  :class:`PBNEnv` serves that
surface on top of :class:`pbn_rl_b200.vec_env.VecPBNEnv` (``num_envs`` instances on the GPU;
the per-instance methods act on instance 0, the ``vec`` attribute exposes the batch).

Call sites this mirrors (reference file:line):
  reset()/step()/done protocol          bdq_model/__init__.py:161-213
  action forms                          bdq_model/__init__.py:176-177 (list of 0-d tensors),
                                        model_tester.py:559-561,622-624 (raw 1-D tensor, duplicates),
                                        graph_classifier/__init__.py:148 ([]), ddqn_per/__init__.py:354 (int)
  graph.setState / setTarget / in_target / render        model_tester.py:602-625
  all_attractors / attracting_states / real_attractors   bdq_model/__init__.py:60,182; train_ddqn.py:164-174
  state_attractor_id / target_attractor_id / rework_probas   bdq_model/__init__.py:180,203
  observation_space.shape / action_space / env.env.env   train_BDQ.py:82,116; model_tester.py:540
  graph.nodes / getNodeByID / get_adj_list / genSTG      gbdq_model/__init__.py:259-277; print_graph.py:15-21

Everything the reference tree does not pin (perturbation model and rate, rewards, the
curriculum formula of ``rework_probas``) is a constructor parameter with a documented default.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .attractors import AttractorSet, find_attractors_stg, load_attractor_pickle, sorted_id_permutation
from .network import PBNNetwork

__all__ = ["PBNEnv", "ControlPBNEnv", "make", "register_gym_ids", "Graph", "Node", "Discrete", "Box", "ENV_IDS"]

MAX_ACTIONS = 8  # PBN_MAX_BINS


# --------------------------------------------------------------------------------------
# minimal spaces (gymnasium is not a dependency; shapes/sample() are what the agents use)
# --------------------------------------------------------------------------------------

class Discrete:
    def __init__(self, n: int, seed: Optional[int] = None):
        self.n = int(n)
        self.shape = ()
        self._rng = np.random.default_rng(seed)

    def sample(self) -> int:
        return int(self._rng.integers(0, self.n))

    def contains(self, x) -> bool:
        return 0 <= int(x) < self.n


class Box:
    def __init__(self, low, high, shape, dtype=np.int8):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def sample(self):
        return np.random.randint(self.low, self.high + 1, size=self.shape).astype(self.dtype)


# --------------------------------------------------------------------------------------
# graph introspection (host only)
# --------------------------------------------------------------------------------------

class Node:
    """One gene: ``index``, ``name``/``ID`` and ``predictors = [(input_ids, truth_table, prob), ...]``
    (the GNN agents read ``predictors[k][0]``, gbdq_model/__init__.py:259-277)."""

    def __init__(self, index: int, name: str, predictors):
        self.index = index
        self.name = name
        self.ID = name
        self.predictors = predictors
        self.value = 0


class Graph:
    def __init__(self, env: "PBNEnv"):
        self._env = env
        net = env.network
        self.nodes = [
            Node(i, g, [(list(f.inputs), f.lut, p) for f, p in zip(net.functions[i], net.probabilities[i])])
            for i, g in enumerate(net.genes)
        ]
        self._by_id = {n.name: n for n in self.nodes}

    def setState(self, state) -> None:
        """Overwrite the current state (model_tester.py:611, graph_classifier/__init__.py:122)."""
        self._env._set_state(state)

    def getState(self):
        return self._env.render()

    def getNodeByID(self, node_id) -> Node:
        key = str(node_id)
        if key in self._by_id:
            return self._by_id[key]
        if "x" + key in self._by_id:
            return self._by_id["x" + key]
        raise KeyError(node_id)

    def get_adj_list(self) -> List[List[int]]:
        return self._env.network.adjacency()

    def genSTG(self) -> Dict[Tuple[int, ...], Tuple[None, Dict[Tuple[int, ...], float]]]:
        """Perturbation-free state-transition graph ``state -> (None, {successor: probability})``
        for small networks (print_graph.py:15-21).  Host enumeration: N <= 16."""
        net = self._env.network
        n = net.n_genes
        if n > 16:
            raise ValueError("genSTG enumerates 2^N states on the host; N=%d is too large" % n)
        out = {}
        for s in range(1 << n):
            succ: Dict[int, float] = {0: 1.0}
            for i, (fs, ps) in enumerate(zip(net.functions, net.probabilities)):
                p1 = 0.0
                for f, p in zip(fs, ps):
                    a = 0
                    for j, g in enumerate(f.inputs):
                        a |= ((s >> g) & 1) << j
                    if (f.lut >> a) & 1:
                        p1 += p
                nxt: Dict[int, float] = {}
                for t, pr in succ.items():
                    if p1 > 0:
                        nxt[t | (1 << i)] = nxt.get(t | (1 << i), 0.0) + pr * p1
                    if p1 < 1:
                        nxt[t] = nxt.get(t, 0.0) + pr * (1 - p1)
                succ = nxt
            key = tuple((s >> i) & 1 for i in range(n))
            out[key] = (None, {tuple((t >> i) & 1 for i in range(n)): pr for t, pr in succ.items()})
        return out


# --------------------------------------------------------------------------------------
# the env
# --------------------------------------------------------------------------------------

def _as_int_list(action) -> List[int]:
    """Every action form the reference passes to ``env.step`` -> list of ints."""
    if action is None:
        return []
    if isinstance(action, (int, np.integer)):
        return [int(action)]
    if hasattr(action, "detach"):  # torch tensor (0-d or 1-d)
        action = action.detach().cpu().reshape(-1).tolist()
        return [int(a) for a in action]
    if isinstance(action, np.ndarray):
        return [int(a) for a in action.reshape(-1)]
    out = []
    for a in action:
        out.extend(_as_int_list(a))
    return out


class PBNEnv:
    """Drop-in PBN control env.  See the module docstring for the mirrored call sites."""

    metadata = {"render_modes": ["human"]}

    def __init__(self, network: Optional[PBNNetwork] = None, attractors: Optional[AttractorSet] = None, *,
                 N: Optional[int] = None, genes: Optional[Sequence[str]] = None, logic_functions=None,
                 horizon: int = 20, min_attractors: Optional[int] = None, device="cuda:0",
                 seed: Optional[int] = None, perturb_p: float = 0.0, perturb_mode: str = "A",
                 r_success: float = 5.0, r_step: float = 0.0, r_action: float = -1.0, r_wrong: float = 0.0, num_envs: int = 1,
                 kernel: str = "auto", name: str = ""):
        from .vec_env import VecPBNEnv  # needs torch + the CUDA extension; no CPU fallback

        if network is None:
            if genes is None or logic_functions is None:
                raise ValueError("pass a PBNNetwork, or genes= and logic_functions= as the reference does")
            network = PBNNetwork.from_logic_functions(genes, logic_functions, name)
        if N is not None and int(N) != network.n_genes:
            raise ValueError("N=%s but the network has %d genes" % (N, network.n_genes))
        self.network = network
        self.N = network.n_genes
        if attractors is None:
            if network.n_genes <= 16:
                attractors, _ = find_attractors_stg(network)  # exhaustive sink-SCC search on the host
            else:
                # large networks: massive GPU rollouts + device closure search (exact attractors with a
                # sampled basin; pbn_rl_b200/discover.py), most visited first, capped like the fork's tables
                from .discover import find_attractors_rollout
                attractors, info = find_attractors_rollout(network, n_rollouts=1 << 16, burn_in=256, device=device,
                                                           seed=0x5EED if seed is None else int(seed))
                self.attractor_search = info
        if min_attractors is not None and len(attractors) < int(min_attractors):
            raise ValueError("found %d attractors, min_attractors=%s" % (len(attractors), min_attractors))
        self.horizon = int(horizon)
        self._seed = 0x5EED if seed is None else int(seed)
        self.vec = VecPBNEnv(network, num_envs, attractors, device=device, seed=self._seed, horizon=horizon,
                             bins=MAX_ACTIONS, perturb_p=perturb_p, perturb_mode=perturb_mode,
                             r_success=r_success, r_step=r_step, r_action=r_action, r_wrong=r_wrong, kernel=kernel)
        self._attractors = attractors
        self.observation_space = Box(0, 1, (self.N,), np.int8)
        self.action_space = Discrete(self.N + 1, seed)
        self.discrete_action_space = self.action_space
        self.graph = Graph(self)
        self.n_steps = 0
        self.state_attractor_id = -1
        self.target_attractor_id = -1
        self._pair_weights = self._default_pair_weights(len(attractors))
        self._pair_stats = np.zeros((len(attractors), len(attractors), 2), dtype=np.float64)  # episodes, len sum

    # ---- wrapper chains of the reference: env.env.env[.env] and env.unwrapped all resolve to the env
    @property
    def env(self) -> "PBNEnv":
        return self

    @property
    def unwrapped(self) -> "PBNEnv":
        return self

    # ---- attractor bookkeeping ------------------------------------------------------------------
    @staticmethod
    def _default_pair_weights(a: int) -> np.ndarray:
        w = np.ones((a, a), dtype=np.float64)
        if a > 1:
            np.fill_diagonal(w, 0.0)
        return w

    @property
    def all_attractors(self) -> List[List[tuple]]:
        return self._attractors.attractors

    @property
    def real_attractors(self) -> List[List[tuple]]:
        return self._attractors.attractors

    @property
    def attracting_states(self) -> set:
        """All states of all attractors, wildcards expanded to 0 (model_tester.py:609 convention)."""
        return {tuple(0 if b == "*" else int(b) for b in s) for attr in self._attractors.attractors for s in attr}

    @property
    def target_nodes(self) -> set:
        """States of the current target attractor."""
        if self.target_attractor_id < 0:
            return set()
        return {tuple(0 if b == "*" else int(b) for b in s) for s in self._attractors.attractors[self.target_attractor_id]}

    def add_attractor(self, states: Sequence[Sequence]) -> int:
        """Register a newly discovered attractor (the table is re-uploaded; agents watch
        ``len(env.all_attractors)``, bdq_model/__init__.py:182-184)."""
        new = AttractorSet(self._attractors.attractors + [list(map(tuple, states))], self.N)
        a = len(new)
        w = self._default_pair_weights(a)
        w[: a - 1, : a - 1] = self._pair_weights
        stats = np.zeros((a, a, 2))
        stats[: a - 1, : a - 1] = self._pair_stats
        self._attractors, self._pair_weights, self._pair_stats = new, w, stats
        self.vec.set_attractors(new, w)
        return a - 1

    def setTarget(self, attractor) -> None:
        """``env.setTarget(all_attractors[k])`` (model_tester.py:614) or ``setTarget(k)``."""
        if isinstance(attractor, (int, np.integer)):
            idx = int(attractor)
        else:
            want = [tuple(s) for s in attractor]
            idx = next((k for k, attr in enumerate(self._attractors.attractors) if [tuple(s) for s in attr] == want), -1)
            if idx < 0:
                raise ValueError("unknown attractor")
        self.target_attractor_id = idx
        self.vec.target_id[0] = idx

    def in_target(self, state) -> bool:
        return self.target_attractor_id >= 0 and self._attractors.contains(self.target_attractor_id, list(state))

    def is_attracting_state(self, state) -> bool:
        return self._attractors.attractor_of(list(state)) >= 0

    def rework_probas(self, ep_len: Optional[int] = None) -> None:
        """Curriculum over (source, target) pairs: after an episode of length ``ep_len`` the pair it
        was played on is sampled in proportion to ``0.1 + mean_len / horizon`` (harder pairs more
        often).  The fork's formula is not in the reference tree; this one is ours."""
        s, t = self.state_attractor_id, self.target_attractor_id
        if ep_len is None or s < 0 or t < 0:
            return
        self._pair_stats[s, t, 0] += 1
        self._pair_stats[s, t, 1] += float(ep_len)
        mean_len = self._pair_stats[s, t, 1] / self._pair_stats[s, t, 0]
        self._pair_weights[s, t] = 0.1 + mean_len / max(self.horizon, 1)
        self.vec.set_attractors(self._attractors, self._pair_weights)

    # ---- gym API ----------------------------------------------------------------------------------
    def _state0(self) -> Tuple[int, ...]:
        bits = self.vec.unpack(self.vec.state[:1])[0].cpu().numpy()
        return tuple(int(b) for b in bits)

    def _set_state(self, state) -> None:
        import torch

        bits = torch.tensor([[0 if b == "*" else int(b) for b in state]], dtype=torch.uint8)
        if self.vec.num_envs == 1:
            self.vec.set_state(bits, packed=False)
        else:
            cur = self.vec.unpack().cpu()
            cur[0] = bits[0]
            self.vec.set_state(cur, packed=False)

    def target_state(self) -> Tuple[int, ...]:
        s = self._attractors.attractors[self.target_attractor_id][0]
        return tuple(0 if b == "*" else int(b) for b in s)

    def reset(self, seed: Optional[int] = None, options=None):
        """``(state, target), info = env.reset()`` (bdq_model/__init__.py:161)."""
        if seed is not None:
            self.vec.step_ctr = int(seed) << 20  # a fresh region of the counter space
        self.vec.reset()
        self.n_steps = 0
        self.state_attractor_id = int(self.vec.source_id[0].item())
        self.target_attractor_id = int(self.vec.target_id[0].item())
        return (self._state0(), self.target_state()), {}

    def step(self, action):
        """``state, reward, terminated, truncated, info = env.step(action)``."""
        import torch

        acts = sorted(set(a for a in _as_int_list(action) if a != 0))
        for a in acts:
            if not 0 <= a <= self.N:
                raise ValueError("action %d outside [0, %d]" % (a, self.N))
        if len(acts) > MAX_ACTIONS:
            raise ValueError("at most %d simultaneous interventions are supported" % MAX_ACTIONS)
        buf = torch.zeros((self.vec.num_envs, MAX_ACTIONS), dtype=torch.uint8)
        buf[0, : len(acts)] = torch.tensor(acts, dtype=torch.uint8)
        self.vec.step(buf.to(self.vec.device))
        self.n_steps += 1
        state = self._state0()
        out = (state, float(self.vec.reward[0].item()), bool(self.vec.terminated[0].item()),
               bool(self.vec.truncated[0].item()), {})
        return out

    def render(self, mode: str = "human"):
        return self._state0()

    def close(self) -> None:
        self.vec.close()


class ControlPBNEnv(PBNEnv):
    """``gym.make("gym-PBN/ControlPBNEnv", ..., control_nodes=[...])`` (train_control_gbdq.py:45-72): only
    the listed genes can be intervened on.  The control agent has one *binary* branch per control node
    (``GraphBranchingQNetwork(observation, 2, len(env.control_nodes))``, control_gbdq_model/__init__.py:35-38)
    and passes the length-``len(control_nodes)`` 0/1 vector to ``env.step`` (:66-86,169): entry ``i`` = 1 flips
    gene ``control_nodes[i]``.  Indices are 0-based positions in ``genes``; entries outside ``[0, N)`` (the
    reference's own example lists node 14 for a 14-gene network) are kept in ``control_nodes`` so the
    agent's branch count matches, but are inert."""

    def __init__(self, *args, control_nodes: Sequence[int] = (), **kwargs):
        super().__init__(*args, **kwargs)
        self.control_nodes = [int(c) for c in control_nodes]
        if sum(1 for c in self.control_nodes if 0 <= c < self.N) > MAX_ACTIONS:
            raise ValueError("at most %d controllable genes are supported" % MAX_ACTIONS)
        self.action_space = Box(0, 1, (len(self.control_nodes),), np.int8)
        self.discrete_action_space = Discrete(2)

    def step(self, action):
        bits = _as_int_list(action)
        if len(bits) != len(self.control_nodes):
            raise ValueError("control action has %d entries, env has %d control nodes" % (len(bits), len(self.control_nodes)))
        if any(b not in (0, 1) for b in bits):
            raise ValueError("control actions are 0/1 per control node")
        flips = [c + 1 for b, c in zip(bits, self.control_nodes) if b and 0 <= c < self.N]
        return super().step(flips)


# --------------------------------------------------------------------------------------
# gym.make-style construction
# --------------------------------------------------------------------------------------

ENV_IDS = ("gym-PBN/BittnerMultiGeneral", "gym-PBN/PBNEnv", "gym-PBN/ControlPBNEnv",
           "gym-PBN/Bittner-7", "gym-PBN/Bittner-10", "gym-PBN/Bittner-28", "gym-PBN/Bittner-70")


def _find_file(candidates: Sequence[Union[str, Path]]) -> Optional[Path]:
    for c in candidates:
        p = Path(c)
        if p.exists():
            return p
    return None


def _bittner(n: int, root: Union[str, Path, None], **kw) -> PBNEnv:
    """The fork keeps its working files relative to the CWD: the network as
    ``kaban/pbn{N}.ispl``, attractors as ``data/attractors_Bittner-{N}.pkl``.  The pickle is in
    ascending gene-ID order (fixture K3), so it is permuted onto the ISPL's ``Vars:`` order."""
    root = Path(root) if root is not None else Path(os.environ.get("PBN_RL_ROOT", "."))
    ispl = kw.pop("ispl_path", None) or _find_file([root / "kaban" / f"pbn{n}.ispl"])
    if ispl is None:
        raise FileNotFoundError("kaban/pbn%d.ispl not found under %s (pass ispl_path= or root=)" % (n, root))
    net = PBNNetwork.from_ispl_file(ispl)
    attractors = kw.pop("attractors", None)
    pkl = kw.pop("attractor_path", None) or _find_file([root / "data" / f"attractors_Bittner-{n}.pkl"])
    if attractors is None and pkl is not None:
        raw = load_attractor_pickle(pkl, net.n_genes)
        try:
            attractors = raw.permuted(sorted_id_permutation(net.genes))
        except ValueError:
            attractors = raw
    return PBNEnv(net, attractors, **kw)


def make(env_id: str, root: Union[str, Path, None] = None, **kwargs) -> PBNEnv:
    """``gym.make("gym-PBN/<id>", **kwargs)`` without gym: same ids, same kwargs
    (train_BDQ.py:50, train_assa_BQN.py:121-124, train_ddqn.py:61, print_graph.py:12)."""
    short = env_id.split("/")[-1]
    if short.endswith("-v0"):
        short = short[:-3]
    if short == "BittnerMultiGeneral":
        n = kwargs.pop("N", None)
        if n is None:
            raise ValueError("BittnerMultiGeneral needs N=")
        return _bittner(int(n), root, **kwargs)
    if short.startswith("Bittner-"):
        return _bittner(int(short.split("-")[1]), root, **kwargs)
    if short in ("PBNEnv", "PBN"):
        return PBNEnv(**kwargs)
    if short == "ControlPBNEnv":
        return ControlPBNEnv(**kwargs)
    raise ValueError("unknown env id %r (known: %s)" % (env_id, ", ".join(ENV_IDS)))


def register_gym_ids() -> bool:
    """Register the ids with gymnasium/gym if one of them is importable (``import gym_PBN`` does this
    as a side effect in the reference, train_BDQ.py:7-8).  Returns False when neither is installed."""
    try:
        import gymnasium as gym
    except ImportError:
        try:
            import gym  # type: ignore
        except ImportError:
            return False
    for env_id in ENV_IDS:
        short = env_id.split("/")[-1]
        try:
            gym.register(id=env_id, entry_point=lambda _id=env_id, **kw: make(_id, **kw), disable_env_checker=True)
        except Exception:
            pass
        del short
    return True
