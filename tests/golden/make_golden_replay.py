#!/usr/bin/env python
"""Generates tests/golden/replay_expected.json by RUNNING the reference's own replay memory in this container
(/root/reference is read, never copied):

  * `ExperienceReplay` of /root/reference/bdq_model/memory.py (loaded by file path as a synthetic package, so the
    package __init__ with its `import gym` never runs): a scripted store / overwrite sequence and `sample()` under a
    seeded python `random`;
  * the tensor-building lines of `update_policy` (/root/reference/bdq_model/__init__.py:100-111), sliced out of the
    source file and executed on the reference's own samples (as make_golden_formats.py does for the parsers);
  * the network input `predict()` builds (bdq_model/__init__.py:92-93).

The fixture pins oracle/replay_oracle.py (tests/test_oracle_replay.py) and, through it, the device replay ring
(tests/test_gpu_replay.py).

    python tests/golden/make_golden_replay.py
"""
import importlib.util
import json
import random
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def load_reference_memory():
    pkg = types.ModuleType("refbdq")
    pkg.__path__ = [str(REF / "bdq_model")]
    sys.modules["refbdq"] = pkg
    spec = importlib.util.spec_from_file_location("refbdq.memory", REF / "bdq_model" / "memory.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refbdq.memory"] = mod
    spec.loader.exec_module(mod)
    return mod


def update_policy_tensor_lines():
    """Lines 100-111 of bdq_model/__init__.py: from `x = memory.sample(batch_size)` to `input_tuples = ...`."""
    src = (REF / "bdq_model" / "__init__.py").read_text().splitlines()
    a = next(i for i, ln in enumerate(src) if ln.strip().startswith("x = memory.sample(batch_size)"))
    b = next(i for i, ln in enumerate(src) if i > a and ln.strip().startswith("input_tuples = torch.stack"))
    body = [ln[8:] if ln.startswith("        ") else ln for ln in src[a:b + 1]]
    return "\n".join(body)


def main():
    mem_mod = load_reference_memory()
    n, capacity, bins = 5, 7, 3
    rng = random.Random(20261018)
    mem = mem_mod.ExperienceReplay(capacity)
    script = []
    for k in range(19):
        state = tuple(rng.randint(0, 1) for _ in range(n))
        target = tuple(rng.randint(0, 1) for _ in range(n))
        action = [rng.randint(0, n) for _ in range(bins)]
        reward = float(rng.choice([-3.0, -2.0, -1.0, 0.0, 2.0, 4.0, 5.0]))
        nxt = tuple(rng.randint(0, 1) for _ in range(n))
        done = rng.random() < 0.3
        mem.store(mem_mod.Transition(state, target, torch.tensor(action), reward, nxt, done))
        script.append(dict(state=state, target=target, action=action, reward=reward, next_state=nxt, done=done))
    buffer = [dict(state=t.state, target=t.target, action=t.action.tolist(), reward=t.reward, next_state=t.next_state,
                   done=bool(t.done)) for t in mem.buffer]
    out = {"n": n, "capacity": capacity, "bins": bins, "script": script, "buffer_after": buffer,
           "current_index": mem.current_index, "len": len(mem), "samples": []}
    code = update_policy_tensor_lines()
    for seed, batch in ((1, 4), (2, 7), (3, 1)):
        random.seed(seed)
        sampled = random.sample(range(len(mem.buffer)), batch)   # what memory.sample draws: random.sample(self.buffer, k)
        random.seed(seed)
        env = {"memory": mem, "batch_size": batch, "np": np, "torch": torch,
               "self": types.SimpleNamespace(config=types.SimpleNamespace(device="cpu"))}
        exec(code, env)
        assert [mem.buffer[i] for i in sampled] == env["x"]       # same draw as indices
        out["samples"].append({
            "seed": seed, "batch": batch, "index": sampled,
            "states": env["states"].tolist(), "targets": env["targets"].tolist(), "actions": env["actions"].tolist(),
            "rewards": env["rewards"].tolist(), "next_states": env["next_states"].tolist(), "masks": env["masks"].tolist(),
            "input_tuples_shape": list(env["input_tuples"].shape), "actions_shape": list(env["actions"].shape),
        })
    # predict(): np.stack((state, target)) -> float tensor [2, N] (unsqueezed to [2, 1, N] for the network)
    st, tg = script[0]["state"], script[0]["target"]
    out["predict_input"] = torch.tensor(np.stack((st, tg))).float().tolist()
    (OUT / "replay_expected.json").write_text(json.dumps(out, indent=1))
    print("wrote", OUT / "replay_expected.json", "with", len(out["samples"]), "samples; tensor code was:\n" + code)


if __name__ == "__main__":
    main()
