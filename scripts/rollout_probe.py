#!/usr/bin/env python
"""Throughput of pbn_rollout (uncontrolled updates, states resident on chip) vs a loop of pbn_step."""
import json
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from pbn_rl_b200 import VecPBNEnv
from pbn_rl_b200.formats import network_from_bnet

E, S = 1 << 20, 256
rows = []
for name, p in (("pbn28", 0.0), ("pbn28", 0.001), ("pbn70", 0.001), ("pbn7", 0.001), ("bb33", 0.0)):
    if name == "bb33":
        net, attrs = network_from_bnet(ROOT / "tests" / "golden" / "bb33.bnet"), None
    else:
        net, attrs = bench.load_workload(name)
    env = VecPBNEnv(net, E, attrs, device="cuda:0", perturb_p=p, perturb_mode="A", horizon=0)
    env.state[:, 0] = torch.randint(0, 1 << min(net.n_genes, 62), (E,), device="cuda")
    env.rollout(8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.rollout(S); e1.record(); torch.cuda.synchronize()
    t_roll = e0.elapsed_time(e1) * 1e-3
    for _ in range(4):
        env.step(None)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(32):
        env.step(None)
    e1.record(); torch.cuda.synchronize()
    t_step = e0.elapsed_time(e1) * 1e-3 / 32
    rows.append({"net": name, "genes": net.n_genes, "perturb_p": p, "envs": E, "rollout_steps": S,
                 "rollout_us_per_update": 1e6 * t_roll / S, "rollout_env_steps_per_s": E * S / t_roll,
                 "step_loop_us_per_update": 1e6 * t_step, "step_loop_env_steps_per_s": E / t_step,
                 "speedup": t_step / (t_roll / S)})
    print(json.dumps(rows[-1]), flush=True)
    env.close()
