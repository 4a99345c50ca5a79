"""One pbn_rollout launch (Bittner-28, 2^20 envs, 64 updates) -- target of the ncu capture in profiles/."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from pbn_rl_b200 import VecPBNEnv
net, attrs = bench.load_workload("pbn28")
env = VecPBNEnv(net, 1 << 20, attrs, device="cuda:0", perturb_p=0.001, perturb_mode="A", horizon=0)
env.state[:, 0] = torch.randint(0, 1 << 28, (1 << 20,), device="cuda")
env.rollout(4)
env.rollout(64)
torch.cuda.synchronize()
print("ok", env.stats())
