"""pbn_step_host, packed form: time per step for several lane counts (PBN_B200_HOST_LANES).  Tuning aid."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from bench import load_workload, ENV_KW
from pbn_rl_b200 import VecPBNEnv

dev = torch.device("cuda:0")
net, attrs = load_workload("pbn28")
E = 1 << 20
env = VecPBNEnv(net, E, attrs, device=dev, auto_reset=True, device_counter="--host-counter" not in sys.argv, **ENV_KW)
env.reset()
pa = torch.randint(0, 29, (E, 3), dtype=torch.uint8)
a16 = env.pinned_actions16()
a16.numpy().view(np.uint16)[...] = env.pack_actions16(pa.numpy())
for lanes in (1, 2, 3, 4, 6, 8):
    os.environ["PBN_B200_HOST_LANES"] = str(lanes)
    for _ in range(5):
        env.step_host(None, chunks=0, compact="packed", actions16=a16)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 40
    for _ in range(n):
        env.step_host(None, chunks=0, compact="packed", actions16=a16)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / n
    print("lanes %d: %.3f ms  %.3e env-steps/s" % (lanes, t * 1e3, E / t), flush=True)
