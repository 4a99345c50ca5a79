"""PDL sequences vs plain steps: first mismatching step / tile (debug aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from pbn_rl_b200 import VecPBNEnv
net, attrs = bench.load_workload("pbn28")
e = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = VecPBNEnv(net, e, attrs, device="cuda:0", perturb_p=0.01, pdl=True, auto_reset=True)
b = VecPBNEnv(net, e, attrs, device="cuda:0", perturb_p=0.01, auto_reset=True)
g = torch.Generator(device="cuda").manual_seed(0)
s0 = torch.randint(0, 1 << 28, (e, 1), generator=g, device="cuda", dtype=torch.int64)
tgt = torch.randint(0, 14, (e,), generator=g, device="cuda", dtype=torch.int32)
for env in (a, b):
    env.state.copy_(s0); env.set_target(tgt)
bad = 0
for k in range(12):
    act = torch.randint(0, 29, (e, 3), generator=g, device="cuda", dtype=torch.uint8)
    a.step(act); b.step(act)
    if k % 4 == 3:
        a.advance_counter()
    torch.cuda.synchronize()
    neq = (a.state != b.state).any(dim=1)
    if neq.any():
        idx = torch.nonzero(neq).reshape(-1)
        print("step %d: %d envs differ, tiles %s" % (k, idx.numel(), sorted(set((idx // 1024).tolist()))[:10]))
        bad += 1
        a.state.copy_(b.state); a.target_id.copy_(b.target_id); a.t.copy_(b.t)
print("mismatching steps:", bad)
