"""Attractor discovery on the golden networks: what is found, how long it takes (docs aid)."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from bench import load_workload
from pbn_rl_b200.discover import find_attractors_rollout
for name, nr in (("pbn7", 1 << 12), ("pbn10", 1 << 12), ("pbn28", 1 << 16), ("pbn28", 1 << 20), ("pbn70", 1 << 16), ("pbn70", 1 << 20)):
    net, attrs = load_workload(name)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    found, info = find_attractors_rollout(net, n_rollouts=nr, burn_in=300, max_attractor_states=1 << 16)
    dt = time.perf_counter() - t0
    sizes = [len(m) for m in info["states"]]
    line = {"net": name, "rollouts": nr, "seconds": round(dt, 2), "attractors": len(found), "sizes": sizes[:40],
            "basin_top": [round(x, 4) for x in sorted(info["basin_fraction"], reverse=True)[:8]],
            "basin_sum": round(sum(info["basin_fraction"]), 4), "distinct_end_states": info["distinct_end_states"],
            "unresolved": len(info["unresolved"])}
    if name == "pbn28":
        members = {s for m in info["states"] for s in m}
        k3 = [sum(int(b) << i for i, b in enumerate(a[0])) for a in attrs.attractors]
        line["k3_states_inside_found_attractors"] = sum(1 for s in k3 if s in members)
    print(json.dumps(line), flush=True)
