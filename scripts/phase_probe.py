#!/usr/bin/env python
"""Phase timeline of CTA 0 of the sliced kernel (flag bit 31 -> %globaltimer stamps).  Development aid."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ctypes as C
import torch
import bench
from pbn_rl_b200 import VecPBNEnv, _cabi

net, attrs = bench.load_workload(sys.argv[1] if len(sys.argv) > 1 else "pbn28")
names = ["A0", "stage", "C1", "A1", "wait1", "B(+C1 odd)", "wait2+C2", "wait3", "E", "D", "F", "G+stats", "loopsync", "bump", "end"]
for envs in (1024, 1 << 20):
    env = VecPBNEnv(net, envs, attrs, device="cuda:0", auto_reset=True, **bench.ENV_KW)
    env.state[:, 0] = torch.randint(0, 1 << 28, (envs,), device="cuda")
    env.set_target(torch.randint(0, len(attrs), (envs,), device="cuda", dtype=torch.int32))
    acts = torch.randint(0, 29, (envs, 3), device="cuda", dtype=torch.uint8)
    ntile = (envs + 1023) // 1024
    final = torch.zeros((envs * net.n_words + 16 + 2 * ntile,), dtype=torch.int64, device="cuda")
    for rep in range(4):
        a = env._args(acts, final, True)
        a.flags |= 0x80000000
        _cabi.check(env.lib.pbn_step(env._h, C.byref(a), env._stream()))
        env.step_ctr += 1
        torch.cuda.synchronize()
    ts = final[envs * net.n_words:envs * net.n_words + 15].cpu().tolist()
    print("E=%d  total %.2f us" % (envs, (ts[14] - ts[0]) / 1e3))
    print("  " + "  ".join("%s %.2f" % (names[i], (ts[i + 1] - ts[i]) / 1e3) for i in range(14)))
    base = envs * net.n_words + 16
    st = final[base:base + 2 * ntile].cpu().numpy().reshape(ntile, 2)
    import numpy as np
    t0 = (st[:, 0] & 0x00FFFFFFFFFFFFFF).astype(np.int64)
    t1 = (st[:, 1] & 0x00FFFFFFFFFFFFFF).astype(np.int64)
    sm = (st[:, 0] >> 56) & 0xFF
    z = t0.min()
    print("  CTAs %d: start min/median/max %.2f %.2f %.2f us; end min/median/max %.2f %.2f %.2f us; duration median %.2f max %.2f" % (
        ntile, 0, np.median(t0 - z) / 1e3, (t0.max() - z) / 1e3, (t1.min() - z) / 1e3, np.median(t1 - z) / 1e3, (t1.max() - z) / 1e3,
        np.median(t1 - t0) / 1e3, (t1 - t0).max() / 1e3))
    per_sm = np.bincount(sm.astype(int), minlength=148)
    print("  CTAs per SM: min %d max %d" % (per_sm[per_sm > 0].min(), per_sm.max()))
