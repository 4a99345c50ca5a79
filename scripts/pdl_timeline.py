#!/usr/bin/env python
"""Timeline of consecutive programmatic-dependent-launch steps (per-CTA %globaltimer stamps).  Development aid."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ctypes as C
import numpy as np
import torch
import bench
from pbn_rl_b200 import VecPBNEnv, _cabi

net, attrs = bench.load_workload("pbn28")
envs, nl = 1 << 20, 6
ntile = envs // 1024
pdl = "--nopdl" not in sys.argv
KW = dict(bench.ENV_KW)
if "--p0" in sys.argv:
    KW["perturb_p"] = 0.0
es = [VecPBNEnv(net, envs, attrs, device="cuda:0", auto_reset=True, device_counter=True, pdl=pdl, env_offset=b * envs, **KW) for b in range(nl)]
for e in es:
    e.state[:, 0] = torch.randint(0, 1 << 28, (envs,), device="cuda")
    e.set_target(torch.randint(0, len(attrs), (envs,), device="cuda", dtype=torch.int32))
acts = torch.randint(0, 29, (envs, 3), device="cuda", dtype=torch.uint8)
finals = [torch.zeros((envs + 16 + 8 * ntile,), dtype=torch.int64, device="cuda") for _ in range(nl)]
for rep in range(3):
    for k in range(nl):
        a = es[k]._args(acts, finals[k], True)
        a.flags |= 0x80000000
        _cabi.check(es[k].lib.pbn_step(es[k]._h, C.byref(a), es[k]._stream()))
    torch.cuda.synchronize()
T = []
RAW = []
for k in range(nl):
    raw = finals[k][envs + 16:envs + 16 + 8 * ntile].cpu().numpy().reshape(ntile, 8)
    RAW.append(raw.copy())
    st = raw[:, [0, 1, 2, 7]] & 0x00FFFFFFFFFFFFFF
    T.append(st.astype(np.int64))
import os
if os.path.isdir("gpurun_out"):
    np.save("gpurun_out/pdl_timeline%s.npy" % ("_p0" if "--p0" in sys.argv else ""), np.stack(RAW))
z = T[0][:, 0].min()
names = ("start", "input arrived", "out planes", "end")
for k in range(nl):
    d = T[k]
    print("   phases p50: start->input %.1f  input->planes %.1f  planes->end %.1f  (CTA lifetime p50 %.1f)" % (
        np.median(d[:, 1] - d[:, 0]) / 1e3, np.median(d[:, 2] - d[:, 1]) / 1e3, np.median(d[:, 3] - d[:, 2]) / 1e3, np.median(d[:, 3] - d[:, 0]) / 1e3))
    print("launch %d: " % k + " | ".join("%s min %.1f p50 %.1f max %.1f" % (names[j], (T[k][:, j].min() - z) / 1e3, (np.median(T[k][:, j]) - z) / 1e3, (T[k][:, j].max() - z) / 1e3) for j in range(4)))
print("period (end max to end max): " + " ".join("%.1f" % ((T[k + 1][:, 3].max() - T[k][:, 3].max()) / 1e3) for k in range(nl - 1)))
