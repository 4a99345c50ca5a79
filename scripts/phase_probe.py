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
for envs in (1 << 20,):
    kw = dict(bench.ENV_KW)
    if len(sys.argv) > 2:
        kw["seed"] = int(sys.argv[2], 0)
    if len(sys.argv) > 3:
        kw["perturb_p"] = float(sys.argv[3])
    env = VecPBNEnv(net, envs, attrs, device="cuda:0", auto_reset=True, **kw)
    env.state[:, 0] = torch.randint(0, 1 << 28, (envs,), device="cuda")
    env.set_target(torch.randint(0, len(attrs), (envs,), device="cuda", dtype=torch.int32))
    acts = torch.randint(0, 29, (envs, 3), device="cuda", dtype=torch.uint8)
    ntile = (envs + 1023) // 1024
    final = torch.zeros((envs * net.n_words + 16 + 8 * ntile,), dtype=torch.int64, device="cuda")
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
    st = final[base:base + 8 * ntile].cpu().numpy().reshape(ntile, 8)
    import numpy as np
    t0 = (st[:, 0] & 0x00FFFFFFFFFFFFFF).astype(np.int64)
    t1 = (st[:, 7] & 0x00FFFFFFFFFFFFFF).astype(np.int64)
    tx = [(st[:, j] & 0x00FFFFFFFFFFFFFF).astype(np.int64) for j in range(8)]
    ta = (st[:, 1] & 0x00FFFFFFFFFFFFFF).astype(np.int64)
    tb = (st[:, 2] & 0x00FFFFFFFFFFFFFF).astype(np.int64)
    sm = (st[:, 0] >> 56) & 0xFF
    z = t0.min()
    print("  CTAs %d: start min/median/max %.2f %.2f %.2f us; end min/median/max %.2f %.2f %.2f us; duration median %.2f max %.2f" % (
        ntile, 0, np.median(t0 - z) / 1e3, (t0.max() - z) / 1e3, (t1.min() - z) / 1e3, np.median(t1 - z) / 1e3, (t1.max() - z) / 1e3,
        np.median(t1 - t0) / 1e3, (t1 - t0).max() / 1e3))
    per_sm = np.bincount(sm.astype(int), minlength=148)
    print("  CTAs per SM: min %d max %d" % (per_sm[per_sm > 0].min(), per_sm.max()))
    if ntile > 1:
        dur = (t1 - t0) / 1e3
        end = (t1 - z) / 1e3
        print("  end-time percentiles (us): " + "  ".join("p%d %.1f" % (q, np.percentile(end, q)) for q in (1, 10, 25, 50, 75, 90, 99, 100)))
        for k in sorted(set(per_sm[per_sm > 0])):
            sel = np.isin(sm, np.nonzero(per_sm == k)[0])
            print("  SMs with %d CTAs: %d CTAs, end median %.1f max %.1f" % (k, sel.sum(), np.median(end[sel]), end[sel].max()))
        for label, x in (("input copy arrived", ta - t0), ("arrived -> out planes done", tb - ta), ("out planes -> end", t1 - tb)):
            x = x / 1e3
            print("  %-28s p1 %.1f  p50 %.1f  p90 %.1f  p99 %.1f  max %.1f" % (label, np.percentile(x, 1), np.median(x), np.percentile(x, 90), np.percentile(x, 99), x.max()))
        for label, j0, j1 in (("E (planes->rows)", 2, 3), ("D (apply events)", 3, 4), ("F (outputs)", 4, 5), ("G (reset+stats)", 5, 6), ("final sync", 6, 7)):
            x = (tx[j1] - tx[j0]) / 1e3
            print("  %-28s p1 %.2f  p50 %.2f  p90 %.2f  p99 %.2f  max %.2f" % (label, np.percentile(x, 1), np.median(x), np.percentile(x, 90), np.percentile(x, 99), x.max()))
        dd = (tx[4] - tx[3]) / 1e3
        print("  tiles with D > 2 us: %d (%s ...)" % ((dd > 2).sum(), np.nonzero(dd > 2)[0][:12].tolist()))
        par = np.arange(ntile) & 1
        for pv in (0, 1):
            m = par == pv
            print("  %s tiles: arrive p50 %.1f p90 %.1f | arrive->planes p50 %.1f p90 %.1f | planes->end p50 %.1f p90 %.1f | end p50 %.1f p90 %.1f max %.1f" % (
                "even" if pv == 0 else "odd ", np.median((ta - t0)[m]) / 1e3, np.percentile((ta - t0)[m], 90) / 1e3,
                np.median((tb - ta)[m]) / 1e3, np.percentile((tb - ta)[m], 90) / 1e3, np.median((t1 - tb)[m]) / 1e3,
                np.percentile((t1 - tb)[m], 90) / 1e3, np.median(end[m]), np.percentile(end[m], 90), end[m].max()))
        n_even = np.array([int(((sm == s_) & (par == 0)).sum()) for s_ in range(148)])
        print("  even tiles per SM histogram:", np.bincount(n_even).tolist())
        for ne in sorted(set(n_even.tolist())):
            sel_sm = np.nonzero(n_even == ne)[0]
            ends = [end[sm == s_].max() for s_ in sel_sm if (sm == s_).any()]
            print("    SMs with %d even tiles: %d SMs, last-CTA end median %.1f max %.1f" % (ne, len(ends), np.median(ends), max(ends)))
        sm_end = np.array([end[sm == s].max() for s in range(148) if (sm == s).any()])
        sm_first = np.array([end[sm == s].min() for s in range(148) if (sm == s).any()])
        print("  per-SM last-CTA end: min %.1f median %.1f max %.1f;  per-SM first-CTA end: min %.1f median %.1f max %.1f" % (
            sm_end.min(), np.median(sm_end), sm_end.max(), sm_first.min(), np.median(sm_first), sm_first.max()))
        order = np.argsort(sm_end)
        print("  slowest SMs:", [(int(np.arange(148)[order[-i]]), round(float(sm_end[order[-i]]), 1)) for i in range(1, 6)],
              " fastest:", [(int(np.arange(148)[order[i]]), round(float(sm_end[order[i]]), 1)) for i in range(5)])
