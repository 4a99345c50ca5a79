#!/usr/bin/env python
"""Per-CUDA-line executed-instruction / stall-sample summary of an .ncu-rep (needs the sources ncu names to exist:
symlink csrc/*.cuh and write net_gen.cuh / net_update.inc of the network into the repo root).  Development aid."""
import collections
import csv
import subprocess
import sys

rep, ntiles = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1024.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
for k in keys + [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]:
    if k in hdr:
        i = hdr.index(k)
        vals = [r[i] for r in rows[2:]]
        if any(v not in ("0", "0.000000") for v in vals):
            print(k.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), units[i], vals)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, agg = None, None, collections.OrderedDict()
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif r[0] not in ("", "Function Name") and hdr:
        try:
            n, sm = int(r[ie]), int(r[isamp])
        except Exception:
            continue
        key = (cur, int(r[0]), r[1][:100])
        old = agg.get(key, (0, 0))
        agg[key] = (old[0] + n, old[1] + sm)
tot = sum(v[0] for v in agg.values())
ts = max(sum(v[1] for v in agg.values()), 1)
print("total warp-instr %d = %.0f per tile, samples %d" % (tot, tot / ntiles, ts))
byfile = collections.Counter()
for (f, l, t), (n, sm) in agg.items():
    byfile[f] += n
print({k: round(v / ntiles) for k, v in byfile.items()})
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
for (f, l, t), (n, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-16s %4d %7.1f/tile %5.1f%% samp %5.1f%%  %s" % (f, l, n / ntiles, 100 * n / tot, 100 * sm / ts, t))
