#!/usr/bin/env python
"""Turn the ncu outputs of a gpurun call into the text summaries committed under profiles/.

    python scripts/profile_extract.py launches gpurun_out/launches.csv "<command>"  > profiles/rNN_launches_bench_summary.txt
    python scripts/profile_extract.py metrics  gpurun_out/raw.csv "<command>"       > profiles/rNN_step_sliced_ncu_metrics.txt
(the SASS-region summary comes from scripts/ncu_summary.py raw.csv src.csv <tiles>)."""
import collections
import csv
import sys


def launches(path, cmd):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, data = rows[0], rows[1:]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in data:
        name = r[kn].split("(")[0][:70]
        tot[name] += float(r[mv].replace(",", "")) / 1e3
        cnt[name] += 1
    total = sum(tot.values())
    print("ncu --metrics gpu__time_duration.sum --clock-control none -c %d %s" % (len(data), cmd))
    print("(first %d launches of the process; cold-cache, serialised: compare shares, not absolutes)\n" % len(data))
    for name in sorted(tot, key=lambda k: -tot[k])[:14]:
        print("%-72s n=%4d total %9.1f us avg %8.2f us share %5.1f%%" % (name, cnt[name], tot[name], tot[name] / cnt[name], 100 * tot[name] / total))


def metrics(path, cmd):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
            "sm__cycles_elapsed.avg.per_second")
    print("ncu --set full --clock-control none --import-source on -k regex:pbn_step_sliced -s 100 -c %d %s" % (len(data), cmd))
    print("(consecutive launches inside the bench; numbers under ncu are not bench values)\n")
    for i, h in enumerate(hdr):
        if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and "pcsamp" not in h):
            print("%-95s %-15s %s" % (h, units[i], "  ".join(r[i] for r in data)))


if __name__ == "__main__":
    {"launches": launches, "metrics": metrics}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
