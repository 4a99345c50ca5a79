#!/usr/bin/env python
"""Step time vs number of envs (latency floor vs throughput slope).  Development aid."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "scripts"))
import bench  # noqa: E402
from perf_probe import time_config  # noqa: E402

net_name = sys.argv[1] if len(sys.argv) > 1 else "pbn28"
kernel = sys.argv[2] if len(sys.argv) > 2 else "sliced"
net, attrs = bench.load_workload(net_name)
for envs in (1024, 16 * 1024, 148 * 1024, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
    us = time_config(net, attrs, envs, kernel, 0.001, True, True, True, batches=4 if envs >= (1 << 22) else 8)
    print("%-6s %-7s E=%8d  %8.2f us/step  %9.3e steps/s  %7.1f GB/s" % (
        net_name, kernel, envs, us, envs / us * 1e6, bench.BYTES_PER_STEP[net.n_words] * envs / us / 1e3), flush=True)
