#!/usr/bin/env python
"""PBN env-steps/s benchmark (BASELINE.json metric): Bittner-28, 2^20 env instances per GPU.

    python bench.py --gpus 1 --steps K --warmup W               # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps K --warmup W   # CPU arm (oracle port on host cores)
    torchrun ... bench.py --gpus N ...                           # one rank per GPU, weak scaling

A "step" is one pbn_step launch over one batch of 2^20 env instances (per GPU): apply the
agent's gene flips, draw predictor selections + perturbations from Philox, synchronous
Boolean update, target-attractor test, reward/done, auto-reset.  Prints ONE JSON line.

Timing: steps are captured into a CUDA graph (the launch-bound inner loop; the Philox step
counter lives in device memory so every replay advances the streams) and replayed; the launches use
programmatic dependent launch, so a step draws its state-independent predictor-selection planes
while the previous kernel drains and only then waits for it; the timed
region is bracketed by barrier + torch.cuda.synchronize() and CUDA events on the launching
stream; max over ranks.  The L2 (126 MB) is defeated by rotating over R independent env
batches whose combined working set exceeds it, plus a pool of distinct action buffers.
"""
import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLD = ROOT / "tests" / "golden"
METRIC = "PBN env-steps/sec (Bittner-28, 2^20 envs) at 1/2/4/8 B200 vs host CPU"
UNIT = "env-steps/s"
BYTES_PER_STEP = {1: 33, 2: 49}  # algorithmic HBM bytes per env-step by words/state


# ----------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------

def load_workload(name):
    """Network + attractor/target table for a golden network (same choices as tests/helpers.py)."""
    import numpy as np
    from pbn_rl_b200 import AttractorSet, PBNNetwork, sorted_id_permutation
    net = PBNNetwork.from_json(GOLD / f"{name}.json")
    if name == "pbn28":
        raw = json.loads((GOLD / "attractors_bittner28.json").read_text())["attractors"]
        attrs = AttractorSet([[tuple(s) for s in a] for a in raw], 28).permuted(sorted_id_permutation(net.genes))
    elif name == "pbn7":
        raw = json.loads((GOLD / "attractors_bittner7.json").read_text())["attractors"]
        attrs = AttractorSet([[tuple(s) for s in a] for a in raw], 7)
    elif name == "pbn10":
        sinks = json.loads((GOLD / "k5_stg.json").read_text())["pbn10"]["sink_sccs"]
        attrs = AttractorSet([[tuple((s >> i) & 1 for i in range(10)) for s in m] for m in sinks], 10)
    else:
        rng = np.random.default_rng(70)
        states = rng.integers(0, 2, size=(16, net.n_genes))
        attrs = AttractorSet([[tuple(int(v) for v in s)] for s in states], net.n_genes)
    return net, attrs


ENV_KW = dict(horizon=20, bins=3, perturb_p=0.001, perturb_mode="A", r_success=5.0, r_step=0.0, r_action=-1.0,
              seed=0x5EED)


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle's per-instance env on the host cores
# ----------------------------------------------------------------------------------------------

_CPU_ENV = {}


def _cpu_env(name, seed):
    """Per-process cached oracle env (built once per worker process)."""
    key = (name, seed)
    if key not in _CPU_ENV:
        import numpy as np
        from oracle import pbn_oracle as O
        onet = O.OracleNetwork.from_json(GOLD / f"{name}.json")
        _, attrs = load_workload(name)
        env = O.OraclePBNEnv(onet, attrs.attractors, horizon=ENV_KW["horizon"], perturb_p=ENV_KW["perturb_p"],
                             perturb_mode=ENV_KW["perturb_mode"], r_success=ENV_KW["r_success"],
                             r_step=ENV_KW["r_step"], r_action=ENV_KW["r_action"], seed=seed)
        env.reset()
        acts = np.random.default_rng(seed).integers(0, onet.n + 1, size=(4096, 3)).tolist()
        _CPU_ENV[key] = (env, acts)
    return _CPU_ENV[key]


def _cpu_worker(args):
    """Step one single-instance oracle env `n_steps` times (or for `seconds`) with random actions."""
    name, seed, n_steps, seconds = args
    env, acts = _cpu_env(name, seed)
    chunk = 64
    done_steps = 0
    t0 = time.perf_counter()
    while True:
        for i in range(chunk):
            _, _, term, trunc, _ = env.step(acts[(done_steps + i) & 4095])
            if term or trunc:
                env.reset()
        done_steps += chunk
        if n_steps is not None and done_steps >= n_steps:
            break
        if seconds is not None and time.perf_counter() - t0 >= seconds:
            break
    return done_steps, time.perf_counter() - t0


class CpuArm:
    """`cores` worker processes, one oracle env each; run() = one bounded sample on all of them."""

    def __init__(self, name, cores):
        import multiprocessing as mp
        self.name, self.cores = name, cores
        self.pool = mp.get_context("fork").Pool(cores) if cores > 1 else None

    def run(self, n_steps=None, seconds=None):
        jobs = [(self.name, 1000 + i, n_steps, seconds) for i in range(self.cores)]
        res = self.pool.map(_cpu_worker, jobs, chunksize=1) if self.pool else [_cpu_worker(jobs[0])]
        return sum(r[0] for r in res), max(max(r[1] for r in res), 1e-9)

    def close(self):
        if self.pool:
            self.pool.close()
            self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    K, Wm = args.steps_ref, args.warmup_ref
    # bounded sample per bench "step", sized so K + W steps take about a minute of wall clock
    # (the per-instance python env does roughly 2e4 env-steps/s per core)
    per_proc = int(min(4096, max(64, (60.0 * 2.0e4) / max(K + Wm, 1))) // 64 * 64)
    arm = CpuArm(args.net, cores)
    for _ in range(Wm):
        arm.run(n_steps=per_proc)
    total, busy = 0, 0.0
    for _ in range(K):
        n, t = arm.run(n_steps=per_proc)
        total += n
        busy += t
    arm.close()
    value = total / busy
    sample = "%d processes x %d env-steps per step, single-env oracle port, %s" % (cores, per_proc, args.net)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": Wm, "ms_per_step": 1e3 * busy / max(K, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, int(os.environ.get("WORLD_SIZE", "1")), None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------

def workload_config(args, world, kernel):
    return {
        "workload": "Bittner-28 PBN (kaban/pbn28.ispl, data/attractors_Bittner-28.pkl), 2^20 env instances per GPU"
        if args.net == "pbn28" and args.envs == 1 << 20 else "%s, %d env instances per GPU" % (args.net, args.envs),
        "network": args.net, "envs_per_gpu": args.envs, "total_envs": args.envs * world, "bins": 3,
        "horizon": 20, "perturb_p": 0.001, "perturb_mode": "A", "auto_reset": True,
        "actions": "uniform in [0,N], pre-generated on device, pool of %d buffers" % args.action_pool,
        "l2": "%d rotating env batches + action pool: working set > 126 MB L2, no flush" % args.batches,
        "graph_steps": min(args.graph_steps, max(1, args.steps or 1)), "kernel": kernel,
        "launch": "CUDA graph of %d step launches%s" % (min(args.graph_steps, max(1, args.steps or 1)), "" if getattr(args, "no_pdl", False) else
                                                        ", programmatic dependent launch (selection planes drawn under the previous kernel's tail)"), "parallelism": "env-sharded x%d" % world,
    }


def make_env(net, attrs, args, device, env_offset, device_counter=True):
    import torch
    from pbn_rl_b200 import VecPBNEnv
    env = VecPBNEnv(net, args.envs, attrs, device=device, env_offset=env_offset, auto_reset=True,
                    device_counter=device_counter, pdl=not args.no_pdl, kernel=args.kernel, **ENV_KW)
    g = torch.Generator(device=device).manual_seed(env_offset + 1)
    n = net.n_genes
    for w in range(net.n_words):
        bits = min(64, n - 64 * w)
        hi = torch.randint(0, 1 << max(bits - 31, 0), (args.envs,), generator=g, device=device, dtype=torch.int64)
        lo = torch.randint(0, 1 << min(bits, 31), (args.envs,), generator=g, device=device, dtype=torch.int64)
        env.state[:, w] = (hi << 31) | lo
    env.set_target(torch.randint(0, len(attrs), (args.envs,), generator=g, device=device, dtype=torch.int32))
    return env


def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    net, attrs = load_workload(args.net)
    W = net.n_words

    # R independent env batches (each `envs` instances); batch b of rank r owns global env ids
    # [(r*R + b) * envs, ...): disjoint Philox streams everywhere.
    R = args.batches
    envs = [make_env(net, attrs, args, device, (rank * R + b) * args.envs) for b in range(R)]
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    pool = [torch.randint(0, net.n_genes + 1, (args.envs, 3), generator=g, device=device, dtype=torch.uint8)
            for _ in range(args.action_pool)]
    kernel = envs[0].kernel

    K = max(1, args.steps)                     # EXACTLY K timed steps: K // G replays of a G-step graph + one tail graph
    G = min(args.graph_steps, K)
    tail_steps = K % G
    Wm = max(3, args.warmup)
    stream = torch.cuda.Stream(device)
    stats_total = torch.zeros(8, dtype=torch.int64, device=device)

    def enqueue(i):
        envs[i % R].step(pool[i % len(pool)])

    with torch.cuda.stream(stream):
        for i in range(Wm):
            enqueue(i)
        for e in envs:
            e.advance_counter()
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for i in range(G):
                enqueue(i)
            for e in envs:
                e.advance_counter()   # PDL launches do not bump the device step counter themselves
        graph.replay()
        tail = None
        if tail_steps:
            tail = torch.cuda.CUDAGraph()
            with torch.cuda.graph(tail, stream=stream):
                for i in range(tail_steps):
                    enqueue(i)
                for e in envs:
                    e.advance_counter()
            tail.replay()
        stream.synchronize()

        sampler = ClockSampler(local_rank)
        sampler.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(K // G):
            graph.replay()
        if tail is not None:
            tail.replay()
        ev1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop()
        ms = ev0.elapsed_time(ev1)

    # episode statistics: the only cross-GPU exchange of the path (one small NCCL all-reduce)
    for e in envs:
        stats_total += e.stats_buf
    t_ms = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats_total, op=dist.ReduceOp.SUM)
    ms = float(t_ms.item())
    value = args.envs * world * K / (ms * 1e-3)
    ms_per_step = ms / K
    bytes_per = BYTES_PER_STEP[W]
    achieved = bytes_per * args.envs / (ms_per_step * 1e-3) / 1e9
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    # DRAM traffic of the step kernel per launch, from the committed ncu capture of this command
    traffic = None
    tfile = ROOT / "profiles" / "traffic.json"
    if tfile.exists():
        try:
            traffic = json.loads(tfile.read_text()).get(args.net, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    # end-to-end through the public API with host buffers (rank-local, every rank does it)
    import numpy as np
    e2e_env = envs[0]
    host_actions = []
    for p in pool[:4]:  # the step's inputs live in pinned host memory, as the contract asks
        pa = e2e_env.pinned_actions()
        pa.copy_(p)
        host_actions.append(pa)
    n_e2e = args.e2e_steps

    def time_e2e(compact):
        for i in range(3):
            e2e_env.step_host(host_actions[i % 4], compact=compact)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ee0.record()
        for i in range(n_e2e):
            out = e2e_env.step_host(host_actions[i % 4], compact=compact)
        ee1.record()
        torch.cuda.synchronize()
        secs = max(time.perf_counter() - t0, ee0.elapsed_time(ee1) * 1e-3)
        t_e2e = torch.tensor([secs], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        assert out["reward"].shape[0] == args.envs
        return args.envs * world * n_e2e / float(t_e2e.item())

    # full-width results (int64 state words, fp32 reward, terminated, truncated) and, for N <= 32, the
    # compact form of the same results (uint32 state, fp32 reward, done byte): fewer PCIe bytes per env-step
    e2e_full = time_e2e(False)
    compact_ok = net.n_genes <= 32
    e2e_value = time_e2e(True) if compact_ok else e2e_full
    h2d, d2h = e2e_env.host_bytes_per_step_compact if compact_ok else e2e_env.host_bytes_per_step
    e2e_note = {
        "results": "uint32 state + fp32 reward + done byte per env (step_host(compact=True))" if compact_ok
        else "int64 state words + fp32 reward + terminated + truncated per env",
        "full_width_value": e2e_full, "full_width_d2h_bytes_per_step": e2e_env.host_bytes_per_step[1],
        "transfer": "pbn_step_host: actions uploaded from pinned memory by the copy engine, results written "
                    "by an export kernel straight into pinned host memory (zero-copy over PCIe), 2 chunks pipelined",
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        arm = CpuArm(args.net, cores)
        n, busy = arm.run(seconds=args.cpu_seconds)
        arm.close()
        cpu = {"value": n / busy, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d processes x %.0f s of single-env oracle steps (%d env-steps), %s" % (cores, args.cpu_seconds, n, args.net)}

    if rank == 0:
        stats = dict(zip(("steps", "episodes", "terminated", "truncated", "ep_len_sum", "flips", "perturbed"),
                         stats_total.cpu().tolist()))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": workload_config(args, world, kernel),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_env_step": bytes_per,
                         "kernel_us": ms_per_step * 1e3},
            "cpu_baseline": cpu,
            "e2e": dict({"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                         "steps": n_e2e}, **e2e_note),
            "gpu_launches": K,
            "clocks": clocks,
            "episode_stats": stats,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # stdout carries exactly one JSON line: anything libraries print to fd 1 (NCCL's version banner, ...) is
    # sent to stderr instead
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--net", default="pbn28")
    ap.add_argument("--envs", type=int, default=1 << 20, help="env instances per GPU")
    ap.add_argument("--batches", type=int, default=8, help="independent env batches rotated to defeat L2")
    ap.add_argument("--action-pool", type=int, default=16)
    ap.add_argument("--graph-steps", type=int, default=64)
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pdl", action="store_true", help="plain stream-serialised launches instead of programmatic dependent launch")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps_ref = args.steps if args.steps is not None else 20
        args.warmup_ref = args.warmup if args.warmup is not None else 3
        run_reference(args)
        return
    args.steps = args.steps if args.steps is not None else 6400
    args.warmup = args.warmup if args.warmup is not None else 64
    run_gpu(args)


if __name__ == "__main__":
    main()
