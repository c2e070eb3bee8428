#!/usr/bin/env python
"""bench.py — placement decisions / second of the NodeSelect hot path.

A "step" is one full scheduling tick (SchedulerAlgo::NodeSelect,
JobScheduler.cpp:5543-5868) over one synthetic pending queue. At N=1 the
workload is BASELINE.json configs[1]: 100k pending jobs x 10k nodes, cpu+mem+
gres(GPU), 4 partitions, multifactor priority + backfill. At N>1 the queue
shards by partition: rank r schedules its own config-2-shaped
set of partitions (a disjoint slice of an N-times larger cluster; weak scaling,
no data-path collective inside the tick; NCCL is used for the barrier and the
max-over-ranks timing reduction).

  value   decisions/s, tables resident in HBM when the timed region starts
          (device time of crane_sched_run, CUDA events on the launching stream)
  e2e     same metric through crane_sched_node_select with HOST buffers
          (H2D of the pending table + D2H of the placements inside the region)
  --impl reference   the CPU oracle (oracle/, a restatement of the reference's
          single-threaded NodeSelect; the reference itself cannot be built in
          this image) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "placement decisions/sec (jobs through the NodeSelect job loop)"
UNIT = "decisions/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in o.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of k_commit from the committed
    `ncu --set full` capture (profiles/r01_ncu_k_commit_summary.csv), per launch."""
    p = os.path.join(ROOT, "profiles", "r01_ncu_k_commit_summary.csv")
    if not os.path.exists(p):
        return None
    tot = 0.0
    for line in open(p):
        f = line.strip().split(",")
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(f[2], 1.0)
            tot += float(f[1]) * mult
    return tot or None


def algorithmic_bytes(cluster, pending, n_decided):
    """B_dec(j) = job_row + M_part(j) * node_row + out_row(j),
    job_row 64 B, node_row 48 B (24 B when the cluster has no gres), out_row
    16 B + 32 B * node_num."""
    part_size = np.diff(cluster.part_off.astype(np.int64))
    node_row = 48 if cluster.n_gres_entries else 24
    p = pending.partition
    ok = p < cluster.n_partitions
    mp = np.where(ok, part_size[np.minimum(p, cluster.n_partitions - 1)], 0)
    b = 64 + mp * node_row + 16 + 32 * pending.node_num.astype(np.int64)
    return int(b.sum()), node_row


def workload(args, rank):
    """Rank r's shard: its own config-shaped set of partitions (weak scaling)."""
    from cranesched_b200 import sharding
    if args.config not in (1, 2, 5):
        raise SystemExit("bench.py --config must be 1, 2 or 5")
    return sharding.shard_workload(args.config, rank, int(os.environ.get("WORLD_SIZE", "1")), args.jobs, args.nodes)


def workload_name(args):
    return {1: "config1: 1k jobs x 128 nodes, cpu+mem, FIFO",
            2: "config2: %dk pending jobs x %dk nodes, cpu+mem+gres(GPU), 4 partitions, multifactor priority + backfill"
               % ((args.jobs or 100_000) // 1000, (args.nodes or 10_000) // 1000),
            5: "config5: backfill stress, %dk jobs x %dk nodes, 1 partition"
               % ((args.jobs or 200_000) // 1000, (args.nodes or 5_000) // 1000)}[args.config]


def run_reference(args, rank, world):
    """CPU arm: the oracle, single thread like the reference's NodeSelect
    ("TODO: do it in parallel", JobScheduler.cpp:5756,5776)."""
    if rank != 0:
        return
    from oracle import pyoracle
    pyoracle.build()
    cfg, cl, rn, pd, now = workload(args, 0)
    sample = args.ref_sample
    times, done = [], 0
    for i in range(args.warmup + args.steps):
        _, ms, done = pyoracle.node_select(cfg, cl, rn, pd, now, max_jobs=sample)
        if i >= args.warmup:
            times.append(ms)
    ms = float(np.mean(times))
    v = done / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64+u64 masks, f64 cost/priority",
            "data": "synthetic", "config": {"workload": workload_name(args)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": "first %d jobs of the priority order of the same queue (cost per job grows as "
                                       "the cluster fills, so this flatters the CPU)" % done},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--jobs", type=int, default=0)
    ap.add_argument("--nodes", type=int, default=0)
    ap.add_argument("--ref-sample", type=int, default=1500, help="jobs per step of the CPU arm")
    ap.add_argument("--cpu-sample", type=int, default=1500, help="jobs of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from cranesched_b200 import abi
    from cranesched_b200.scheduler import GpuScheduler

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg, cl, rn, pd, now = workload(args, rank)
    sched = GpuScheduler(cfg, local_rank)
    sched.set_cluster(cl)
    out = abi.Placements.for_pending(pd, pinned=True)
    # pinned copies of the pending table for the e2e leg
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sched.upload(rn, pd)
    sched.sync()
    dev_ms = []
    timing_last = None

    def step(record):
        flush.fill_(1)  # L2 flush between timed iterations
        torch.cuda.synchronize()
        sched.run(now)
        ms = sched.sync()
        if record:
            dev_ms.append(ms)

    for _ in range(args.warmup):
        step(False)
    barrier()
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(True)
        barrier()
        wall = time.perf_counter() - t0
    timing_last = sched.timing()
    sched.fetch(out)
    n_decided = int(min(pd.n, cfg.scheduled_batch_size))
    dev_total_ms = float(np.sum(dev_ms))

    # ---- e2e: host buffers through crane_sched_node_select --------------------
    e2e_ms = []
    for i in range(2 + args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        sched.node_select(now, rn, pd, out)
        dt = (time.perf_counter() - t1) * 1e3
        if i >= 2:
            e2e_ms.append(dt)
    barrier()
    h2d = sum(getattr(pd, f).nbytes for f in pd.__dataclass_fields__ if getattr(pd, f) is not None)
    d2h = out.nbytes()

    # max over ranks
    t_dev = torch.tensor([dev_total_ms, float(np.sum(e2e_ms)), wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_total_ms, e2e_total_ms, wall_ms = [float(x) for x in t_dev.tolist()]
    total_decided = n_decided * world
    value = total_decided * args.steps / (dev_total_ms / 1e3)
    e2e_value = total_decided * args.steps / (e2e_total_ms / 1e3)

    if rank == 0:
        peak, peak_src = load_peaks()
        alg_bytes, node_row = algorithmic_bytes(cl, pd, n_decided)
        commit_ms = timing_last["commit_ms"]
        achieved = alg_bytes / (commit_ms / 1e3) / 1e9
        placed_now = int((out.reason == 0).sum())
        reserved = int(((out.reason != 0) & (out.n_alloc > 0)).sum())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "int64 cpu / u64 mem + bit masks; f64 cost and priority", "data": "synthetic",
            "config": {"workload": workload_name(args), "per_gpu_jobs": pd.n, "per_gpu_nodes": cl.n_nodes,
                       "partitions_per_gpu": cl.n_partitions, "sharding": "by partition (independent LocalSchedulers); every rank schedules its own copy of the same synthetic draw (equal per-GPU work)",
                       "l2": "256 MiB flush write between timed iterations",
                       "started_now": placed_now, "backfill_reserved": reserved},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_total_ms / args.steps},
            "gpu_launches": int(timing_last["kernel_launches"]) * args.steps,
            "phases_ms": {k: round(float(v), 4) for k, v in timing_last.items() if k.endswith("_ms")},
            "wall_ms_per_step_incl_flush": wall_ms / args.steps,
            "roofline": {"bound": "hbm", "kernel": "k_commit", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic() if args.config == 2 and not args.jobs else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "per decision 64 B job row + M_part x %d B node row + 16 B + 32 B x node_num "
                                 "(SURVEY.md 8d); the job loop is a dependency chain, so the binding limit is "
                                 "per-job latency, not DRAM" % node_row},
            "clocks": clocks.summary(),
        }
        if not args.no_cpu_baseline and world == 1:
            from oracle import pyoracle
            pyoracle.build()
            _, ms, done = pyoracle.node_select(cfg, cl, rn, pd, now, max_jobs=args.cpu_sample)
            line["cpu_baseline"] = {
                "value": done / (ms / 1e3), "unit": UNIT, "cores": 1, "kind": "port",
                "sample": "oracle (CPU restatement of NodeSelect, 1 thread like the reference) on the first %d jobs "
                          "of the priority order of the same queue, %.1f s" % (done, ms / 1e3)}
        print(json.dumps(line))
    sched.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
