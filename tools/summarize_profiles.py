"""Turns the raw ncu outputs brought back in gpurun_out/ into the small tracked
summaries under profiles/.

  python tools/summarize_profiles.py launches gpurun_out/launches_v6.csv profiles/r01_launches_summary.csv "note"
  python tools/summarize_profiles.py full gpurun_out/prof_commit_v6.ncu-rep profiles/r01_ncu_k_commit_summary.csv "note"
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEEP = (
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
    "l1tex__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "sm__cycles_active.avg", "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
)
STALL = "smsp__average_warps_issue_stalled_"


def launches(src, dst, note):
    rows = [l for l in open(src) if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v_ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
        k = r["Kernel Name"].split("(")[0]
        tot[k] += v_ms
        cnt[k] += 1
    all_ms = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# {note}\n")
        f.write("kernel,launches,total_ms,share\n")
        for k in sorted(tot, key=lambda k: -tot[k]):
            f.write(f"{k},{cnt[k]},{tot[k]:.3f},{tot[k] / all_ms:.4f}\n")


def full(src, dst, note):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rd[0], rd[1], rd[2]
    with open(dst, "w") as f:
        f.write(f"# {note}\n")
        for h, u, v in zip(hdr, units, vals):
            name = h.split(".TriageCompute.")[-1] if ".TriageCompute." in h else h
            if name in KEEP or name.startswith(STALL) and name.endswith("_not_issued.ratio") is False and name.endswith(".ratio"):
                f.write(f"{name},{v},{u}\n")


if __name__ == "__main__":
    mode, src, dst, note = sys.argv[1:5]
    {"launches": launches, "full": full}[mode](src, dst, note)
