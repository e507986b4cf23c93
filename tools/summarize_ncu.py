"""Developer tool: turns ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.
usage: python tools/summarize_ncu.py <tag> <launches.csv> [<report.ncu-rep> ...]"""
import collections
import csv
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if len(r) > 10 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            v = float(d["Metric Value"].replace(",", ""))
            u = d["Metric Unit"]
            v *= {"usecond": 1e-3, "us": 1e-3, "nsecond": 1e-6, "ns": 1e-6, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3}.get(u, 1.0)
            a = agg.setdefault(d["Kernel Name"], [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("kernel,launches,total_ms,avg_ms,share_pct\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{v[0]},{v[1]:.4f},{v[1]/v[0]:.4f},{100*v[1]/tot:.2f}\n")
    return tot


def report(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for r in rows[2:]:
            f.write(f"## launch {r[hdr.index('ID')]}: {r[hdr.index('Kernel Name')]}\n")
            for k in KEYS:
                for i, h in enumerate(hdr):
                    if h == k or h.endswith("." + k):
                        f.write(f"{h},{units[i]},{r[i]}\n")
            f.write("\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs("profiles", exist_ok=True)
    tot = launches(sys.argv[2], f"profiles/{tag}_launches.csv")
    print("total kernel ms in launch list:", tot)
    for rep in sys.argv[3:]:
        name = os.path.basename(rep).replace(".ncu-rep", "")
        report(rep, f"profiles/{tag}_{name}.txt")
