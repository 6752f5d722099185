#!/usr/bin/env python
"""Summarise ncu output for profiles/: a launch list (--metrics gpu__time_duration.sum CSV) into per-kernel
shares, and/or a full capture (.ncu-rep) into the handful of metrics DESIGN.md / bench.py quote.

    python tools/ncu_summary.py --launches gpurun_out/launches.csv --rep gpurun_out/prof.ncu-rep > profiles/xyz.txt
"""
import argparse
import collections
import csv
import io
import subprocess

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        a = agg.setdefault(row["Kernel Name"][:96], [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values()) or 1.0
    print("# launch list %s (cold-cache, serialised: compare SHARES, not absolutes)" % path)
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        print("%-98s n=%3d total=%9.3f ms avg=%8.4f ms share=%5.1f%%" % (k, n, ms, ms / n, 100 * ms / tot))


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("# full capture %s" % path)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("## kernel:", d.get("Kernel Name", "?")[:110])
        for k in KEYS:
            if k in d:
                print("%-84s %-14s %s" % (k, units[hdr.index(k)], d[k]))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    a = ap.parse_args()
    if a.launches:
        launches(a.launches)
    if a.rep:
        rep(a.rep)
