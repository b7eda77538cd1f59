#!/usr/bin/env python
"""Turn what gpurun brought back (gpurun_out/) into the tracked summaries under profiles/.

    python profiles/summarize.py launches <launches.csv> <out.md>      # per-kernel totals of an ncu launch list
    python profiles/summarize.py full <report.ncu-rep> <out.md>         # key metrics of every kernel in a --set full report
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")
    return name.replace("(bool)", "")


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3}
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        k = short(r[iK])
        tot[k] += float(r[iV].replace(",", "")) * scale.get(r[iU], 1.0)
        cnt[k] += 1
    T = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"ncu launch list `{src}`: {sum(cnt.values())} launches, {T:.3f} ms of kernel time "
                "(cold-cache, serialised: compare shares, not absolutes)\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in tot.most_common():
            f.write(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / T:.2f}% |\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"Key metrics of `{src}` (ncu --set full --clock-control none)\n\n")
        for r in rows[2:]:
            f.write(f"### `{short(r[hdr.index('Kernel Name')])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
