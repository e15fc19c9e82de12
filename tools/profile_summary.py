#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/profile_summary.py <tag> gpurun_out/launches.csv gpurun_out/prof.ncu-rep [kernel-label]

writes profiles/<tag>_launches.md (per-kernel share of the step, from the gpu__time_duration pass) and
profiles/<tag>_<kernel>.md (key metrics of the --set full capture + SASS evidence)."""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        agg.setdefault(r[ki].split("(")[0], []).append(v * scale)
    return agg


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def sass_histogram(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return None, 0, 0
    name = rows[0][1] if len(rows[0]) > 1 else ""
    hdr = rows[1]
    isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
    hist = collections.Counter()
    tot = 0
    for r in rows[2:]:
        if not r[iex].isdigit():
            continue
        t = r[isrc].split()
        op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
        hist[op.split(".")[0]] += int(r[iex])
        tot += int(r[iex])
    return (name, hist, tot), len(rows) - 2, tot


def main():
    tag, lpath, rep = sys.argv[1], sys.argv[2], sys.argv[3]
    label = sys.argv[4] if len(sys.argv) > 4 else "kernel"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if os.path.exists(lpath):
        agg = launches(lpath)
        tot = sum(sum(v) for v in agg.values())
        with open(os.path.join(ROOT, "profiles", tag + "_launches.md"), "w") as fh:
            fh.write("# %s: per-kernel device time (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n" % tag)
            fh.write("Cold-cache, serialised launches: compare SHARES, not absolutes.\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
                fh.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k, len(v), sum(v), 100 * sum(v) / tot))
    if os.path.exists(rep):
        m = raw_metrics(rep)
        (name, hist, tot), ninstr, _ = sass_histogram(rep)
        with open(os.path.join(ROOT, "profiles", "%s_%s.md" % (tag, label)), "w") as fh:
            fh.write("# %s: ncu --set full --clock-control none capture of `%s`\n\n" % (tag, label))
            fh.write("kernel: `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % name)
            for w in WANT:
                if w in m:
                    fh.write("| %s | %s | %s |\n" % (w, m[w][0], m[w][1]))
            fh.write("\nSASS: %d instructions in the kernel; executed warp-instructions by opcode:\n\n| opcode | share |\n|---|---:|\n" % ninstr)
            for k, v in hist.most_common(16):
                fh.write("| %s | %.2f%% |\n" % (k, 100.0 * v / tot))
    print("ok")


if __name__ == "__main__":
    main()
