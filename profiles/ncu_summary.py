#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel): key raw metrics, stall reasons, and warp-stall samples per CUDA source line.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines]"""
import collections
import csv
import io
import subprocess
import sys


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    n_lines = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    print("kernel:", m.get("Kernel Name"), "grid", m.get("Grid Size"), "block", m.get("Block Size"))
    keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__occupancy_limit_registers",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sectors_op_read.sum", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
    for k in keys:
        if k in m:
            print("  %-70s %s %s" % (k, m[k], u.get(k, "")))
    st = []
    for h, v in zip(hdr, vals):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                st.append((float(v), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("stall reasons (warps per issue-active cycle):", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--print-source", "cuda,sass"))))
    cur = None
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 8 or r[0] in ("Line No", "Function Name"):
            continue
        if r[0] != "" and r[2] == "-":
            try:
                agg[(cur, int(r[0]), r[1][:90])] = (int(r[6]), int(r[7]))
            except ValueError:
                pass
    tot = sum(v[0] for v in agg.values()) or 1
    byfile = collections.Counter()
    for (f, _, _), (s, _) in agg.items():
        byfile[f] += s
    print("samples by file:", {k: "%.1f%%" % (100.0 * v / tot) for k, v in byfile.most_common()})
    for (f, l, src), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n_lines]:
        print("  %5.2f%% %-14s %4d exec=%10d  %s" % (100.0 * s / tot, f, l, e, src))


if __name__ == "__main__":
    main()
