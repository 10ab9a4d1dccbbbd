#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.
usage: ncu_summary.py report.ncu-rep out.txt [--source]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
    "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_uniform.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append("== %s  grid %s block %s" % (name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append("  %-85s %s %s" % (k, r[i], units[i]))
    if "--source" in sys.argv:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                             capture_output=True, text=True).stdout
        lines.append("\n== top SASS lines by samples")
        rs = list(csv.reader(io.StringIO(src)))
        if rs:
            h = rs[0]
            try:
                si = h.index("# Samples") if "# Samples" in h else [i for i, x in enumerate(h) if "Samples" in x][0]
                body = [r for r in rs[1:] if len(r) > si and r[si].replace(".", "").isdigit()]
                body.sort(key=lambda r: -float(r[si]))
                for r in body[:40]:
                    lines.append("  %8s  %s" % (r[si], " | ".join(r[1:3])[:150]))
            except Exception as e:
                lines.append("  (source page not parsed: %s)" % e)
    if "--traffic-json" in sys.argv:
        import json
        jpath = sys.argv[sys.argv.index("--traffic-json") + 1]
        models = int(sys.argv[sys.argv.index("--models") + 1])
        r = rows[2]

        def num(k):
            v, u = float(r[hdr.index(k)]), units[hdr.index(k)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        json.dump({"kernel": r[hdr.index("Kernel Name")][:80], "models_per_launch": models,
                   "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                   "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
                   "source": rep.split("/")[-1] + " (ncu --set full --clock-control none)"},
                  open(jpath, "w"), indent=1)
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:70]))


if __name__ == "__main__":
    main()
