#!/usr/bin/env python
"""Print selected metrics of an .ncu-rep: ncu_keys.py report.ncu-rep [substr ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
subs = sys.argv[2:] or [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit",
    "sm__warps_active.avg.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
    "sm__pipe_fp64_cycles_active.avg.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "l1tex__data_pipe_lsu_wavefronts_mem_lgds",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate",
    "lts__t_sector_hit_rate", "lts__throughput.avg.pct", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
    "l1tex__lsu_writeback_active.avg.pct", "l1tex__lsuin_requests.avg.pct",
    "smsp__thread_inst_executed_per_inst_executed", "issue_stalled_long_scoreboard_per",
    "issue_stalled_short_scoreboard_per", "issue_stalled_math_pipe", "issue_stalled_wait_per",
    "issue_stalled_not_selected_per", "issue_stalled_mio_throttle_per", "issue_stalled_lg_throttle_per",
    "issue_stalled_barrier_per", "issue_stalled_branch", "issue_stalled_dispatch", "issue_stalled_no_inst",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fp64", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_uniform.sum", "l1tex__m_xbar2l1tex_read_bytes.sum "]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:90], r[h.index("Grid Size")], r[h.index("Block Size")])
    for i, k in enumerate(h):
        if any(s.strip() in k and (not s.endswith(" ") or k == s.strip()) for s in subs):
            print("  %-90s %s %s" % (k, r[i], u[i]))
