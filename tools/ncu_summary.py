#!/usr/bin/env python
"""Condense an `ncu --set full` report into the handful of numbers DESIGN.md / profiles/ quote.  Usage: ncu_summary.py rep [rep..]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (hmma) cycles active % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active % of peak"),
    ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "tensor(hmma) inst %"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe (UTCHMMA issue) %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor hmma cycles active (avg/SM)"),
    ("sm__cycles_active.avg", "SM cycles active (avg)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts %"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall long_scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall wait"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "selected"),
    ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall math_pipe_throttle"),
]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    h, u = rd[0], rd[1]
    for row in rd[2:]:
        print(f"== {rep.split('/')[-1]}: {row[h.index('Kernel Name')][:90]}")
        for k, label in KEYS:
            cand = [i for i, n in enumerate(h) if n == k or n.endswith("." + k)]
            if cand:
                i = cand[0]
                print(f"   {label:38s} {row[i]:>16s} {u[i]}")
