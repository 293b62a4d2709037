#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py rep.ncu-rep [out.txt] [title]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U, V = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__pcsamp_warps_issue_stalled', 'smsp__pcsamp_sample_count']
out = [sys.argv[3] if len(sys.argv) > 3 else rep]
for h, u, v in zip(H, U, V):
    if any(h.startswith(k) for k in keep) and 'not_issued' not in h:
        out.append('%-84s %-16s %s' % (h, u, v))
txt = "\n".join(out) + "\n"
if len(sys.argv) > 2 and sys.argv[2] != '-':
    open(sys.argv[2], 'w').write(txt)
print(txt)
