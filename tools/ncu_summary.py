#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box) into a text table under profiles/ and, for the headline decode kernel,
into profiles/ncu_latest.json -- the one place bench.py reads profiler counters from (nothing is pasted into bench.py).

    python tools/ncu_summary.py rep.ncu-rep out.txt "title" [--kernel REGEX] [--frames F] [--json profiles/ncu_latest.json]

--kernel picks the first launch whose name matches (default: the first launch of the report); --frames is the number of
frames that launch decoded (for the per-frame DRAM figure)."""
import argparse
import csv
import json
import os
import re
import subprocess

KEEP = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__block_size', 'launch__grid_size',
        'launch__cluster', 'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__pcsamp_warps_issue_stalled', 'smsp__pcsamp_sample_count']


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out", nargs="?", default="-")
    ap.add_argument("title", nargs="?", default=None)
    ap.add_argument("--kernel", default=None)
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--json", default=None)
    ap.add_argument("--fixed10-into", default=None, help="merge this (fixed-10-iterations) capture's counters into an existing json")
    ap.add_argument("--mean-iters", type=float, default=0.0, help="mean decoder iterations per frame in that launch")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    name_col = H.index("Kernel Name")
    row = rows[2]
    if a.kernel:
        row = next(r for r in rows[2:] if re.search(a.kernel, r[name_col]))
    out = [a.title or a.rep, "kernel: " + row[name_col][:160]]
    vals = {}
    for h, u, v in zip(H, U, row):
        if any(h.startswith(k) for k in KEEP) and 'not_issued' not in h:
            out.append('%-84s %-16s %s' % (h, u, v))
            vals[h] = (v, u)
    txt = "\n".join(out) + "\n"
    if a.out != '-':
        open(a.out, 'w').write(txt)
    print(txt)
    if a.fixed10_into:
        num = lambda k: float(vals[k][0].replace(",", "")) if k in vals else None
        d = json.load(open(a.fixed10_into))
        rd, wr = to_bytes(*vals['dram__bytes_read.sum']), to_bytes(*vals['dram__bytes_write.sum'])
        d.update({"fixed10_summary": os.path.relpath(a.out) if a.out != '-' else None, "fixed10_frames_in_launch": a.frames,
                  "fixed10_dram_bytes_per_frame": (rd + wr) / a.frames if a.frames else None,
                  "fixed10_issue_active_pct": num('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                  "fixed10_l2_hit_rate_pct": num('lts__t_sector_hit_rate.pct'),
                  "fixed10_warp_inst_per_frame_iteration": (num('smsp__inst_executed.sum') / a.frames / a.mean_iters
                                                            if a.frames and a.mean_iters else None)})
        json.dump(d, open(a.fixed10_into, "w"), indent=1)
        print("merged into", a.fixed10_into)
    if a.json:
        rd, wr = to_bytes(*vals['dram__bytes_read.sum']), to_bytes(*vals['dram__bytes_write.sum'])
        try:
            commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
        except Exception:
            commit = None
        num = lambda k: float(vals[k][0].replace(",", "")) if k in vals else None
        d = {"kernel": row[name_col][:160], "summary": os.path.relpath(a.out) if a.out != '-' else None, "report": os.path.basename(a.rep),
             "read_at_commit": commit, "frames_in_launch": a.frames, "dram_bytes_read": rd, "dram_bytes_write": wr,
             "dram_bytes_per_frame": (rd + wr) / a.frames if a.frames else None,
             "gpu_time_ms": num('gpu__time_duration.sum'),
             "issue_active_pct": num('smsp__issue_active.avg.pct_of_peak_sustained_active'),
             "warps_active_pct": num('sm__warps_active.avg.pct_of_peak_sustained_active'),
             "alu_pipe_pct": num('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
             "fma_fp16_pipe_pct": num('sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active'),
             "registers_per_thread": num('launch__registers_per_thread'),
             "l2_hit_rate_pct": num('lts__t_sector_hit_rate.pct'),
             "warp_inst_executed": num('smsp__inst_executed.sum')}
        if a.frames and a.mean_iters and d["warp_inst_executed"]:
            d["mean_iters_in_launch"] = a.mean_iters
            d["warp_inst_per_frame_iteration"] = d["warp_inst_executed"] / a.frames / a.mean_iters
        if vals.get('gpu__time_duration.sum', ("", ""))[1] == "us" and d["gpu_time_ms"]:
            d["gpu_time_ms"] /= 1e3
        json.dump(d, open(a.json, "w"), indent=1)
        print("wrote", a.json)


if __name__ == "__main__":
    main()
