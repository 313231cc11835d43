#!/usr/bin/env python
"""Summarise an .ncu-rep (run where ncu is installed; no GPU needed): key raw metrics, executed SASS by opcode and
stall reasons.  Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_global_op_ld.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__cycles_active.avg',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct', 'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct',
        'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct']


def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, capture_output=True, text=True).stdout


def main():
    rows = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:80s} {r[i]:>18s} {units[i]}")
    src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "sass"]))))
    kernel, hdr2 = None, None
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = 0

    def flush():
        nonlocal agg, tot
        if kernel and tot:
            print("== executed SASS by opcode:", kernel, "total warp instructions", tot)
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:18]:
                print(f"  {k:12s} {v[0]:>12d} {100 * v[0] / tot:5.1f}%  avg lanes {v[1] / max(v[0], 1):5.1f}  stall samples {v[2]}")
        agg, tot = collections.defaultdict(lambda: [0, 0, 0]), 0

    for r in src:
        if not r:
            continue
        if r[0] == "Kernel Name":
            flush()
            kernel = r[1]
            continue
        if r[0] == "Address":
            hdr2 = r
            continue
        if hdr2 is None or len(r) < len(hdr2) // 2:
            continue
        op = r[hdr2.index("Source")].split()
        if op and op[0].startswith("@"):
            op = op[1:]
        if not op:
            continue
        o = op[0].split(".")[0]
        e = int(r[hdr2.index("Instructions Executed")] or 0)
        agg[o][0] += e
        agg[o][1] += int(r[hdr2.index("Thread Instructions Executed")] or 0)
        agg[o][2] += int(r[hdr2.index("# Samples")] or 0)
        tot += e
    flush()


if __name__ == "__main__":
    main()
