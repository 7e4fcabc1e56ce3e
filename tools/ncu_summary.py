#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics per captured launch and the
executed-instruction mix by opcode.  Usage: python tools/ncu_summary.py file.ncu-rep [--sass]"""
import collections
import csv
import io
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__grid_size', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("----")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"{w:90s} {r[i]} {units[i]}")
    if "--sass" in sys.argv:
        rows = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass"]))))
        hdr, data, name = None, [], None

        def flush():
            if not data:
                return
            ia, ie, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
            tot = sum(int(r[ie]) for r in data)
            ops, samp = collections.Counter(), collections.Counter()
            for r in data:
                parts = r[ia].split()
                op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
                ops[op] += int(r[ie])
                samp[op] += int(r[isamp])
            print("====", name, "total warp-inst", tot)
            for op, c in ops.most_common(18):
                print(f"  {op:10s} {c:12d} {100 * c / max(tot, 1):5.1f}%  samples {samp[op]}")
        for r in rows:
            if r and r[0] == 'Kernel Name':
                flush()
                name, hdr, data = r[1], None, []
            elif r and r[0] == 'Address':
                hdr = r
            elif hdr:
                data.append(r)
        flush()


if __name__ == "__main__":
    main()
