#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md / profiles/ quote.
    python tools/ncu_summary.py gpurun_out/prof_agents.ncu-rep [more.ncu-rep ...]
"""
import csv
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors.sum', 'lts__t_sectors_op_red.sum',
    'lts__t_sectors_op_atom.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum',
    'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fmaheavy.sum',
    'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max',
    'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        print(f'=== {rep}')
        for r in rows[2:]:
            print(f"--- launch {r[hdr.index('ID')]}: {r[hdr.index('Kernel Name')][:70]}")
            for w in WANT:
                if w in hdr:
                    print(f'  {w:88s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}')


if __name__ == '__main__':
    main()
