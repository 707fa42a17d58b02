#!/bin/bash
# usage: ncu_metrics.sh file.ncu-rep  -> key metrics of the first profiled kernel
ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin))
hdr=r[0]; vals=r[2] if len(r)>2 else r[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors.sum','lts__t_sectors_op_red.sum','lts__t_sectors_op_atom.sum','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']
for w in want:
    if w in hdr:
        i=hdr.index(w); print(f'{w:85s} {r[1][i]:>10s} {vals[i]}')
"
