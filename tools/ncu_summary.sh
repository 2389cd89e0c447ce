#!/bin/bash
# tools/ncu_summary.sh report.ncu-rep "title" > profiles/xxx_summary.txt : the numbers DESIGN.md / profiles/README.md quote
rep="$1"; title="${2:-}"
echo "# $title"
echo "# source: $(basename "$rep") (ncu --set full --clock-control none --import-source on; one launch)"
ncu -i "$rep" --page details 2>/dev/null | grep -E "^  [a-zA-Z_:]+.*\(|^    (Duration|DRAM Throughput|Memory Throughput|L1/TEX Cache Throughput|L2 Cache Throughput|Compute \(SM\) Throughput|Issue Slots Busy|Issued Ipc Active|Mem Busy|L1/TEX Hit Rate|L2 Hit Rate|Active Warps Per Scheduler|Eligible Warps Per Scheduler|Warp Cycles Per Issued Instruction|Registers Per Thread|Theoretical Active Warps per SM|Achieved Active Warps Per SM|Block Size|Grid Size|Local Memory|Dynamic Shared|Static Shared)" | sed 's/  */ /g'
ncu -i "$rep" --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]; v=rows[2]
want=['dram__bytes_read.sum','dram__bytes_write.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__inst_executed.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum','lts__t_sectors.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_no_instruction_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct']
for i,name in enumerate(h):
    if name in want: print(' ',name, v[i], u[i])
"
echo "# per source line (top 16): share of warp instructions, active threads per instruction, share of stall samples"
python3 "$(dirname "$0")/ncu_lines.py" "$rep" 16
