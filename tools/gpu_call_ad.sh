#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_warp -c 1 -f -o gpurun_out/prof_knn8_r02 python tools/decode_profile.py 250000 200 65 > gpurun_out/ad_ncu.log 2>&1
tail -2 gpurun_out/ad_ncu.log
python tools/ncu_summary.py gpurun_out/prof_knn8_r02.ncu-rep > gpurun_out/prof_knn8_r02_summary.csv
ncu -i gpurun_out/prof_knn8_r02.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
for k in ('smsp__inst_executed.sum','launch__registers_per_thread','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','gpu__time_duration.sum','launch__grid_size','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'):
    if k in h: print('%-90s %s'%(k, r[h.index(k)]))
"
