#!/bin/bash
# one `ncu --set full` capture (with source) of kernels matching a regex in the headline workload, 32-element chunk
# usage: tools/ncu_full.sh <tag> <kernel regex> <skip> <count>  -> gpurun_out/<tag>.ncu-rep + <tag>_summary.csv + <tag>_source.csv
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,launch__waves_per_multiprocessor,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/$1 -f \
    python bench.py --steps 1 --warmup 3 --elements 32 --no-cpu --no-e2e --no-configs > gpurun_out/$1.log 2>&1
ncu -i gpurun_out/$1.ncu-rep --page raw --csv --metrics $M > gpurun_out/$1_summary.csv 2>&1
ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>&1
head -c 400 gpurun_out/$1_summary.csv; echo
