#!/bin/bash
# real-form pipeline (default): launch list with DRAM traffic (one 32-element chunk) + full ncu captures of the real GEMM
# (argument "all": also the hexahedron integration kernel and the real potrf tile kernel)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r01_rs_launches_traffic_b32.csv \
    python bench.py --steps 1 --warmup 3 --elements 32 --no-cpu --no-e2e > gpurun_out/ncu_rs_launch.log 2>&1
tail -1 gpurun_out/ncu_rs_launch.log | cut -c1-120
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,launch__waves_per_multiprocessor,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active
run() { # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/prof_$1 -f \
      python bench.py --steps 1 --warmup 3 --elements 32 --no-cpu --no-e2e > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv --metrics $M > gpurun_out/r01_$1_ncu_full_summary.csv 2>&1
  head -c 600 gpurun_out/r01_$1_ncu_full_summary.csv; echo
}
run gemm_real gemm_nc 864 12
if [ "$1" = "all" ]; then
  run tp3_rs tp3_kernel 8 1
  run potrf_real potrf_inv 330 3
fi
