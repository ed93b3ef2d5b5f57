#!/bin/bash
# ncu launch list (time + DRAM bytes per launch) of ONE single-stream pass over a 32-element chunk of the headline workload
# usage: tools/launch_list.sh <tag>   -> gpurun_out/<tag>_launches_b32.csv
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/$1_launches_b32.csv \
    python bench.py --steps 1 --warmup 3 --elements 32 --no-cpu --no-e2e --no-configs > gpurun_out/$1_launches.log 2>&1
tail -1 gpurun_out/$1_launches.log | cut -c1-160
