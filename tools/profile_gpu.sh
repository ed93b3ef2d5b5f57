#!/bin/bash
# launch list of one bench step + one full ncu capture of the dominant kernel (1 GPU, small batch: ncu replays each kernel)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size" 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none -s 453 -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --elements 32 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:gemm_nc -s 245 -c 16 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 1 --warmup 3 --elements 32 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
