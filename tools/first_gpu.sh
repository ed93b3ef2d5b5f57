#!/bin/bash
# first end-to-end GPU session: smoke, tests, bench, launch list, one full ncu capture of the GEMM kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --elements 128 --e2e-elements 64 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err
cat gpurun_out/bench_first.json; tail -5 gpurun_out/bench_first.err
