#!/bin/bash
# Round-2 evidence for profiles/: launch list with DRAM traffic (one single-stream pass, 32-element chunk), full ncu captures of
# the real GEMM (panel update / solve / HERK launches), the hexahedron integration kernel and the diagonal-tile kernel.
# Run on the GPU box: tools/profile_r02.sh ; results land in gpurun_out/ (copy the summaries to profiles/).
tools/launch_list.sh r02
python tools/launch_summary.py gpurun_out/r02_launches_b32.csv > gpurun_out/r02_launches_b32_summary.csv
python tools/launch_summary.py gpurun_out/r02_launches_b32.csv grid > gpurun_out/r02_launches_b32_by_grid.csv
tools/ncu_full.sh r02_gemm_real gemm_nc 872 12 > /dev/null
tools/ncu_full.sh r02_tp3 tp3_kernel 8 1 > /dev/null
tools/ncu_full.sh r02_potrf_real potrf_inv 330 2 > /dev/null
ls -la gpurun_out | tail -12
