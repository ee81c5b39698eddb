#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_icp_gpu.py -m gpu -x -q 2>&1 | tail -15
for v in 2 1; do echo "== variant $v"; ICP_VARIANT=$v timeout 300 python profiles/icp_iter_times.py 30 3 2>&1 | head -2; done | tee gpurun_out/r02_iter9.log
echo "== cell 0.35"; CELL_SCALE=0.35 timeout 300 python profiles/icp_iter_times.py 30 3 2>&1 | head -2 | tee -a gpurun_out/r02_iter9.log
ICP_ITERS=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_iteration_kernel -s 10 -c 1 \
    -o gpurun_out/r02_icp_staged9 python profiles/profile_workload.py icp > gpurun_out/r02_ncu9.log 2>&1; tail -2 gpurun_out/r02_ncu9.log
bash profiles/tune_icp.sh "-DICP_MIN_BLOCKS=4" "-DICP_MIN_BLOCKS=5" "-DICP_TRANSPOSE_SMEM=1" "-DICP_MIN_BLOCKS=4 -DICP_CELL_SCALE=0.35" "" 2>&1 | tee gpurun_out/r02_tune9.log
