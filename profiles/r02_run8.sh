#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_icp_gpu.py -m gpu -x -q 2>&1 | tail -15
for v in 2 1; do echo "== variant $v"; ICP_VARIANT=$v timeout 300 python profiles/icp_iter_times.py 30 3 2>&1 | head -2; done | tee gpurun_out/r02_iter8.log
CELL_SCALE=0.35 timeout 300 python profiles/icp_iter_times.py 30 3 2>&1 | head -2 | tee -a gpurun_out/r02_iter8.log
ICP_ITERS=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_iteration_kernel -s 10 -c 1 \
    -o gpurun_out/r02_icp_staged8 python profiles/profile_workload.py icp > gpurun_out/r02_ncu8.log 2>&1; tail -2 gpurun_out/r02_ncu8.log
O3DB_NVCC_EXTRA="-DICP_TIMING=1" bash open3d_b200/csrc/build.sh > /dev/null 2>&1
echo "== timing variant 2"; ICP_VARIANT=2 timeout 300 python profiles/icp_timing.py 2>&1 | tail -10 | tee gpurun_out/r02_timing8_v2.log
bash open3d_b200/csrc/build.sh > /dev/null 2>&1
