#!/bin/bash
# round 2, GPU call 5: per-block timestamps of a steady-state iteration (ICP_TIMING build), both variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O3DB_NVCC_EXTRA="-DICP_TIMING=1" bash open3d_b200/csrc/build.sh > /dev/null 2>&1
for v in 2 1; do echo "== variant $v"; ICP_VARIANT=$v timeout 300 python profiles/icp_timing.py 2>&1 | tee gpurun_out/r02_timing_v$v.log; done
O3DB_NVCC_EXTRA="-DICP_TIMING=1 -DICP_PDL=0" bash open3d_b200/csrc/build.sh > /dev/null 2>&1
echo "== variant 2, PDL off"; ICP_VARIANT=2 timeout 300 python profiles/icp_timing.py 2>&1 | tee gpurun_out/r02_timing_v2_nopdl.log
bash open3d_b200/csrc/build.sh > /dev/null 2>&1
timeout 900 python -m pytest tests/test_icp_gpu.py -m gpu -x -q 2>&1 | tail -3
