#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O3DB_NVCC_EXTRA="-DICP_TIMING=1" bash open3d_b200/csrc/build.sh > /dev/null 2>&1
echo "== variant 2"; ICP_VARIANT=2 timeout 300 python profiles/icp_timing.py 2>&1 | tee gpurun_out/r02_timing6_v2.log
O3DB_NVCC_EXTRA="-DICP_TIMING=1 -DO3DB_FENCE_ACQREL=1" bash open3d_b200/csrc/build.sh > /dev/null 2>&1
echo "== variant 2, acq_rel fence"; ICP_VARIANT=2 timeout 300 python profiles/icp_timing.py 2>&1 | tee gpurun_out/r02_timing6_v2_acqrel.log
bash open3d_b200/csrc/build.sh > /dev/null 2>&1
timeout 900 python -m pytest tests/test_icp_gpu.py -m gpu -x -q 2>&1 | tail -25
