#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_icp_gpu.py -m gpu -q 2>&1 | tail -8
for v in 2 1; do echo "== variant $v"; ICP_VARIANT=$v timeout 300 python profiles/icp_iter_times.py 30 3 2>&1 | head -2; done | tee gpurun_out/r02_iter10.log
ICP_ITERS=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_iteration_kernel -s 10 -c 1 \
    -o gpurun_out/r02_icp_staged10 python profiles/profile_workload.py icp > gpurun_out/r02_ncu10.log 2>&1; tail -2 gpurun_out/r02_ncu10.log
bash profiles/tune_icp.sh "-DICP_MIN_BLOCKS=4" "-DICP_MIN_BLOCKS=5" "-DICP_TRANSPOSE_SMEM=1" "-DICP_CERTIFY=0" "" 2>&1 | tee gpurun_out/r02_tune10.log
