#!/bin/bash
# round 2, GPU call 4: transposed warp reduction (1 accumulator register per lane), flat slow path, staged vs direct
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_icp_gpu.py -m gpu -x -q > gpurun_out/r02_t4.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_t4.log
for v in 2 1; do
  echo "== variant $v"
  ICP_VARIANT=$v timeout 300 python profiles/icp_iter_times.py 30 3 2>&1 | tee gpurun_out/r02_iter4_v$v.log | head -2
done
bash profiles/tune_icp.sh "-DICP_MIN_BLOCKS=4" "-DICP_MIN_BLOCKS=5" "-DICP_MIN_BLOCKS=4 -DICP_DEFAULT_VARIANT=1" "-DICP_MIN_BLOCKS=5 -DICP_DEFAULT_VARIANT=1" "-DICP_DEFAULT_VARIANT=1" "-DICP_THIN_FACTOR=32" "" 2>&1 | tee gpurun_out/r02_tune4.log
ICP_ITERS=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_iteration_kernel -s 10 -c 1 \
    -o gpurun_out/r02_icp_staged4 python profiles/profile_workload.py icp > gpurun_out/r02_ncu4.log 2>&1; tail -2 gpurun_out/r02_ncu4.log
