#!/bin/bash
# round 2, GPU call 1: seeded ICP search — parity, per-iteration times, knob sweep, ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_icp_gpu.py -m gpu -x -q > gpurun_out/r02_t1.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_t1.log
timeout 300 python profiles/icp_iter_times.py 30 3 > gpurun_out/r02_iter_seeded.log 2>&1; cat gpurun_out/r02_iter_seeded.log
bash profiles/tune_icp.sh "-DICP_SEEDED=0" "-DICP_MIN_BLOCKS=4" "-DICP_MIN_BLOCKS=5" "-DICP_MIN_BLOCKS=2" "-DICP_CELL_SCALE=0.34" "-DICP_CELL_SCALE=0.25" "-DICP_THIN_FACTOR=32" "" 2>&1 | tee gpurun_out/r02_tune1.log
ICP_ITERS=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_iteration_kernel -s 10 -c 1 \
    -o gpurun_out/r02_icp_seeded python profiles/profile_workload.py icp > gpurun_out/r02_ncu1.log 2>&1; tail -2 gpurun_out/r02_ncu1.log
