#!/bin/bash
# new integrate16 kernel (TMA-staged depth tile, PDL, whole-frame drop rule): TSDF tests, full suite, bench, timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tsdf_gpu.py tests/test_odometry_gpu.py -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/r02_pytest14a.log
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest14b.log
cat gpurun_out/slam_100_frames_vs_oracle.txt
timeout 900 python bench.py --steps 5 --warmup 3 --metric tsdf > gpurun_out/r02_bench14_tsdf.json 2> gpurun_out/r02_bench14.err; tail -c 1800 gpurun_out/r02_bench14_tsdf.json
TSDF_COLOR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"integrate16_kernel|touch_kernel" -s 60 -c 2 \
    -o gpurun_out/r02_tsdf_depth14 python profiles/profile_workload.py tsdf > gpurun_out/r02_ncu14b.log 2>&1; tail -2 gpurun_out/r02_ncu14b.log
TSDF_COLOR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"integrate16_kernel" -s 30 -c 1 \
    -o gpurun_out/r02_tsdf_color14 python profiles/profile_workload.py tsdf > gpurun_out/r02_ncu14c.log 2>&1; tail -2 gpurun_out/r02_ncu14c.log
