#!/bin/bash
# information matrix + full GPU suite; SLAM launch list; ICP per-phase timing (ICP_TIMING build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest21.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pyramid|odometry|clip_transform|integrate|range_|ray_cast|touch" -c 300 --csv --log-file gpurun_out/r02_launches21_slam.csv python profiles/profile_workload.py slam > gpurun_out/r02_ll21.log 2>&1; tail -1 gpurun_out/r02_ll21.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ray_cast -s 3 -c 1 -o gpurun_out/r02_raycast21 python profiles/profile_workload.py raycast > gpurun_out/r02_ncu21.log 2>&1; tail -1 gpurun_out/r02_ncu21.log
O3DB_NVCC_EXTRA="-DICP_TIMING=1" bash open3d_b200/csrc/build.sh > /dev/null 2>&1 && timeout 300 python profiles/icp_timing.py 2>&1 | tail -24 | tee gpurun_out/r02_icp_timing21.log
