#!/bin/bash
# fused pyramid + PDL odometry + division-free ray march: tests, bench (slam), launch list of the SLAM loop, warm-cache ICP capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_odometry_gpu.py tests/test_tsdf_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/r02_pytest20.log
cat gpurun_out/slam_100_frames_vs_oracle.txt
timeout 900 python bench.py --steps 3 --warmup 3 --metric tsdf --skip-cpu > gpurun_out/r02_bench20_tsdf.json 2> gpurun_out/r02_bench20.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench20_tsdf.json').read().strip().splitlines()[-1])
print('tsdf fps', d['value'], 'colour fps', d['depth_color']['value'], 'raycast', d['raycast'], 'slam', d['dense_slam'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:o3db -c 300 --csv --log-file gpurun_out/r02_launches20_slam.csv python profiles/profile_workload.py slam > gpurun_out/r02_ll20.log 2>&1; tail -2 gpurun_out/r02_ll20.log
ICP_ITERS=24 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:icp_iteration -s 20 -c 1 \
    -o gpurun_out/r02_icp20_warm python profiles/profile_workload.py icp > gpurun_out/r02_ncu20.log 2>&1; tail -2 gpurun_out/r02_ncu20.log
