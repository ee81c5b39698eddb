#!/bin/bash
# ncu evidence for the SLAM loop with the level-resident odometry kernel: warm launch list + one full capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KN='regex:odometry|pyramid_level|clip_transform_pair|ray_cast|range_|integrate16|touch_kernel'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k "$KN" -s 96 -c 72 --csv \
  --log-file gpurun_out/r02_launches_slam35.csv python profiles/slam_time.py 12 > gpurun_out/r02_ncu35a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:odometry_level_kernel -s 16 -c 2 \
  -o gpurun_out/r02_level_kernel -f python profiles/slam_time.py 12 > gpurun_out/r02_ncu35b.log 2>&1
tail -3 gpurun_out/r02_ncu35a.log gpurun_out/r02_ncu35b.log
ls -la gpurun_out/r02_level_kernel.ncu-rep
