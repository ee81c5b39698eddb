#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for s in 200 400 600; do SLAM_START=$s timeout 200 python profiles/slam_time.py 100 2>&1 | tail -1; done | tee gpurun_out/r02_slam28.log
