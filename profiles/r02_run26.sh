#!/bin/bash
# dense-SLAM segments as bench.py assigns them to ranks, on ONE GPU: is the N=8 slowdown data-dependent?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for s in 0 100 300 500 700; do SLAM_START=$s timeout 200 python profiles/slam_time.py 100 2>&1 | tail -1; done | tee gpurun_out/r02_slam26.log
