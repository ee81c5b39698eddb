#!/bin/bash
# Rebuilds the library with different odometry knobs ON THE GPU BOX and prints the dense-SLAM ms/frame for each.
cd "$(dirname "$0")/.."
for v in "$@"; do
  O3DB_NVCC_EXTRA="$v" bash open3d_b200/csrc/build.sh > /dev/null 2>&1 || { echo "BUILD FAILED: $v"; continue; }
  echo "VARIANT '$v' $(timeout 200 python profiles/slam_time.py 100 2>&1 | tail -1)"
done
