#!/bin/bash
# Rebuilds the library with different compile-time knobs ON THE GPU BOX and prints the
# average icp_iteration_kernel launch time for each (bench.py --skip-tsdf --skip-cpu).
# Knobs (all -D macros of open3d_b200/csrc): ICP_THIN_FACTOR (8), ICP_CELL_SCALE (0.5), ICP_FLUSH_EVERY (32),
# ICP_MIN_BLOCKS (3), ICP_TWO_PASS (1), ICP_DEFAULT_VARIANT (1 | 2 = TMA tiles), ICP_PREFETCH_SRC (0; round-2 candidate),
# ODO_BLOCKS_PER_SM (4; odometry iteration kernel).  Example:
#   bash profiles/tune_icp.sh "-DICP_PREFETCH_SRC=1" "-DICP_THIN_FACTOR=32" "-DICP_PREFETCH_SRC=1 -DICP_THIN_FACTOR=32"
cd "$(dirname "$0")/.."
for v in "$@"; do
  O3DB_NVCC_EXTRA="$v" bash open3d_b200/csrc/build.sh > /dev/null 2>&1 || { echo "BUILD FAILED: $v"; continue; }
  timeout 150 python bench.py --steps 3 --warmup 3 --skip-tsdf --skip-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('VARIANT', '''$v''', 'iter_us', round(d['roofline']['avg_launch_us'],1), 'value', round(d['value'],1), 'frac', round(d['roofline']['frac'],4), 'fitness', d['result']['fitness'], 'err', d['result']['transformation_error_vs_ground_truth'])
"
done
