#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_icp_gpu.py -m gpu -x -q 2>&1 | tail -15
O3DB_NVCC_EXTRA="-DICP_TIMING=1" bash open3d_b200/csrc/build.sh > /dev/null 2>&1
echo "== timing variant 2"; ICP_VARIANT=2 timeout 300 python profiles/icp_timing.py 2>&1 | tail -12 | tee gpurun_out/r02_timing7_v2.log
bash open3d_b200/csrc/build.sh > /dev/null 2>&1
for cs in 0.5 0.35 0.25; do
  timeout 200 python bench.py --steps 3 --warmup 3 --skip-tsdf --skip-cpu --cell-scale $cs 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('CELL_SCALE $cs', 'iter_us', round(d['roofline']['avg_launch_us'],1), 'value', round(d['value'],1), 'build_ms', round(d['index_build_ms'],2), 'e2e', round(d['e2e']['value'],1), 'fitness', d['result']['fitness'])"
  CELL_SCALE=$cs timeout 200 python profiles/icp_iter_times.py 30 2 2>&1 | head -2
done 2>&1 | tee gpurun_out/r02_cells7.log
bash profiles/tune_icp.sh "-DICP_THIN_MERGE_MAX=0" "-DICP_MIN_BLOCKS=4 -DICP_CELL_SCALE=0.25" "-DICP_MIN_BLOCKS=4" "-DICP_CELL_SCALE=0.25 -DICP_DEFAULT_VARIANT=1" "" 2>&1 | tee gpurun_out/r02_tune7.log
