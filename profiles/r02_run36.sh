#!/bin/bash
# pyramid kernel: spatial filter weights once per block; level-resident odometry kernel: 3 pixels per thread in flight.
# Tests under the default and the alternative variants, SLAM loop timing per variant, full suite, the TSDF / SLAM bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_odometry_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest36a.log
O3DB_ODO_PYR_UNROLL=1 O3DB_ODO_LEVEL_BATCH=4 timeout 300 python -m pytest tests/test_odometry_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest36b.log
{
echo "## default (batch 3, tap rows rolled)"; timeout 200 python profiles/slam_time.py 100 2>&1 | tail -1
echo "## O3DB_ODO_PYR_UNROLL=1"; O3DB_ODO_PYR_UNROLL=1 timeout 200 python profiles/slam_time.py 100 2>&1 | tail -1
echo "## O3DB_ODO_LEVEL_BATCH=1"; O3DB_ODO_LEVEL_BATCH=1 timeout 200 python profiles/slam_time.py 100 2>&1 | tail -1
echo "## O3DB_ODO_LEVEL_BATCH=4"; O3DB_ODO_LEVEL_BATCH=4 timeout 200 python profiles/slam_time.py 100 2>&1 | tail -1
} | tee gpurun_out/r02_slam36.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest36c.log
timeout 600 python bench.py --steps 10 --warmup 3 --metric tsdf > gpurun_out/r02_bench36_tsdf.json 2> gpurun_out/r02_bench36.err
python - <<'PY'
import json
try:
    t=json.loads(open('gpurun_out/r02_bench36_tsdf.json').read().strip().splitlines()[-1])
    print('tsdf', round(t['value']), 'e2e', round(t['e2e']['value']), 'colour', round(t['depth_color']['value']), 'slam', t['dense_slam'].get('frames_per_sec'), t['dense_slam'].get('ms_per_frame'), t['dense_slam'].get('gpu_launches_per_frame'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -n 3 gpurun_out/r02_bench36.err
