#!/bin/bash
# level-resident odometry kernel (thread-block cluster), device-side state init, zero-copy result; VoxelBlockGrid auto-grow
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export O3DB_ODO_VERBOSE=1
timeout 600 python -m pytest tests/test_odometry_gpu.py tests/test_tsdf_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02_pytest33a.log
{
echo "## default (cluster 16 if available, zero-copy result)"; timeout 300 python profiles/slam_time.py 100 2>&1 | tail -2
echo "## O3DB_ODO_LEVEL_CLUSTER=8"; O3DB_ODO_LEVEL_CLUSTER=8 timeout 300 python profiles/slam_time.py 100 2>&1 | tail -2
echo "## cluster 16, level 2 only (O3DB_ODO_LEVEL_PX_PER_THREAD=3)"; O3DB_ODO_LEVEL_PX_PER_THREAD=3 timeout 300 python profiles/slam_time.py 100 2>&1 | tail -2
echo "## O3DB_ODO_LEVEL_CLUSTER=0 (per-iteration kernels, zero-copy result)"; O3DB_ODO_LEVEL_CLUSTER=0 timeout 300 python profiles/slam_time.py 100 2>&1 | tail -2
echo "## O3DB_ODO_LEVEL_CLUSTER=0 O3DB_ODO_NO_ZERO_COPY=1"; O3DB_ODO_LEVEL_CLUSTER=0 O3DB_ODO_NO_ZERO_COPY=1 timeout 300 python profiles/slam_time.py 100 2>&1 | tail -2
} | tee gpurun_out/r02_slam33.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest33b.log
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/r02_bench33_icp.json 2> gpurun_out/r02_bench33.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench33_icp.json').read().strip().splitlines()[-1])
    print('icp', round(d['value']), 'us', round(d['roofline']['avg_launch_us'],1), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'slam', d.get('tsdf',{}).get('dense_slam'))
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 gpurun_out/r02_bench33.err
