#!/bin/bash
# default build (768-thread ICP block, e2e copy overlap, probe hardening): full GPU suite + both bench lines + per-iteration times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest24.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench24_icp.json 2> gpurun_out/r02_bench24.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench24_icp.json').read().strip().splitlines()[-1])
print('icp', d['value'], 'us', d['roofline']['avg_launch_us'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], 'build ms', d['index_build_ms'])
t=d['tsdf']; print('tsdf', t['depth_only']['value'], t['depth_color']['value'], 'cpu', t['cpu_baseline'])"
timeout 300 python profiles/icp_iter_times.py 30 3 2>&1 | head -2 | tee gpurun_out/r02_iter24.log
timeout 300 python profiles/slam_time.py 100 2>&1 | tail -1 | tee gpurun_out/r02_slam24.log
