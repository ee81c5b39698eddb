#!/bin/bash
# status ring (no in-stream copies), dynamic units, rcp; fat-block ICP variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r02_pytest22.log
timeout 900 python bench.py --steps 5 --warmup 3 --metric tsdf --skip-cpu > gpurun_out/r02_bench22_tsdf.json 2> gpurun_out/r02_bench22.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench22_tsdf.json').read().strip().splitlines()[-1])
print('tsdf fps', d['value'], 'e2e', d['e2e']['value'], 'integrate us', d['roofline']['avg_launch_us'], 'touch us', d['roofline']['touch_kernel_avg_us'], 'colour fps', d['depth_color']['value'], 'raycast ms', d['raycast']['ms_per_frame'], 'slam', d['dense_slam'].get('frames_per_sec'), 'icp', d['icp']['value'])"
bash profiles/tune_icp.sh "-DICP_STAGED_THREADS=768 -DICP_MIN_BLOCKS=1" "-DICP_STAGED_THREADS=512 -DICP_MIN_BLOCKS=1" "-DICP_STAGED_THREADS=768 -DICP_MIN_BLOCKS=1 -DICP_DEFER=4" "" 2>&1 | tee gpurun_out/r02_tune22.log
