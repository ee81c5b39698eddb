#!/bin/bash
# 1 GPU: ICP compile-time variants (transposed reduction cadence / smem transpose / occupancy), then the default build: tests + TSDF host-only probe + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash profiles/tune_icp.sh "-DICP_DEFER=4 -DICP_MIN_BLOCKS=2" "-DICP_DEFER=2 -DICP_MIN_BLOCKS=2" "-DICP_DEFER=8 -DICP_MIN_BLOCKS=2" "-DICP_DEFER=4" "-DICP_TRANSPOSE_SMEM=1" "-DICP_MIN_BLOCKS=2" "" 2>&1 | tee gpurun_out/r02_tune19.log
timeout 900 python -m pytest tests/test_tsdf_gpu.py tests/test_forwarders_gpu.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest19.log
echo "== tsdf host only"; O3DB_TSDF_HOST_ONLY=1 timeout 600 python bench.py --steps 3 --warmup 3 --metric tsdf --skip-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('host-only fps', d['value'], 'e2e', d['e2e']['value'])"
timeout 900 python bench.py --steps 5 --warmup 3 --metric tsdf --skip-cpu > gpurun_out/r02_bench19_tsdf.json 2> gpurun_out/r02_bench19.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench19_tsdf.json').read().strip().splitlines()[-1])
print('tsdf fps', d['value'], 'e2e', d['e2e']['value'], 'integrate us', d['roofline']['avg_launch_us'], 'touch us', d['roofline']['touch_kernel_avg_us'], 'colour fps', d['depth_color']['value'], 'slam', d['dense_slam'].get('frames_per_sec'))"
