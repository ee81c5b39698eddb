#!/bin/bash
# chunk-blocked working source (one bulk copy per chunk): ICP tests + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_icp_gpu.py tests/test_baseline_sizes_gpu.py tests/test_forwarders_gpu.py -m gpu -q -x -k "not tsdf and not slam" 2>&1 | tail -6 | tee gpurun_out/r02_pytest25.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-tsdf --skip-cpu > gpurun_out/r02_bench25_icp.json 2> gpurun_out/r02_bench25.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench25_icp.json').read().strip().splitlines()[-1])
print('icp', d['value'], 'us', d['roofline']['avg_launch_us'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
timeout 300 python profiles/icp_iter_times.py 30 3 2>&1 | head -2 | tee gpurun_out/r02_iter25.log
