#!/bin/bash
# device-timer exec stats in integrate16 + bench.py: TSDF tests, both bench lines (final artefacts)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tsdf_gpu.py tests/test_forwarders_gpu.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest32.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench32_icp.json 2> gpurun_out/r02_bench32.err
timeout 900 python bench.py --steps 10 --warmup 3 --metric tsdf > gpurun_out/r02_bench32_tsdf.json 2>> gpurun_out/r02_bench32.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench32_icp.json').read().strip().splitlines()[-1])
print('icp', round(d['value']), 'us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2))
t=json.loads(open('gpurun_out/r02_bench32_tsdf.json').read().strip().splitlines()[-1])
r=t['roofline']; c=t['depth_color']['roofline']
print('tsdf', round(t['value']), 'e2e', round(t['e2e']['value']), 'colour', round(t['depth_color']['value']), 'slam', t['dense_slam'].get('frames_per_sec'))
print(' depth: events us', r['avg_launch_us'], 'frac', r['frac'], 'device-timer us', r['avg_exec_us_device_timer'], 'frac', r['frac_device_timer'])
print(' colour: events us', c['avg_launch_us'], 'frac', c['frac'], 'device-timer us', c['avg_exec_us_device_timer'], 'frac', c['frac_device_timer'])
PY
tail -3 gpurun_out/r02_bench32.err
