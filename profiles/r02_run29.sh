#!/bin/bash
# round-end records on one B200: full GPU suite, both bench lines, the reference arm, ncu captures of the final kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest29.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench29_icp.json 2> gpurun_out/r02_bench29.err
timeout 900 python bench.py --steps 10 --warmup 3 --metric tsdf > gpurun_out/r02_bench29_tsdf.json 2>> gpurun_out/r02_bench29.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench29_reference.json 2>> gpurun_out/r02_bench29.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench29_icp.json').read().strip().splitlines()[-1])
print('icp', round(d['value']), 'us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'cpu', d['cpu_baseline']['value'])
t=json.loads(open('gpurun_out/r02_bench29_tsdf.json').read().strip().splitlines()[-1])
print('tsdf', round(t['value']), 'e2e', round(t['e2e']['value']), 'colour', round(t['depth_color']['value']), 'slam', t['dense_slam'].get('frames_per_sec'), 'cpu', t['cpu_baseline']['value'])
print(open('gpurun_out/r02_bench29_reference.json').read().strip().splitlines()[-1][:400])
PY
ICP_ITERS=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_iteration -s 10 -c 1 -o gpurun_out/r02_icp29 python profiles/profile_workload.py icp > gpurun_out/r02_ncu29a.log 2>&1; tail -1 gpurun_out/r02_ncu29a.log
TSDF_COLOR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"integrate16_kernel|touch_kernel" -s 60 -c 2 -o gpurun_out/r02_tsdf_depth29 python profiles/profile_workload.py tsdf > gpurun_out/r02_ncu29b.log 2>&1; tail -1 gpurun_out/r02_ncu29b.log
TSDF_COLOR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"integrate16_kernel" -s 30 -c 1 -o gpurun_out/r02_tsdf_color29 python profiles/profile_workload.py tsdf > gpurun_out/r02_ncu29c.log 2>&1; tail -1 gpurun_out/r02_ncu29c.log
