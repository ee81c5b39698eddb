#!/bin/bash
# last GPU call of the round: the TSDF / SLAM bench line with the volume built before the SLAM clock starts
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 80 python bench.py --steps 10 --warmup 3 --metric tsdf --skip-cpu > gpurun_out/r02_bench37_tsdf.json 2> gpurun_out/r02_bench37.err
python - <<'PY'
import json
try:
    t=json.loads(open('gpurun_out/r02_bench37_tsdf.json').read().strip().splitlines()[-1])
    print('tsdf', round(t['value']), 'e2e', round(t['e2e']['value']), 'colour', round(t['depth_color']['value']), 'slam', t['dense_slam'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -n 3 gpurun_out/r02_bench37.err
