#!/bin/bash
# Python fast paths on the frame loop (no per-iteration log copy unless asked, cached extrinsic, cheap stream / pointer
# lookups) on top of the level-resident odometry kernel: full GPU suite, SLAM loop timing + launch list, both bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest34.log
timeout 300 python profiles/slam_time.py 100 2>&1 | tail -2 | tee gpurun_out/r02_slam34.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r02_launches_slam34.csv python profiles/slam_time.py 10 > gpurun_out/r02_ncu34.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench34_icp.json 2> gpurun_out/r02_bench34.err
timeout 900 python bench.py --steps 10 --warmup 3 --metric tsdf > gpurun_out/r02_bench34_tsdf.json 2>> gpurun_out/r02_bench34.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench34_icp.json').read().strip().splitlines()[-1])
    print('icp', round(d['value']), 'us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'cpu', d.get('cpu_baseline',{}).get('value'))
    t=json.loads(open('gpurun_out/r02_bench34_tsdf.json').read().strip().splitlines()[-1])
    r=t['roofline']; c=t['depth_color']['roofline']
    print('tsdf', round(t['value']), 'e2e', round(t['e2e']['value']), 'colour', round(t['depth_color']['value']), 'slam', t['dense_slam'].get('frames_per_sec'), t['dense_slam'].get('gpu_launches_per_frame'))
    print(' depth: events us', r['avg_launch_us'], 'device-timer us', r['avg_exec_us_device_timer'], 'frac', r['frac_device_timer'])
except Exception as e:
    print('bench parse failed', e)
PY
tail -3 gpurun_out/r02_bench34.err
