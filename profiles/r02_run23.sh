#!/bin/bash
# 8 GPUs: sharded == single at 8 ranks (peer exchange), full bench N=8 (weak + strong + configs[3] + TSDF replicas)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py 2>&1 | grep -E "multigpu|Error|error" | head -8 | tee gpurun_out/r02_multigpu23.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench23_n8.json 2> gpurun_out/r02_bench23_n8.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench23_n8.json').read().strip().splitlines()[-1])
m=d['multi_gpu']
print('N=8 value',round(d['value']),'kern_us',round(d['roofline']['avg_launch_us'],1),'e2e',round(d['e2e']['value']),'strong',round(m['strong_scaling']['value']), m['transport'][:30])
print('cfg3', json.dumps(m['config3'])[:700])
t=d.get('tsdf') or {}
print('tsdf', t.get('depth_only',{}).get('value'), t.get('depth_color',{}).get('value'), (t.get('dense_slam') or {}).get('frames_per_sec'))
PY
tail -3 gpurun_out/r02_bench23_n8.err
