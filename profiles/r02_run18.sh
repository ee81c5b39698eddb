#!/bin/bash
# 2 GPUs, flag-in-data peer exchange: sharded == single, bench N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py 2>&1 | grep multigpu | tee gpurun_out/r02_multigpu18.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --skip-tsdf > gpurun_out/r02_bench18_n2.json 2> gpurun_out/r02_bench18_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench18_n2.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'kern_us',d['roofline']['avg_launch_us'],'e2e',round(d['e2e']['value']),'strong',d['multi_gpu']['strong_scaling']['value'], d['multi_gpu']['transport'][:40], 'cfg3', d['multi_gpu']['config3'].get('seconds_total'))
PY
