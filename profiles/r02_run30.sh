#!/bin/bash
# round-end records on two B200: sharded == single (both transports), bench N=2 with the in-kernel exchange
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py 2>&1 | grep multigpu | tee gpurun_out/r02_multigpu30.log
O3DB_COMM_NO_PEER=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tests/multigpu_check.py 2>&1 | grep multigpu | tee -a gpurun_out/r02_multigpu30.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench30_n2.json 2> gpurun_out/r02_bench30_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench30_n2.json').read().strip().splitlines()[-1])
m=d['multi_gpu']; t=d.get('tsdf') or {}
print('N=2 value',round(d['value']),'kern_us',round(d['roofline']['avg_launch_us'],1),'e2e',round(d['e2e']['value']),'strong',round(m['strong_scaling']['value']), m['transport'][:30], 'cfg3 s', m['config3'].get('seconds_total'))
print('tsdf', t.get('depth_only',{}).get('value'), t.get('depth_color',{}).get('value'), t.get('dense_slam'))
PY
O3DB_COMM_NO_PEER=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --skip-tsdf --config3-points 500000 > gpurun_out/r02_bench30_n2_nccl.json 2>> gpurun_out/r02_bench30_n2.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench30_n2_nccl.json').read().strip().splitlines()[-1])
print('NCCL N=2 value',round(d['value']),'kern_us',round(d['roofline']['avg_launch_us'],1), d['multi_gpu']['transport'][:30])"
