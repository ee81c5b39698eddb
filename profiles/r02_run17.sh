#!/bin/bash
# 2 GPUs: sharded == single on both transports, bench N=2 (weak + strong + configs[3]); then N=1 sanity of TSDF touch change
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
timeout 900 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -k "sharded" 2>&1 | tail -15 | tee gpurun_out/r02_pytest17a.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py 2>&1 | tail -4 | tee gpurun_out/r02_multigpu17.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --skip-tsdf > gpurun_out/r02_bench17_n2.json 2> gpurun_out/r02_bench17_n2.err; tail -c 3000 gpurun_out/r02_bench17_n2.json; tail -5 gpurun_out/r02_bench17_n2.err
O3DB_COMM_NO_PEER=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --skip-tsdf --config3-points 500000 > gpurun_out/r02_bench17_n2_nccl.json 2>> gpurun_out/r02_bench17_n2.err; tail -c 1500 gpurun_out/r02_bench17_n2_nccl.json
timeout 900 python -m pytest tests/test_tsdf_gpu.py tests/test_forwarders_gpu.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r02_pytest17b.log
