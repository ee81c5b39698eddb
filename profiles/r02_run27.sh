#!/bin/bash
# dense-SLAM loop under torchrun with 2 ranks (segments 0.. and 100..): does the multi-process slowdown of the N=8 bench reproduce?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 profiles/slam_time.py 100 2>&1 | grep -E "slam ms|Error" | tee gpurun_out/r02_slam27.log
