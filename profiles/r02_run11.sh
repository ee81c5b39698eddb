#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== variant 2"; timeout 300 python profiles/icp_iter_times.py 30 3 2>&1 | head -2 | tee gpurun_out/r02_iter11.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench11.json 2> gpurun_out/r02_bench11.err; tail -c 3000 gpurun_out/r02_bench11.json
