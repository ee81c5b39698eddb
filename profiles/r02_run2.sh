#!/bin/bash
# round 2, GPU call 2: parallel final reduction + PDL — full parity suite, per-iteration times, knob sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t2.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_t2.log
timeout 300 python profiles/icp_iter_times.py 30 3 > gpurun_out/r02_iter2.log 2>&1; cat gpurun_out/r02_iter2.log
bash profiles/tune_icp.sh "-DICP_PDL=0" "-DICP_MIN_BLOCKS=4" "-DICP_THIN_FACTOR=32" "-DICP_THIN_FACTOR=32 -DICP_MIN_BLOCKS=4" "-DICP_FLUSH_EVERY=64" "" 2>&1 | tee gpurun_out/r02_tune2.log
