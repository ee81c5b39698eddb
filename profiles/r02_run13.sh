#!/bin/bash
# BASELINE-size parity tests + the reference-solver colour gradients on the GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -s --durations=10 2>&1 | tail -40 | tee gpurun_out/r02_pytest13a.log
timeout 900 python -m pytest tests/test_icp_gpu.py -m gpu -q -k "color" 2>&1 | tail -15 | tee gpurun_out/r02_pytest13b.log
