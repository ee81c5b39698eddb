#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./profiles/probes/tma_probe | tee gpurun_out/r02_tma_probe.log
timeout 300 python profiles/r02_repro_tsdf.py 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_tsdf_gpu.py tests/test_odometry_gpu.py -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/r02_pytest16a.log
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest16b.log
cat gpurun_out/slam_100_frames_vs_oracle.txt
timeout 900 python bench.py --steps 5 --warmup 3 --metric tsdf > gpurun_out/r02_bench16_tsdf.json 2> gpurun_out/r02_bench16.err; tail -c 1800 gpurun_out/r02_bench16_tsdf.json
O3DB_TSDF_NO_TILE=1 timeout 900 python bench.py --steps 5 --warmup 3 --metric tsdf --skip-cpu > gpurun_out/r02_bench16_tsdf_notile.json 2>> gpurun_out/r02_bench16.err; tail -c 600 gpurun_out/r02_bench16_tsdf_notile.json
TSDF_COLOR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"integrate16_kernel|touch_kernel" -s 60 -c 2 \
    -o gpurun_out/r02_tsdf_depth16 python profiles/profile_workload.py tsdf > gpurun_out/r02_ncu16b.log 2>&1; tail -2 gpurun_out/r02_ncu16b.log
TSDF_COLOR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"integrate16_kernel" -s 30 -c 1 \
    -o gpurun_out/r02_tsdf_color16 python profiles/profile_workload.py tsdf > gpurun_out/r02_ncu16c.log 2>&1; tail -2 gpurun_out/r02_ncu16c.log
