#!/bin/bash
# Round-2 re-entry baseline: GPU tests, bench, launch list, --set full captures (ICP steady state, TSDF depth-only + colour, touch)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest12.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench12.json 2> gpurun_out/r02_bench12.err; tail -c 2500 gpurun_out/r02_bench12.json
timeout 300 python profiles/icp_iter_times.py 30 3 2>&1 | head -3 | tee gpurun_out/r02_iter12.log
ICP_ITERS=12 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches12.csv python profiles/profile_workload.py all > gpurun_out/r02_ll12.log 2>&1
ICP_ITERS=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_iteration -s 10 -c 1 \
    -o gpurun_out/r02_icp12 python profiles/profile_workload.py icp > gpurun_out/r02_ncu12a.log 2>&1; tail -2 gpurun_out/r02_ncu12a.log
TSDF_COLOR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"integrate_kernel|touch_kernel" -s 60 -c 2 \
    -o gpurun_out/r02_tsdf_depth12 python profiles/profile_workload.py tsdf > gpurun_out/r02_ncu12b.log 2>&1; tail -2 gpurun_out/r02_ncu12b.log
TSDF_COLOR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"integrate_kernel" -s 30 -c 1 \
    -o gpurun_out/r02_tsdf_color12 python profiles/profile_workload.py tsdf > gpurun_out/r02_ncu12c.log 2>&1; tail -2 gpurun_out/r02_ncu12c.log
ls -la gpurun_out
