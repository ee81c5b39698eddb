#!/bin/bash
# smoke(), SLAM launch list of the round-end build, odometry block-count variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke31.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pyramid|odometry|clip_transform|integrate|range_|ray_cast|touch" -c 300 --csv --log-file gpurun_out/r02_launches31_slam.csv python profiles/profile_workload.py slam > gpurun_out/r02_ll31.log 2>&1; tail -1 gpurun_out/r02_ll31.log
bash profiles/tune_slam.sh "-DODO_BLOCKS_PER_SM=1" "-DODO_BLOCKS_PER_SM=2" "" 2>&1 | tee gpurun_out/r02_tune_slam31.log
