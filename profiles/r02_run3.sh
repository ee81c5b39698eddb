#!/bin/bash
# round 2, GPU call 3: staged (TMA + cp.async ring) vs direct ICP kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_icp_gpu.py -m gpu -x -q > gpurun_out/r02_t3.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_t3.log
for v in 2 1; do
  echo "== variant $v"
  ICP_VARIANT=$v timeout 300 python profiles/icp_iter_times.py 30 3 2>&1 | tee gpurun_out/r02_iter3_v$v.log | head -2
  ICP_VARIANT=$v ICP_ITERS=12 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:icp_iteration --csv \
      --log-file gpurun_out/r02_launches3_v$v.csv python profiles/profile_workload.py icp > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_launches3_v$v.csv")) if len(r)>5 and r[0].isdigit()]
print("kernel_us(warm, serialized):", " ".join(r[-1] for r in rows))
PY
done
bash profiles/tune_icp.sh "-DICP_MIN_BLOCKS=2" "-DICP_MIN_BLOCKS=4" "-DICP_DEFAULT_VARIANT=1" "" 2>&1 | tee gpurun_out/r02_tune3.log
ICP_ITERS=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_iteration_kernel -s 10 -c 1 \
    -o gpurun_out/r02_icp_staged python profiles/profile_workload.py icp > gpurun_out/r02_ncu3.log 2>&1; tail -2 gpurun_out/r02_ncu3.log
