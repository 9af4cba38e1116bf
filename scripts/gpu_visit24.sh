#!/bin/bash
# GPU visit 24: ncu launch list of the default bench command with the final binary of the round
TAG=${1:-r01u}
mkdir -p gpurun_out
KN='regex:local_|neighbor_|max_dt_|bc_kernel|prolong_kernel|restrict_kernel|write_face|g_.*_kernel|ns_.*_kernel|admissible'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -c 60 --csv --log-file gpurun_out/launches_${TAG}_euler.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_euler.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -c 80 --csv --log-file gpurun_out/launches_${TAG}_ns.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --pde navier_stokes > gpurun_out/ncu_launch_ns.log 2>&1
tail -n 2 gpurun_out/ncu_launch_euler.log | cut -c1-200
