#!/bin/bash
# ncu captures: launch list of our kernels over one bench step + full-set capture of the dominant kernels.
# usage: bash scripts/gpu_profile.sh <tag> [kernel regex for the full capture]
TAG=${1:-r01}
PAT=${2:-local_euler_kernel}
mkdir -p gpurun_out
KN='regex:local_|neighbor_|max_dt_|bc_kernel|prolong_kernel|restrict_kernel|reduce_min|write_face|stage_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -s 20 -c 40 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --n 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$PAT" -s 4 -c 2 -f -o gpurun_out/prof_$TAG \
  python bench.py --n 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/
