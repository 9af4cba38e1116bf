#!/bin/bash
# GPU visit 20: parity suite + bench lines with the running single-precision screen in the global max_dt reduction
TAG=${1:-r01r}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_def.log 2>&1; echo "rc=$?" >> gpurun_out/bench_def.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pde navier_stokes > gpurun_out/bench_ns.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ns.log
timeout 900 python bench.py --steps 10 --warmup 3 --dim 2 --no-cpu-baseline > gpurun_out/bench_2d_def.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_def.log
KN='regex:local_|neighbor_|max_dt_|bc_kernel|prolong_kernel|restrict_kernel|write_face|g_.*_kernel|ns_.*_kernel|admissible'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -c 60 --csv --log-file gpurun_out/launches_${TAG}_euler.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_euler.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:max_dt_euler" -s 3 -c 2 -f -o gpurun_out/prof_${TAG}_maxdt \
  python bench.py --n 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_maxdt.log 2>&1
for f in pytest_gpu bench_def bench_ns bench_2d_def; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-200; done
