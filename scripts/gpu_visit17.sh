#!/bin/bash
# GPU visit 17: full parity suite, 2-D Euler bench lines (with e2e) and ncu of the restructured 2-D Local kernel (three CTAs per SM)
TAG=${1:-r01o}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --dim 2 > gpurun_out/bench_2d_def.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_def.log
timeout 900 python bench.py --steps 10 --warmup 3 --dim 2 --no-cpu-baseline --mesh cartesian > gpurun_out/bench_2d_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_car.log
timeout 900 python bench.py --steps 10 --warmup 3 --dim 2 --no-cpu-baseline --pde navier_stokes > gpurun_out/bench_2d_ns.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_ns.log
KN='regex:local_|neighbor_|max_dt_|bc_kernel|prolong_kernel|restrict_kernel|write_face|g_.*_kernel|ns_.*_kernel|admissible'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -c 60 --csv --log-file gpurun_out/launches_${TAG}_2d.csv \
  python bench.py --dim 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_2d.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:local_euler_pipe2d" -s 4 -c 2 -f -o gpurun_out/prof_${TAG}_2d \
  python bench.py --dim 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_2d.log 2>&1
for f in pytest_gpu bench_2d_def bench_2d_car bench_2d_ns; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-200; done
