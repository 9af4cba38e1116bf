#!/bin/bash
# GPU visit 8: parity tests, 3-D Euler (deformed / cartesian) and Navier-Stokes bench lines after the bank-conflict line map, ncu of both Local kernels
TAG=${1:-r01i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_def.log 2>&1; echo "rc=$?" >> gpurun_out/bench_def.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mesh cartesian > gpurun_out/bench_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_car.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pde navier_stokes > gpurun_out/bench_ns.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ns.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pde navier_stokes --mesh cartesian > gpurun_out/bench_ns_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ns_car.log
KN='regex:local_|neighbor_|max_dt_|bc_kernel|prolong_kernel|restrict_kernel|cfl_|fill_kernel|write_face|g_.*_kernel|ns_.*_kernel|gather_faces|scatter_faces|admissible'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -s 30 -c 50 --csv --log-file gpurun_out/launches_${TAG}_euler.csv \
  python bench.py --n 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_euler.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:local_euler_pipe|neighbor_euler" -s 8 -c 4 -f -o gpurun_out/prof_${TAG}_euler \
  python bench.py --n 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_euler.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ns_local" -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_ns \
  python bench.py --n 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --pde navier_stokes > gpurun_out/ncu_full_ns.log 2>&1
for f in pytest_gpu bench_def bench_car bench_ns bench_ns_car; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-400; done
