#!/bin/bash
# GPU visit 13: parity suite, 2-D Navier-Stokes bench lines with the batched reconcile kernel, launch list of the 2-D NS step
TAG=${1:-r01n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --dim 2 --no-cpu-baseline --pde navier_stokes > gpurun_out/bench_2d_ns.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_ns.log
timeout 900 python bench.py --steps 10 --warmup 3 --dim 2 --no-cpu-baseline --pde navier_stokes --mesh cartesian > gpurun_out/bench_2d_ns_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_ns_car.log
KN='regex:local_|neighbor_|max_dt_|bc_kernel|prolong_kernel|restrict_kernel|write_face|g_.*_kernel|ns_.*_kernel|admissible'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -c 80 --csv --log-file gpurun_out/launches_${TAG}_2d_ns.csv \
  python bench.py --dim 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --pde navier_stokes > gpurun_out/ncu_launch_2d_ns.log 2>&1
for f in pytest_gpu bench_2d_ns bench_2d_ns_car; do echo "== $f"; tail -n 4 gpurun_out/$f.log | cut -c1-300; done
