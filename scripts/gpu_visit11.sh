#!/bin/bash
# GPU visit 11 (final single-GPU pass of round 1): parity tests, every bench line, reference arm, launch lists and ncu --set full of the
# dominant kernels at the committed state
TAG=${1:-r01l}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_def.log 2>&1; echo "rc=$?" >> gpurun_out/bench_def.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "rc=$?" >> gpurun_out/bench_reference.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mesh cartesian > gpurun_out/bench_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_car.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pde navier_stokes > gpurun_out/bench_ns.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ns.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pde navier_stokes --mesh cartesian > gpurun_out/bench_ns_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ns_car.log
timeout 900 python bench.py --steps 10 --warmup 3 --dim 2 > gpurun_out/bench_2d_def.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_def.log
timeout 900 python bench.py --steps 10 --warmup 3 --dim 2 --no-cpu-baseline --mesh cartesian > gpurun_out/bench_2d_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_car.log
timeout 900 python bench.py --steps 10 --warmup 3 --dim 2 --no-cpu-baseline --pde navier_stokes > gpurun_out/bench_2d_ns.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_ns.log
KN='regex:local_|neighbor_|max_dt_|bc_kernel|prolong_kernel|restrict_kernel|cfl_|fill_kernel|write_face|g_.*_kernel|ns_.*_kernel|gather_faces|scatter_faces|admissible'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -c 60 --csv --log-file gpurun_out/launches_${TAG}_euler.csv \
  python bench.py --n 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_euler.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -c 80 --csv --log-file gpurun_out/launches_${TAG}_ns.csv \
  python bench.py --n 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --pde navier_stokes > gpurun_out/ncu_launch_ns.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -c 60 --csv --log-file gpurun_out/launches_${TAG}_2d.csv \
  python bench.py --dim 2 --n 512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_2d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:local_euler_pipe|neighbor_euler|max_dt_euler" -s 6 -c 5 -f -o gpurun_out/prof_${TAG}_euler \
  python bench.py --n 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_euler.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:local_euler_pipe" -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_euler_car \
  python bench.py --n 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --mesh cartesian > gpurun_out/ncu_full_euler_car.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:ns_local|ns_reconcile|g_.*_kernel" -s 10 -c 5 -f -o gpurun_out/prof_${TAG}_ns \
  python bench.py --n 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --pde navier_stokes > gpurun_out/ncu_full_ns.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:local_euler_pipe2d|neighbor_euler" -s 6 -c 3 -f -o gpurun_out/prof_${TAG}_2d \
  python bench.py --dim 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_2d.log 2>&1
du -sh gpurun_out
for f in pytest_gpu bench_def bench_reference bench_car bench_ns bench_ns_car bench_2d_def bench_2d_car bench_2d_ns; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-260; done
