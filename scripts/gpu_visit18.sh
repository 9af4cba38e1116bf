#!/bin/bash
# GPU visit 18: parity suite, racecheck + 2-D bench lines + ncu of the 2-D Euler Local kernel with padded strides / vector line accesses
TAG=${1:-r01p}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -q -x -k "test_box and (2-6-101 or 2-4-150 or 2-8-70)" > gpurun_out/racecheck_pipe2d.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck_pipe2d.log
timeout 900 python bench.py --steps 10 --warmup 3 --dim 2 > gpurun_out/bench_2d_def.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_def.log
timeout 900 python bench.py --steps 10 --warmup 3 --dim 2 --no-cpu-baseline --mesh cartesian > gpurun_out/bench_2d_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_car.log
timeout 900 ncu --set full --clock-control none -k "regex:local_euler_pipe2d" -s 4 -c 2 -f -o gpurun_out/prof_${TAG}_2d \
  python bench.py --dim 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_2d.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:local_euler_pipe2d" -s 4 -c 2 -f -o gpurun_out/prof_${TAG}_2d_car \
  python bench.py --dim 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --mesh cartesian > gpurun_out/ncu_full_2d_car.log 2>&1
for f in pytest_gpu racecheck_pipe2d bench_2d_def bench_2d_car; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-200; done
