#!/bin/bash
# GPU visit 22: 3-D Cartesian Euler Local kernel, classic (72 KB, vector line map, three CTAs) vs lean (52 KB, 128 registers, four CTAs), same box
TAG=${1:-r01t}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for rep in 1 2; do
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mesh cartesian > gpurun_out/bench_car_classic_$rep.log 2>&1; echo "rc=$?" >> gpurun_out/bench_car_classic_$rep.log
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mesh cartesian --pipe-mode 3 > gpurun_out/bench_car_lean4_$rep.log 2>&1; echo "rc=$?" >> gpurun_out/bench_car_lean4_$rep.log
done
timeout 900 ncu --set full --clock-control none -k "regex:local_euler_pipe" -s 4 -c 2 -f -o gpurun_out/prof_${TAG}_euler_car \
  python bench.py --n 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --mesh cartesian --pipe-mode 3 > gpurun_out/ncu_full_euler_car.log 2>&1
for f in pytest_gpu bench_car_classic_1 bench_car_lean4_1 bench_car_classic_2 bench_car_lean4_2; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-200; done
