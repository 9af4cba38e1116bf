#!/bin/bash
# GPU visit 15: parity suite (fused admissibility), headline + Cartesian + 2-D bench lines with the fused-check measurement in `aux`
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_def.log 2>&1; echo "rc=$?" >> gpurun_out/bench_def.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mesh cartesian > gpurun_out/bench_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_car.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dim 2 > gpurun_out/bench_2d_def.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_def.log
for f in pytest_gpu bench_def bench_car bench_2d_def; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-200; done
