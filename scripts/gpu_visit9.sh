#!/bin/bash
# GPU visit 9: parity tests (set_jacobian, is_admissible, Riemann BC), Euler bench lines with the line map restricted to the Cartesian kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_def.log 2>&1; echo "rc=$?" >> gpurun_out/bench_def.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mesh cartesian > gpurun_out/bench_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_car.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pde navier_stokes > gpurun_out/bench_ns.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ns.log
for f in pytest_gpu bench_def bench_car bench_ns; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-300; done
