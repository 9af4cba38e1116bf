#!/bin/bash
# quick GPU check: parity tests + one full-size bench line per mesh type
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_def.log 2>&1; echo "rc=$?" >> gpurun_out/bench_def.log
timeout 600 python bench.py --steps 10 --warmup 3 --mesh cartesian --no-cpu-baseline > gpurun_out/bench_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_car.log
for f in pytest_gpu bench_def bench_car; do echo "== $f"; tail -n 4 gpurun_out/$f.log; done
