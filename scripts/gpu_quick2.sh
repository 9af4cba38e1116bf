#!/bin/bash
# quick look: Euler parity subset + the headline bench line (kernel_seconds_per_step shows the max_dt kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "soup_all or test_box or soup_options or full_size or update_euler" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
true > gpurun_out/bench_def.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --dim 2 > gpurun_out/bench_2d_def.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_def.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --dim 2 --mesh cartesian > gpurun_out/bench_2d_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_car.log
for f in pytest_gpu bench_def bench_2d_def bench_2d_car; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-200; done
