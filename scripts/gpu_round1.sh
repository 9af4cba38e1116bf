#!/bin/bash
# GPU visit: smoke, parity tests, small + full bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --n 40 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n40.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n40.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
timeout 600 python bench.py --steps 10 --warmup 3 --mesh cartesian --no-cpu-baseline > gpurun_out/bench_full_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full_car.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --n 60 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
for f in smoke pytest_gpu bench_n40 bench_full bench_full_car; do echo "== $f"; tail -n 4 gpurun_out/$f.log; done
