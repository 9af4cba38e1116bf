#!/bin/bash
# GPU visit: smoke, parity tests, both bench arms at full size
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full.log
timeout 600 python bench.py --steps 10 --warmup 3 --mesh cartesian --no-cpu-baseline > gpurun_out/bench_full_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_full_car.log
for f in smoke pytest_gpu bench_ref bench_full bench_full_car; do echo "== $f"; tail -n 4 gpurun_out/$f.log | cut -c1-2500; done
