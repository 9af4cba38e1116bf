#!/bin/bash
# GPU visit 26: parity suite + Navier-Stokes bench with the running single-precision screen in g_max_dt (Navier-Stokes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pde navier_stokes > gpurun_out/bench_ns.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ns.log
for f in pytest_gpu bench_ns; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-200; done
