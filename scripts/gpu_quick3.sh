#!/bin/bash
# quick look: full parity suite, smoke(), Navier-Stokes bench lines (3-D deformed, 2-D deformed) after hoisting the loads of Neighbor_reconcile
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pde navier_stokes > gpurun_out/bench_ns.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ns.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pde navier_stokes --dim 2 > gpurun_out/bench_2d_ns.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_ns.log
for f in pytest_gpu smoke bench_ns bench_2d_ns; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-200; done
