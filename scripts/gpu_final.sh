#!/bin/bash
# what the driver runs at round end: GPU parity suite, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.log 2>&1; echo "rc=$?" >> gpurun_out/bench_reference.log
timeout 900 python bench.py > gpurun_out/bench_def.log 2>&1; echo "rc=$?" >> gpurun_out/bench_def.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pde navier_stokes > gpurun_out/bench_ns.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ns.log
for f in pytest_gpu smoke bench_reference bench_def bench_ns; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-220; done
