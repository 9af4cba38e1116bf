#!/bin/bash
# GPU visit 10: parity tests, racecheck of the aliased Navier-Stokes Local kernel, NS bench lines with three resident CTAs per SM, ncu
TAG=${1:-r01k}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -q -x -k "navier_stokes_3d_line_kernel" > gpurun_out/racecheck_ns.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck_ns.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pde navier_stokes > gpurun_out/bench_ns.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ns.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pde navier_stokes --mesh cartesian > gpurun_out/bench_ns_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ns_car.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ns_local" -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_ns \
  python bench.py --n 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --pde navier_stokes > gpurun_out/ncu_full_ns.log 2>&1
for f in pytest_gpu racecheck_ns bench_ns bench_ns_car; do echo "== $f"; tail -n 4 gpurun_out/$f.log | cut -c1-300; done
