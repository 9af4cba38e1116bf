#!/bin/bash
# GPU visit 12: parity suite with the config-class tests and the 2-D batched Navier-Stokes kernel, racecheck of its aliasing, 2-D NS bench + ncu
TAG=${1:-r01m}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -q -x -k "navier_stokes_2d_line_kernel and 6-37" > gpurun_out/racecheck_ns2d.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck_ns2d.log
timeout 900 python bench.py --steps 10 --warmup 3 --dim 2 --no-cpu-baseline --pde navier_stokes > gpurun_out/bench_2d_ns.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_ns.log
timeout 900 python bench.py --steps 10 --warmup 3 --dim 2 --no-cpu-baseline --pde navier_stokes --mesh cartesian > gpurun_out/bench_2d_ns_car.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2d_ns_car.log
timeout 900 ncu --set full --clock-control none -k "regex:ns_local|g_reconcile|g_neighbor" -s 8 -c 5 -f -o gpurun_out/prof_${TAG}_2d_ns \
  python bench.py --dim 2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --pde navier_stokes > gpurun_out/ncu_full_2d_ns.log 2>&1
for f in pytest_gpu racecheck_ns2d bench_2d_ns bench_2d_ns_car; do echo "== $f"; tail -n 4 gpurun_out/$f.log | cut -c1-300; done
