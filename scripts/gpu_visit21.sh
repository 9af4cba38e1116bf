#!/bin/bash
# GPU visit 21: lean (67 KB, three CTAs per SM) vs classic (105 KB, two CTAs) layout of the 3-D deformed Euler Local kernel, same box
TAG=${1:-r01s}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -q -x -k "test_box and (3-6-5-True or 3-4-6-True)" > gpurun_out/racecheck_pipe3d.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck_pipe3d.log
for rep in 1 2; do
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pipe-mode 2 > gpurun_out/bench_def_classic_$rep.log 2>&1; echo "rc=$?" >> gpurun_out/bench_def_classic_$rep.log
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_def_lean_$rep.log 2>&1; echo "rc=$?" >> gpurun_out/bench_def_lean_$rep.log
done
timeout 900 ncu --set full --clock-control none -k "regex:local_euler_pipe_kernel" -s 4 -c 2 -f -o gpurun_out/prof_${TAG}_euler \
  python bench.py --n 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_euler.log 2>&1
for f in pytest_gpu racecheck_pipe3d bench_def_classic_1 bench_def_lean_1 bench_def_classic_2 bench_def_lean_2; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-200; done
