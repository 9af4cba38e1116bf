#!/bin/bash
# GPU visit 16: parity suite (update_euler with and without graph), launch-bound mesh measurement
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python scripts/small_mesh_bench.py > gpurun_out/small_mesh.log 2>&1; echo "rc=$?" >> gpurun_out/small_mesh.log
for f in pytest_gpu small_mesh; do echo "== $f"; tail -n 6 gpurun_out/$f.log | cut -c1-700; done
