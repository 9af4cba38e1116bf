#!/bin/bash
# GPU visit 14: parity suite; compute-sanitizer memcheck over the tests of this round's new kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py tests/test_config_classes.py -q -x -m gpu \
  -k "set_jacobian or shared_normals or calc_jacobian or vertex_sharing or av_glue or aux_bcs or is_admissible or riemann or 2d_line or cylinder or (test_box and 2-6-101) or 3d_line_kernel" \
  > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck.log
for f in pytest_gpu memcheck; do echo "== $f"; tail -n 6 gpurun_out/$f.log | cut -c1-300; done
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck.log
