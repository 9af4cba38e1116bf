#!/bin/bash
# GPU visit 23: memcheck of the kernels that changed last (lean 3-D Local, padded 2-D Local, running-screen max_dt)
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x \
  -k "(test_box and not navier) or other_local_kernels or time_step_scale or (running_screen and 3-3-30)" > gpurun_out/memcheck_r01v.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck_r01v.log
tail -n 6 gpurun_out/memcheck_r01v.log | cut -c1-200
