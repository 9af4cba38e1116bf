#!/bin/bash
# multi-GPU visit: parity of the partitioned path (Euler + viscous) and weak-scaling bench lines
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py > gpurun_out/multigpu_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/multigpu_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_def_$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_def_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --pde navier_stokes > gpurun_out/bench_ns_$N.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ns_$N.log
for f in multigpu_check_$N bench_def_$N bench_ns_$N; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-1200; done
