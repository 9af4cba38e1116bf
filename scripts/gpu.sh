#!/bin/bash
# One script for every GPU visit (replaces the per-visit one-offs of round 1). Runs the named steps in order on the box gpurun
# provides, logs under gpurun_out/ and prints the tail of each log.
#   gpurun --timeout 1500 -- 'bash scripts/gpu.sh tests smoke ref bench'
#   gpurun --timeout 1200 -- 'TAG=r02a bash scripts/gpu.sh launches full:local_euler_pipe'
#   gpurun --gpus 2 --timeout 1200 -- 'N=2 bash scripts/gpu.sh multi_check multi_bench'
# steps:
#   tests            pytest -m gpu (whole suite)           tests:<expr>   pytest -m gpu -k <expr>
#   smoke            __graft_entry__.smoke()
#   ref              bench.py --impl reference             bench          default bench line
#   bench:<args>     bench.py --steps 10 --warmup 3 --no-cpu-baseline <args with , for space>   (e.g. bench:--pde,navier_stokes)
#   launches[:<args>]  ncu launch list (gpu__time_duration) of bench steps at --n 64
#   full:<regex>[:<args>]  ncu --set full capture of the kernels matching <regex> (report named after regex + args: several per visit are fine)
#   multi_check / multi_bench[:<args>]   torchrun over $N GPUs: scripts/multigpu_check.py / bench.py --gpus $N
#   py:<script>[:<args>]   python <script> <args>
TAG=${TAG:-r02}
N=${N:-2}
mkdir -p gpurun_out
(nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader; echo "nproc $(nproc)"; lscpu | grep -m1 "Model name"; free -g | head -2; ulimit -a | grep -E "memory|locked") > gpurun_out/box_$TAG.log 2>&1
logs=(box_$TAG)
run() { # name, command...
  local name=$1; shift
  "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log
  logs+=($name)
}
KN='regex:local_|neighbor_|max_dt_|bc_kernel|prolong_kernel|restrict_kernel|reduce_min|write_face|stage_|ns_|g_|gather|scatter|admissible|scale_dt'
port=29520
for step in "$@"; do
  kind=${step%%:*}; rest=""; [[ "$step" == *:* ]] && rest=${step#*:}
  a1=${rest%%:*}; a2=""; [[ "$rest" == *:* ]] && a2=${rest#*:}
  a1s=${a1//,/ }; a2s=${a2//,/ }
  case $kind in
    tests) if [ -n "$a1" ]; then run pytest_gpu_$TAG timeout 1500 python -m pytest tests -m gpu -q -k "$a1s"; else run pytest_gpu_$TAG timeout 1800 python -m pytest tests -m gpu -q; fi ;;
    smoke) run smoke_$TAG timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ;;
    ref) run bench_reference_$TAG timeout 900 python bench.py --impl reference ;;
    bench) if [ -n "$a1" ]; then name=bench_$(echo "$a1" | tr -c 'a-zA-Z0-9\n' '_')_$TAG; run $name timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $a1s;
           else run bench_default_$TAG timeout 900 python bench.py; fi ;;
    launches) run ncu_launch_$TAG timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KN" -s 20 -c 60 --csv \
                --log-file gpurun_out/launches_$TAG.csv python bench.py --n 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $a1s ;;
    full) name=$(echo "$a1$a2" | tr -c 'a-zA-Z0-9\n' '_'); run ncu_full_${name}_$TAG timeout 900 ncu --set full --clock-control none --import-source on \
                -k "regex:$a1" -s 4 -c 2 -f -o gpurun_out/prof_${name}_$TAG python bench.py --n 64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $a2s ;;
    multi_check) port=$((port+1)); run multigpu_check_${N}_$(echo "$a1" | tr -c 'a-zA-Z0-9\n' '_')_$TAG timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
                --master-addr 127.0.0.1 --master-port $port scripts/multigpu_check.py $a1s ;;
    multi_bench) port=$((port+1)); name=bench_multi_${N}_$(echo "$a1" | tr -c 'a-zA-Z0-9\n' '_')_$TAG; run $name timeout 900 python -m torch.distributed.run \
                --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 10 --warmup 3 $a1s ;;
    py) name=$(basename "$a1" .py)_$TAG; run $name timeout 1200 python $a1 $a2s ;;
    *) echo "unknown step $step" ;;
  esac
done
for f in "${logs[@]}"; do echo "== $f"; tail -n 4 gpurun_out/$f.log | cut -c1-1500; done
