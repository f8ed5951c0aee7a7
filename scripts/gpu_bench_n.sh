#!/bin/bash
# usage: gpu_bench_n.sh N [extra bench.py flags]   (under gpurun --gpus N): the driver's launch line for N ranks
N=${1:-8}; shift
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; tail -c 2500 gpurun_out/bench_n$N.log; grep -v "^$" gpurun_out/bench_n$N.err | grep -iv "warn\|OMP_NUM\|\*\*\*" | tail -5 | cut -c1-300
