#!/bin/bash
# usage: gpu_r2_n8.sh N   (under gpurun --gpus N)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; tail -c 3500 gpurun_out/bench_n$N.log; tail -3 gpurun_out/bench_n$N.err
