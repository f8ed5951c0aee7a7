#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-extra > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; tail -c 700 gpurun_out/bench_n$N.log; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
