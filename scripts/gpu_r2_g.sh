#!/bin/bash
# N = 2: default bench line (inference weak scaling + training legs with the overlapped gradient all-reduce and p2p SyncBN)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q -x -k "training_step_with_own or gather_into or depth_losses" > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_g.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench N=2 rc=$?"
tail -c 2500 gpurun_out/bench_n2.log; tail -3 gpurun_out/bench_n2.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench N=1 rc=$?"
tail -c 1500 gpurun_out/bench_n1.log
