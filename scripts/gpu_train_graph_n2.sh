#!/bin/bash
# graphed training step at N=2: SyncBatchNorm peer-memory exchange with device-side epochs inside the graph
mkdir -p gpurun_out
BATCH=16 STEPS=4 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  scripts/train_graph_check.py > gpurun_out/tg_n2.log 2> gpurun_out/tg_n2.err; echo "check rc=$?"; tail -c 1500 gpurun_out/tg_n2.log; tail -8 gpurun_out/tg_n2.err
