#!/bin/bash
# graphed training step: parity test at N=1, eager vs graph timing at BASELINE config 2 (and config 3)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -q -x -k "graphed_train_step" > gpurun_out/pytest_v2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_v2.log
BATCH=16 STEPS=3 timeout 400 python scripts/train_graph_check.py > gpurun_out/tg_n1.log 2> gpurun_out/tg_n1.err; echo "check rc=$?"; tail -c 1500 gpurun_out/tg_n1.log; tail -5 gpurun_out/tg_n1.err
CFG=3 BATCH=16 STEPS=3 timeout 400 python scripts/train_graph_check.py > gpurun_out/tg_n1_cfg3.log 2> gpurun_out/tg_n1_cfg3.err; echo "check cfg3 rc=$?"; tail -c 1500 gpurun_out/tg_n1_cfg3.log; tail -5 gpurun_out/tg_n1_cfg3.err
