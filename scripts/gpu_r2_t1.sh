#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_t1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_t1.log
timeout 300 python scripts/train_syncbn_n1.py 2>&1 | tail -2
