#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q -k "conv3x3 or decoder or config2 or bf16 or pointwise" > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_q.log
timeout 600 python bench.py --no-cpu --no-train --no-extra > gpurun_out/bench_q.log 2> gpurun_out/bench_q.err; echo "bench rc=$?"; tail -c 2400 gpurun_out/bench_q.log | head -c 1500; tail -2 gpurun_out/bench_q.err
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"pointwise_x3" -s 1 -c 2 \
    -o gpurun_out/prof_r2_pointwise2 -f python scripts/ncu_step.py 2 > gpurun_out/ncu_pw.log 2>&1
echo "ncu pw rc=$?"; tail -2 gpurun_out/ncu_pw.log
ls -la gpurun_out/*.ncu-rep
