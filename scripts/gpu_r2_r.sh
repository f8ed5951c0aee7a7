#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r.log
timeout 600 python bench.py --no-cpu --no-train --no-extra > gpurun_out/bench_r.log 2> gpurun_out/bench_r.err; echo "bench rc=$?"; tail -c 2600 gpurun_out/bench_r.log | head -c 1900; tail -2 gpurun_out/bench_r.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python scripts/ncu_step.py 2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
