#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_k.log
timeout 900 python bench.py > gpurun_out/bench_k.log 2> gpurun_out/bench_k.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/bench_k.log; tail -2 gpurun_out/bench_k.err
timeout 600 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg5.log 2> gpurun_out/bench_cfg5.err; echo "bench cfg5 rc=$?"; tail -c 1500 gpurun_out/bench_cfg5.log
timeout 300 python scripts/ncu_step.py 2 > gpurun_out/step_k.log 2>&1; echo "step rc=$?"; tail -1 gpurun_out/step_k.log
timeout 300 python scripts/profile_step.py infer > gpurun_out/profile_step_k.log 2>&1; echo "profile rc=$?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python scripts/ncu_step.py 2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"; tail -2 gpurun_out/ncu_launches.log
