#!/bin/bash
# round-2 call B: full GPU suite -> bench (config 2 default line) -> launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
tail -c 900 gpurun_out/bench_ref.log
if [ "$1" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-train --no-extra > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
fi
