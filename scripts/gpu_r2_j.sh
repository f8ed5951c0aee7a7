#!/bin/bash
mkdir -p gpurun_out
for c in 3 4 5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg$c.log 2> gpurun_out/bench_cfg$c.err; echo "bench cfg$c rc=$?"
  tail -c 1800 gpurun_out/bench_cfg$c.log; tail -2 gpurun_out/bench_cfg$c.err
done
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-train --no-extra > gpurun_out/bench_short.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-train --no-extra > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
