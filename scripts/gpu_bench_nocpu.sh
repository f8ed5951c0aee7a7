#!/bin/bash
# the default bench line without the CPU-baseline leg (fits a short GPU slot)
mkdir -p gpurun_out
timeout 88 python bench.py --no-cpu > gpurun_out/bench_nocpu.log 2> gpurun_out/bench_nocpu.err; echo "bench rc=$?"; tail -c 900 gpurun_out/bench_nocpu.log
