#!/bin/bash
# validation of HEAD after the re-entry: full GPU suite, smoke, the default bench line (what the driver runs)
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_v1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_v1.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_v1.log 2>&1; tail -2 gpurun_out/smoke_v1.log
timeout 600 python bench.py > gpurun_out/bench_v1.log 2> gpurun_out/bench_v1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_v1.log; tail -5 gpurun_out/bench_v1.err
