#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "pointwise" > gpurun_out/pytest_h1.log 2>&1; echo "pytest(pointwise) rc=$?"; tail -4 gpurun_out/pytest_h1.log
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|passed|failed" gpurun_out/pytest_all.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 5000 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 300 python scripts/profile_head.py 3 > gpurun_out/profile_head_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'depth_losses_kernel|gather_embed_nhwc_kernel' -s 2 -c 2 -f \
   -o gpurun_out/prof_r2_small python scripts/profile_head.py 2 > gpurun_out/ncu_small_full.log 2>&1
echo "ncu rc=$?"
