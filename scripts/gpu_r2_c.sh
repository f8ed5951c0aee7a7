#!/bin/bash
# round-2 call C: full GPU suite -> bench -> ncu full capture of the hot-path kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 5500 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 300 python scripts/profile_head.py 3 > gpurun_out/profile_head_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:'head_chain_kernel|depth_losses_kernel|conv3x3_kernel|patch_embed_kernel|gather_embed_nhwc_kernel|attention_kernel' \
   -s 6 -c 8 -f -o gpurun_out/prof_r2_head python scripts/profile_head.py 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
