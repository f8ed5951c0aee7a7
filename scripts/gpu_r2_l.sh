#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_l.log
timeout 900 python bench.py --no-cpu > gpurun_out/bench_l.log 2> gpurun_out/bench_l.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_l.log; tail -2 gpurun_out/bench_l.err
timeout 300 python scripts/ncu_step.py 2 > gpurun_out/step_l.log 2>&1; echo "step rc=$?"; tail -1 gpurun_out/step_l.log
timeout 300 python scripts/profile_step.py infer > gpurun_out/profile_step_l.log 2>&1; echo "profile rc=$?"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -c 130 \
    -k regex:"gather_embed_nhwc|depth_losses|upsample_nhwc|bias_act_pool|se_gate|head_chain|patch_embed|pointwise_x3|conv3x3_kernel" \
    -o gpurun_out/prof_r2_step -f python scripts/ncu_step.py 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/prof_r2_step.ncu-rep
