#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_n.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_n.log
timeout 900 python bench.py --no-cpu > gpurun_out/bench_n.log 2> gpurun_out/bench_n.err; echo "bench rc=$?"; tail -c 3600 gpurun_out/bench_n.log; tail -2 gpurun_out/bench_n.err
timeout 300 python scripts/profile_step.py infer > gpurun_out/profile_step_n.log 2>&1; echo "profile rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python scripts/ncu_step.py 2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none -c 30 \
    -k regex:"gather_embed_nhwc|depth_losses|upsample_nhwc|conv3x3_kernel|head_chain|bias_act_pool|pointwise_x3|patch_embed" \
    -o gpurun_out/prof_r2_step -f python scripts/ncu_step.py 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_r2_step.ncu-rep --page raw --csv > gpurun_out/prof_r2_step_raw.csv 2>/dev/null
ls -la gpurun_out/
sz=$(stat -c %s gpurun_out/prof_r2_step.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 50000000 ]; then rm -f gpurun_out/prof_r2_step.ncu-rep; echo "report dropped (too large), csv kept"; fi
