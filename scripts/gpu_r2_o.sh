#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_o.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_o.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"pointwise_x3" -s 2 -c 1 \
    -o gpurun_out/prof_r2_pointwise -f python scripts/ncu_step.py 2 > gpurun_out/ncu_pw.log 2>&1
echo "ncu pw rc=$?"; tail -2 gpurun_out/ncu_pw.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"depth_losses" -c 1 \
    -o gpurun_out/prof_r2_losses -f python scripts/ncu_step.py 2 > gpurun_out/ncu_ls.log 2>&1
echo "ncu ls rc=$?"; tail -2 gpurun_out/ncu_ls.log
ls -la gpurun_out/*.ncu-rep
