#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"depth_losses" -c 1 \
    -o gpurun_out/prof_r2_losses2 -f python scripts/ncu_step.py 2 > gpurun_out/ncu_ls.log 2>&1
echo "ncu ls rc=$?"; tail -2 gpurun_out/ncu_ls.log
