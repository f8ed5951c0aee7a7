#!/bin/bash
mkdir -p gpurun_out
timeout 500 python scripts/pdl_sweep.py > gpurun_out/pdl_sweep.log 2> gpurun_out/pdl_sweep.err; echo "sweep rc=$?"; tail -c 1500 gpurun_out/pdl_sweep.log; tail -3 gpurun_out/pdl_sweep.err
