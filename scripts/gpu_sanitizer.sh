#!/bin/bash
# one compute-sanitizer tool per gpurun call (B200_PROFILING.md): bash scripts/gpu_sanitizer.sh memcheck|racecheck
TOOL=${1:-memcheck}
mkdir -p gpurun_out
timeout 300 python scripts/sanitizer_workload.py > gpurun_out/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -8 gpurun_out/sanitizer_plain.log; exit 1; }
tail -1 gpurun_out/sanitizer_plain.log | cut -c1-300
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 30 --error-exitcode 9 python scripts/sanitizer_workload.py > gpurun_out/sanitizer_$TOOL.log 2>&1
echo "compute-sanitizer $TOOL rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer workload ok|Error|hazard|Invalid|=========     at" gpurun_out/sanitizer_$TOOL.log | head -30
