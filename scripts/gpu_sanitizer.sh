#!/bin/bash
# one compute-sanitizer tool per gpurun call (B200_PROFILING.md): bash scripts/gpu_sanitizer.sh memcheck|racecheck
TOOL=${1:-memcheck}
mkdir -p gpurun_out
K='conv3x3_tc or conv3x3_autograd or patch_embed_tc or test_head_golden or gemm_nt_tc or depth_losses_fused or range_attention or split_bf16 or upsample_concat_nhwc_pair or gather_into_encoder'
timeout 300 python -m pytest tests -m gpu -q -x -k "$K" > gpurun_out/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitizer_plain.log; exit 1; }
tail -1 gpurun_out/sanitizer_plain.log
timeout 2400 compute-sanitizer --tool $TOOL --print-limit 20 --error-exitcode 9 \
  python -m pytest tests -m gpu -q -x -k "$K" > gpurun_out/sanitizer_$TOOL.log 2>&1
echo "compute-sanitizer $TOOL rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" gpurun_out/sanitizer_$TOOL.log | tail -15
