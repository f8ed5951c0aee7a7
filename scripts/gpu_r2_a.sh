#!/bin/bash
# round-2 bring-up call: bf16 UMMA layouts -> new tensor-core kernels -> model-level parity -> rest -> smoke
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 120 mde_biological_vision_systems_b200/lib/bf16_selftest > gpurun_out/bf16_selftest.log 2>&1; echo "bf16_selftest rc=$?"
cat gpurun_out/bf16_selftest.log
K1='split or conv3x3 or patch_embed or range_attention or upsample_concat_nhwc_pair or gemm_nt'
timeout 900 python -m pytest tests -m gpu -q --durations=5 -k "$K1" > gpurun_out/pytest_k1.log 2>&1; echo "pytest(k1) rc=$?"
tail -30 gpurun_out/pytest_k1.log
K2='head or decoder or full_model or mvit or config or before_attn or channels_last_model or autocast or full_size'
timeout 1500 python -m pytest tests -m gpu -q -s --durations=8 -k "($K2) and not ($K1)" > gpurun_out/pytest_k2.log 2>&1; echo "pytest(k2) rc=$?"
grep -E "passed|failed|Error|error|assert|max |vs oracle|golden" gpurun_out/pytest_k2.log | tail -40
timeout 900 python -m pytest tests -m gpu -q --durations=5 -k "not ($K1) and not ($K2)" > gpurun_out/pytest_rest.log 2>&1; echo "pytest(rest) rc=$?"
tail -8 gpurun_out/pytest_rest.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
