#!/bin/bash
# quick re-validation after a kernel change: full GPU suite + the inference bench line (no CPU / training / extra legs)
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_q.log
timeout 90 python bench.py --no-cpu --no-train --no-extra > gpurun_out/bench_q.log 2> gpurun_out/bench_q.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_q.log').read().strip().splitlines()[-1])
print(l["ms_per_step"], l["value"], "eager", l["eager"]["ms_per_step"], "hot", l["hot_path"]["ms_per_step"], "e2e", l["e2e"]["ms_per_step"], "bf16", l["bf16_mode"]["ms_per_step"])
print({k:v for k,v in l["kernels"].items() if k in ("encoder_layers_tc","upsample_concat_nhwc","conv3x3_head","up4.conv_a","patch_embed")})
PY
