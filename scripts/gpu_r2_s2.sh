#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_s2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_s2.log
timeout 600 python bench.py --no-cpu --no-train --no-extra > gpurun_out/bench_s2.log 2> gpurun_out/bench_s2.err; echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_s2.log').read().strip().splitlines()[-1])
print(l["ms_per_step"], "stem", l["kernels"].get("stem_conv"), "tf32bb", l["tf32_backbone"]["ms_per_step"], "bf16", l["bf16_mode"]["ms_per_step"])
PY
