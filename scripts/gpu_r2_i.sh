#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|passed|failed" gpurun_out/pytest_all.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','tf32_backbone','hot_path','e2e','configs','train'): print(k, l.get(k))
print(l['kernels'])
PY
tail -3 gpurun_out/bench.err
bash scripts/gpu_sanitizer.sh memcheck
