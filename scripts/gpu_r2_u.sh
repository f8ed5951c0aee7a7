#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "loss or silog or chamfer or config2 or golden or train" > gpurun_out/pytest_u.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_u.log
timeout 600 python bench.py --no-cpu --no-train --no-extra > gpurun_out/bench_u.log 2> gpurun_out/bench_u.err; echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_u.log').read().strip().splitlines()[-1])
print(l["ms_per_step"], l["kernels"]["loss_fused"], l["hot_path"]["ms_per_step"])
PY
