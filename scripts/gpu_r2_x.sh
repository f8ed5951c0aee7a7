#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "loss or silog or chamfer or config2 or golden or gather" > gpurun_out/pytest_x.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_x.log
timeout 600 python bench.py --no-cpu --no-train --no-extra > gpurun_out/bench_x.log 2> gpurun_out/bench_x.err; echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_x.log').read().strip().splitlines()[-1])
print(l["ms_per_step"], "loss", l["kernels"]["loss_fused"], "gather", l["kernels"]["gather_embed"], "hot", l["hot_path"]["ms_per_step"])
PY
