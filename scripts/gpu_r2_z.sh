#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "pointwise or encoder or config2 or decoder or bf16" > gpurun_out/pytest_z.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_z.log
timeout 600 python bench.py --no-cpu --no-train --no-extra > gpurun_out/bench_z.log 2> gpurun_out/bench_z.err; echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_z.log').read().strip().splitlines()[-1])
print(l["ms_per_step"], "pointwise", l["kernels"]["pointwise"], "conv2", l["kernels"]["decoder.conv2"], "tf32bb", l["tf32_backbone"]["ms_per_step"], "bf16", l["bf16_mode"]["ms_per_step"])
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python scripts/ncu_step.py 2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
