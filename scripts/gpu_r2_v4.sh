#!/bin/bash
# kernel batch 1 (attention 8 rows/warp single wave, regressor 4 rows/warp, 8-channel resize): full suite + bench line
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_v4.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_v4.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_v4.log 2>&1; tail -2 gpurun_out/smoke_v4.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench_v4.log 2> gpurun_out/bench_v4.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_v4.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_v4.log').read().strip().splitlines()[-1])
print(l["ms_per_step"], l["value"], "hot", l["hot_path"]["ms_per_step"], "e2e", l["e2e"]["ms_per_step"])
print({k:v for k,v in l["kernels"].items() if k in ("encoder_layers_tc","regressor_bins","upsample_concat_nhwc","loss_fused","head_chain","gather_embed")})
print(l.get("train"), l.get("configs"))
PY
