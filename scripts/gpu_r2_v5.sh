#!/bin/bash
# programmatic dependent launch A/B + fused regressor launch: suite with PDL on, bench with PDL off / on
mkdir -p gpurun_out
MDE_PDL=1 timeout 700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_v5.log 2>&1; echo "pytest(PDL=1) rc=$?"; tail -4 gpurun_out/pytest_v5.log
for p in 0 1; do
  MDE_PDL=$p timeout 400 python bench.py --no-cpu --no-train --no-extra > gpurun_out/bench_v5_pdl$p.log 2> gpurun_out/bench_v5_pdl$p.err; echo "bench PDL=$p rc=$?"; tail -3 gpurun_out/bench_v5_pdl$p.err
  python - <<PY
import json
l=json.loads(open('gpurun_out/bench_v5_pdl$p.log').read().strip().splitlines()[-1])
print("PDL=$p", l["ms_per_step"], "eager", l["eager"]["ms_per_step"], "hot", l["hot_path"]["ms_per_step"], "e2e", l["e2e"]["ms_per_step"], l["config"]["launch"], l.get("graph_error"))
print({k:v for k,v in l["kernels"].items() if k in ("encoder_layers_tc","regressor_bins","upsample_concat_nhwc","fold_queries","patch_embed")})
PY
done
