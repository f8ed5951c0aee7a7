#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "gather or losses or depth_losses or chamfer or silog or config2" > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|passed|failed" gpurun_out/pytest_f.log | tail -5
timeout 600 python scripts/profile_train.py > gpurun_out/profile_train_tc.log 2>&1; echo "profile tc rc=$?"; head -45 gpurun_out/profile_train_tc.log | cut -c1-170
MDE_TRAIN_CONV=cudnn timeout 600 python scripts/profile_train.py > gpurun_out/profile_train_cudnn.log 2>&1; echo "profile cudnn rc=$?"; head -30 gpurun_out/profile_train_cudnn.log | cut -c1-170
MDE_CUDNN_BENCHMARK=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-train --no-extra > gpurun_out/bench_cudnnbench.log 2>&1; echo "bench(cudnn.benchmark) rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_cudnnbench.log').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','tf32_backbone','hot_path')}); print(l['kernels'])
PY
timeout 300 python scripts/profile_head.py 3 > gpurun_out/profile_head_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none \
   -k regex:'depth_losses_kernel|gather_embed_nhwc_kernel' -s 2 -c 4 --csv --log-file gpurun_out/small_kernels.csv \
   python scripts/profile_head.py 2 > gpurun_out/ncu_small.log 2>&1
echo "ncu rc=$?"; grep -E "depth_losses|gather_embed_nhwc" gpurun_out/small_kernels.csv | awk -F'","' '{print substr($5,1,50), $(NF-2), $NF}' | head -8
