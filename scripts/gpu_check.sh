#!/bin/bash
# One gpurun call: tcgen05 probe -> parity tests (non-tensor-core first, then tensor-core) -> smoke -> bench -> ncu launch list.
# Every stage runs under its own timeout and logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1; echo "tc_probe rc=$?"
tail -4 gpurun_out/tc_probe.log | cut -c1-600
TC='tc or head or full_model or mvit or full_size'
timeout 900 python -m pytest tests -m gpu -q --durations=6 -k "not ($TC)" > gpurun_out/pytest_simt.log 2>&1; echo "pytest(simt) rc=$?"
tail -25 gpurun_out/pytest_simt.log
timeout 900 python -m pytest tests -m gpu -q --durations=6 -k "$TC" > gpurun_out/pytest_tc.log 2>&1; echo "pytest(tc) rc=$?"
tail -25 gpurun_out/pytest_tc.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -2 gpurun_out/bench.log
if [ "$1" = "ncu" ]; then
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_short.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
fi
