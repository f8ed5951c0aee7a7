#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|passed|failed" gpurun_out/pytest_all.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 300 python scripts/profile_head.py 3 > gpurun_out/profile_head_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none \
   -k regex:'depth_losses_kernel|gather_embed_nhwc_kernel|split_bf16|nchw_to_nhwc' -s 4 -c 8 --csv --log-file gpurun_out/small_kernels.csv \
   python scripts/profile_head.py 2 > gpurun_out/ncu_small.log 2>&1
echo "ncu rc=$?"; grep -E "depth_losses|gather_embed_nhwc|nchw_to_nhwc" gpurun_out/small_kernels.csv | awk -F'","' '{print substr($5,1,50), $(NF-2), $NF}' | head -20
