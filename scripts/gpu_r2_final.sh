#!/bin/bash
# final single-GPU evidence run of the round (budget: ~6 GPU-minutes): full GPU suite, smoke, the default bench line (+ cpu baseline,
# training legs, configs 3 / 5), the launch list of one step, ncu --set full of the kernels that changed last (raw page exported on
# the box; the reports are dropped), the reference arm, config 4
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
timeout 400 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_final.log
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python scripts/ncu_step.py 2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 200 ncu --profile-from-start off --set full --clock-control none \
    -k regex:"attention_kernel|regressor_bins|upsample_nhwc|copy_channels|head_chain|depth_losses|gather_embed_dense" -c 12 \
    -o gpurun_out/prof_final_a -f python scripts/ncu_step.py 2 > gpurun_out/ncu_full_a.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/prof_final_a.ncu-rep --page raw --csv > gpurun_out/prof_final_a_raw.csv 2>/dev/null
rm -f gpurun_out/prof_final_a.ncu-rep
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.log 2> gpurun_out/bench_final_reference.err; echo "reference arm rc=$?"; tail -c 300 gpurun_out/bench_final_reference.log
timeout 150 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_final_cfg4.log 2> gpurun_out/bench_final_cfg4.err; echo "bench cfg4 rc=$?"; tail -c 400 gpurun_out/bench_final_cfg4.log
ls -la gpurun_out | head -30
