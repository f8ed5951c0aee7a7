#!/bin/bash
# final single-GPU evidence run of the round: full GPU suite, smoke, default bench (+ cpu baseline), configs 3/4/5, reference arm,
# launch list, ncu --set full summaries (raw pages as csv; the reports themselves are dropped to stay under the transfer limit)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
timeout 900 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -c 1200 gpurun_out/bench_final.log
for c in 3 4 5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_final_cfg$c.log 2> gpurun_out/bench_final_cfg$c.err; echo "bench cfg$c rc=$?"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.log 2> gpurun_out/bench_final_reference.err; echo "reference arm rc=$?"; tail -c 400 gpurun_out/bench_final_reference.log
timeout 300 python scripts/ncu_step.py 2 > gpurun_out/step_final.log 2>&1; tail -1 gpurun_out/step_final.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python scripts/ncu_step.py 2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
i=0
for spec in "head_chain|depth_losses|gather_embed_dense|patch_embed|upsample_nhwc:0:9" "conv3x3_kernel:0:10" "pointwise_x3|bias_act_pool:0:8" "pointwise_x3:36:4"; do
  i=$((i+1)); k=${spec%%:*}; rest=${spec#*:}; sk=${rest%%:*}; c=${rest##*:}
  src=""; if [ $i -eq 4 ]; then src="--import-source on"; fi
  timeout 600 ncu --profile-from-start off --set full --clock-control none $src -s $sk -c $c -k regex:"$k" -o gpurun_out/prof_final_$i -f \
      python scripts/ncu_step.py 2 > gpurun_out/ncu_full_$i.log 2>&1
  echo "ncu full $i rc=$?"
  ncu -i gpurun_out/prof_final_$i.ncu-rep --page raw --csv > gpurun_out/prof_final_${i}_raw.csv 2>/dev/null
  if [ $i -eq 4 ]; then ncu -i gpurun_out/prof_final_$i.ncu-rep --page source --csv > gpurun_out/prof_final_${i}_source.csv 2>/dev/null; fi
  rm -f gpurun_out/prof_final_$i.ncu-rep
done
ls -la gpurun_out | head -40
