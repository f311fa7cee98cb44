#!/bin/bash
# round-2 final measurements on ONE GPU: bench lines of all configs (with CPU baseline), reference arms, ncu launch list
# of the default command, ncu --set full of the config 4 / 5 evaluation and step kernels (summarised on the box)
mkdir -p gpurun_out
for c in 2 4 5 3; do
  timeout -k 10 900 python bench.py --config $c > gpurun_out/bench_r2_c$c.json 2> gpurun_out/bench_r2_c$c.err
  echo "config $c rc=$?"; tail -c 300 gpurun_out/bench_r2_c$c.json | tail -1
done
for c in 2 4 5; do
  timeout -k 10 600 python bench.py --impl reference --config $c --steps 2 --warmup 1 > gpurun_out/bench_r2_reference_c$c.json 2> gpurun_out/bench_r2_reference_c$c.err
  echo "reference $c rc=$?"; tail -c 400 gpurun_out/bench_r2_reference_c$c.json | tail -1
done
B="--no-cpu-baseline --no-gp --no-default-options --prewarm-ms 50 --clock-sampler none"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 4 --warmup 3 $B > gpurun_out/launches_bench.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:eval_quad -s 12 -c 1 -o gpurun_out/evq_c5_r2 -f python bench.py --config 5 --steps 2 --warmup 3 $B > gpurun_out/ncu_evq5.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:step_quad -s 12 -c 1 -o gpurun_out/stq_c5_r2 -f python bench.py --config 5 --steps 2 --warmup 3 $B > gpurun_out/ncu_stq5.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:eval_quad -s 12 -c 1 -o gpurun_out/evq_c4_r2 -f python bench.py --config 4 --steps 2 --warmup 3 $B > gpurun_out/ncu_evq4.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:step_quad -s 12 -c 1 -o gpurun_out/stq_c4_r2 -f python bench.py --config 4 --steps 2 --warmup 3 $B > gpurun_out/ncu_stq4.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gp_predict_fused -c 1 -o gpurun_out/gp_fused_r2 -f python tools/gp_bench.py > gpurun_out/ncu_gp.log 2>&1
for r in evq_c5_r2 stq_c5_r2 evq_c4_r2 stq_c4_r2 gp_fused_r2; do
  [ -f gpurun_out/$r.ncu-rep ] && python tools/ncu_summary.py gpurun_out/$r.ncu-rep gpurun_out/ncu_${r}_summary.csv "round 2 final" > /dev/null 2>&1
  rm -f gpurun_out/$r.ncu-rep
done
ls -la gpurun_out/ | grep -E "r2|launches" | head -40; du -sh gpurun_out
