#!/bin/bash
# round-2 profile run (ONE GPU): full GPU test-suite, estimator micro-bench, ncu launch list of the default bench
# command, ncu --set full captures of the hot kernels (config 2) and of the config 4 / 5 evaluation kernels
mkdir -p gpurun_out
timeout -k 10 300 python -u -m pytest tests/test_lmpc_condensed.py -m gpu -q --timeout -k 10 300 > gpurun_out/pytest_lmpc.log 2>&1; grep -E '^E ' gpurun_out/pytest_lmpc.log | head -12; tail -2 gpurun_out/pytest_lmpc.log
timeout -k 10 200 python tools/est_bench.py > gpurun_out/est_bench.json 2> gpurun_out/est_bench.err; tail -1 gpurun_out/est_bench.json | cut -c1-600
B="--no-cpu-baseline --no-gp --no-default-options --prewarm-ms 50 --clock-sampler none"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 4 --warmup 3 $B > gpurun_out/launches_bench.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:eval_quad -s 12 -c 2 -o gpurun_out/evq_r2 -f python bench.py --steps 2 --warmup 3 $B > gpurun_out/ncu_evq.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:riccati6 -s 12 -c 2 -o gpurun_out/ric_r2 -f python bench.py --steps 2 --warmup 3 $B > gpurun_out/ncu_ric.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:step_quad -s 12 -c 2 -o gpurun_out/stq_r2 -f python bench.py --steps 2 --warmup 3 $B > gpurun_out/ncu_stq.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:eval_quad -s 12 -c 1 -o gpurun_out/evq_c4_r2 -f python bench.py --config 4 --steps 2 --warmup 3 $B > gpurun_out/ncu_evq4.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:riccati_tma2 -s 12 -c 1 -o gpurun_out/ric_c4_r2 -f python bench.py --config 4 --steps 2 --warmup 3 $B > gpurun_out/ncu_ric4.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:eval_quad -s 40 -c 1 -o gpurun_out/evq_c5_r2 -f python bench.py --config 5 --batch 1024 --steps 2 --warmup 3 $B > gpurun_out/ncu_evq5.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:UkfStep -s 2 -c 1 -o gpurun_out/ukf_r2 -f python tools/est_bench.py 4096 > gpurun_out/ncu_ukf.log 2>&1
for r in evq_r2 ric_r2 stq_r2 evq_c4_r2 ric_c4_r2 evq_c5_r2 ukf_r2; do
  [ -f gpurun_out/$r.ncu-rep ] && python tools/ncu_summary.py gpurun_out/$r.ncu-rep gpurun_out/ncu_${r}_summary.csv "round 2" > /dev/null 2>&1
done
rm -f gpurun_out/evq_c4_r2.ncu-rep gpurun_out/ric_c4_r2.ncu-rep gpurun_out/evq_c5_r2.ncu-rep gpurun_out/ukf_r2.ncu-rep gpurun_out/stq_r2.ncu-rep
ls -la gpurun_out/ | head -40; du -sh gpurun_out
