#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -u -m pytest tests -m gpu -q --timeout -k 10 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; grep -E "^E  |FAILED|passed|failed" gpurun_out/pytest_gpu.log | head -30
timeout -k 10 200 python tools/est_bench.py > gpurun_out/est_bench.json 2> gpurun_out/est_bench.err; tail -1 gpurun_out/est_bench.json | cut -c1-900; tail -3 gpurun_out/est_bench.err
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:ukf_lanes -s 2 -c 1 -o gpurun_out/ukf_r2 -f python tools/est_bench.py 16384 > gpurun_out/ncu_ukf.log 2>&1
python tools/ncu_summary.py gpurun_out/ukf_r2.ncu-rep gpurun_out/ncu_ukf_lanes_r2_summary.csv "round 2, B=16384" > /dev/null 2>&1; rm -f gpurun_out/ukf_r2.ncu-rep
head -20 gpurun_out/ncu_ukf_lanes_r2_summary.csv
