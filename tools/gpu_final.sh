#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 600 gpurun_out/bench_final.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 400 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu-baseline --no-gp --steps 2 --warmup 3 > gpurun_out/launches_bench.log 2>&1
wc -l gpurun_out/launches.csv
