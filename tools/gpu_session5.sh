#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"eval_quad|step_quad" -s 8 -c 2 -o gpurun_out/evst -f python bench.py --no-cpu-baseline --no-gp --steps 2 --warmup 1 --warm-mu 1e-4 > gpurun_out/ncu_evst.log 2>&1
tail -c 300 gpurun_out/ncu_evst.log
