#!/bin/bash
# round-2 final re-measure after the last kernel changes: config 5 (150-step window), config 2 / 3 (GP line), reference arms
mkdir -p gpurun_out
for c in 5 2 3 4; do
  timeout -k 10 900 python bench.py --config $c > gpurun_out/bench_r2_c$c.json 2> gpurun_out/bench_r2_c$c.err
  echo "config $c rc=$?"; tail -c 200 gpurun_out/bench_r2_c$c.json | tail -1
done
for c in 2 4 5; do
  timeout -k 10 600 python bench.py --impl reference --config $c --steps 2 --warmup 1 > gpurun_out/bench_r2_reference_c$c.json 2> gpurun_out/bench_r2_reference_c$c.err
  echo "reference $c rc=$?"; tail -c 300 gpurun_out/bench_r2_reference_c$c.json | tail -1
done
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
