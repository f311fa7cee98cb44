#!/bin/bash
mkdir -p gpurun_out
{
for f in tools/_bin/ric_*; do echo -n "$f  "; timeout 60 $f 4096 | tail -1; done
} > gpurun_out/ric_bench.txt 2>&1
cat gpurun_out/ric_bench.txt
