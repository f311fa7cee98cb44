#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-gp > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 300 gpurun_out/bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/multigpu_check.py > gpurun_out/multigpu_check.log 2>&1
tail -3 gpurun_out/multigpu_check.log
