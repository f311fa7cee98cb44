#!/bin/bash
# usage: tools/gpu_defines.sh "<nvcc defines>" — parity test of the evaluation kernel + a short config-2 bench with
# modules built under HB_NVCC_DEFINES (separately hashed; prebuild them here with the same variable first)
export HB_NVCC_DEFINES="$1"
timeout -k 5 300 python -m pytest tests/test_nmpc_gpu.py -m gpu -q --timeout 200 -k "quad_eval or full_size_properties or cstr_batch" 2>&1 | tail -3
timeout -k 5 200 python bench.py --steps 20 --no-cpu-baseline --no-default-options --no-gp 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], {k: round(v['avg_us'], 1) for k, v in d['kernels'].items()}, d.get('roofline'))
"
