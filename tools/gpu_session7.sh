#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --no-gp --steps 10 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<'PY'
import json,glob
for f in ['gpurun_out/bench_q.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms/step %.3f'%d['ms_per_step'], {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()})
    except Exception as e:
        print(f,'ERR',e)
PY
