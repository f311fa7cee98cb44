#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for v in 2 3 5 7 8; do
  timeout 300 python bench.py --no-cpu-baseline --no-gp --steps 10 --riccati-variant $v > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_v*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.0f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()})
    except Exception as e:
        print(f,'ERR',e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:riccati4 -s 6 -c 2 -o gpurun_out/ric4 -f python bench.py --no-cpu-baseline --no-gp --steps 2 --warmup 1 --riccati-variant 7 > gpurun_out/ncu_ric4.log 2>&1
tail -3 gpurun_out/ncu_ric4.log
