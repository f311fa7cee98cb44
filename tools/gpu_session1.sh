#!/bin/bash
# one gpurun call: microbench, GPU tests, bench per riccati variant, ncu of the new sweep
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu && /tmp/microbench > gpurun_out/microbench.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err
for v in 3 4 5 6; do
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
timeout 600 ncu --set full --clock-control none --import-source on -k regex:riccati3 -s 6 -c 2 -o gpurun_out/ric3 -f python bench.py --no-cpu-baseline --no-gp --steps 2 --warmup 1 --riccati-variant 3 > gpurun_out/ncu_ric3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:riccati3 -s 6 -c 1 -o gpurun_out/ric5 -f python bench.py --no-cpu-baseline --no-gp --steps 2 --warmup 1 --riccati-variant 5 > gpurun_out/ncu_ric5.log 2>&1
ls -la gpurun_out | tail -20
