#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 200 python bench.py --no-cpu-baseline --no-gp --steps 20 > gpurun_out/bench_c1_$i.json 2> gpurun_out/bench_c1.err
timeout 200 python bench.py --no-cpu-baseline --no-gp --steps 20 --no-compaction > gpurun_out/bench_c0_$i.json 2> gpurun_out/bench_c0.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c?_?.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms/step %.3f'%d['ms_per_step'], d['config']['ipm_iterations_launched_per_step'][-3:], {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()})
    except Exception as e:
        print(f,'ERR',e)
PY
