#!/bin/bash
# tuning runs: bench modules built with HB_NVCC_DEFINES variants (prebuilt in _gen/), default build first and last
mkdir -p gpurun_out
i=0
for defs in "" "-DHB_RCP_NEWTON_STEPS=1" "-DHB_R6_SKIPFWD=1" "-DHB_STEPQ_MINB=3" "-DHB_RCP_NEWTON_STEPS=1 -DHB_R6_SKIPFWD=1" ""; do
  export HB_NVCC_DEFINES="$defs"
  timeout 120 python bench.py --no-cpu-baseline --no-gp --steps 20 > gpurun_out/bench_x$i.json 2> gpurun_out/bench_x$i.err
  i=$((i+1))
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_x?.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d['config']
        print(f, 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms/step %.3f'%d['ms_per_step'], 'it %.3f'%c['ipm_iterations_mean'], c['ipm_iterations_launched_per_step'][-2:], 'ok %.3f'%c['status_optimal_frac'], {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()})
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-300:])
PY
