#!/bin/bash
# tuning runs for config 5 (path following, MLP dynamics): warm-start shift and barrier parameter; plus a short config-4
# and config-2 line (per-iteration kernel table)
mkdir -p gpurun_out
i=0
for v in "--config 5 --batch 1024" "--config 5 --batch 1024 --opt b200.shift_warm_start=1" "--config 5 --batch 1024 --opt b200.shift_warm_start=1 --warm-mu 1e-9" "--config 5 --batch 1024 --warm-mu 1e-9" "--config 5 --opt b200.shift_warm_start=1 --warm-mu 1e-9" "--config 4" "--config 2"; do
  timeout -k 10 600 python bench.py --steps 20 --no-cpu-baseline --no-default-options --no-gp $v > gpurun_out/bench_v$i.json 2> gpurun_out/bench_v$i.err
  echo "variant $i [$v] rc=$?"
  i=$((i+1))
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_v?.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d['config']
        print(f, 'value %.0f'%d['value'], 'ms/step %.2f'%d['ms_per_step'], 'it %.2f max %d'%(c['ipm_iterations_mean'], c['ipm_iterations_max']), 'launched %.1f'%c['ipm_iterations_launched_per_step_mean'],
              'opt %.3f acc %.3f'%(c['status_optimal_frac'], c['status_acceptable_frac']), {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()})
        for row in d['extras'].get('kernel_us_by_iteration', [])[:8]: print('     ', row)
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-400:])
PY
