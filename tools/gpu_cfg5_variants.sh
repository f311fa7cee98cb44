#!/bin/bash
# tuning runs for config 5 (path following, MLP dynamics): evaluation kernel family, warm-start barrier parameter,
# convexification gate
mkdir -p gpurun_out
i=0
for v in "" "--opt b200.eval_variant=0" "--opt b200.regularization_mu=0" "--warm-mu 1e-4" "--warm-mu 1e-8" "--opt b200.eval_variant=0 --opt b200.step_variant=1"; do
  timeout -k 10 600 python bench.py --config 5 --batch 1024 --steps 10 --no-cpu-baseline --no-default-options $v > gpurun_out/bench_v$i.json 2> gpurun_out/bench_v$i.err
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
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-400:])
PY
