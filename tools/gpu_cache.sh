#!/bin/bash
# call cache (one walk over the GP / MLP table per collocation point instead of two): parity tests + configs 5 and 3
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_configs_gpu.py tests/test_gp_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3
i=0
for v in "--config 5 --steps 60" "--config 3 --steps 60"; do
  timeout -k 10 600 python bench.py --no-cpu-baseline --no-default-options --no-gp $v > gpurun_out/bench_k$i.json 2> gpurun_out/bench_k$i.err
  echo "variant $i [$v] rc=$?"
  i=$((i+1))
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_k?.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d['config']
        print(f, 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms/step %.2f'%d['ms_per_step'], 'it %.2f max %d'%(c['ipm_iterations_mean'], c['ipm_iterations_max']), 'launched %.1f'%c['ipm_iterations_launched_per_step_mean'],
              'opt %.3f acc %.3f'%(c['status_optimal_frac'], c['status_acceptable_frac']), {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()})
        for row in d['extras'].get('kernel_us_by_iteration', [])[:1]: print('     ', row)
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-400:])
PY
