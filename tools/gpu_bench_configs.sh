#!/bin/bash
# bench lines of the BASELINE configurations on one GPU; logs under gpurun_out/
mkdir -p gpurun_out
for c in ${CONFIGS:-2 4 5 3}; do
  timeout -k 10 ${BENCH_TIMEOUT:-900} python bench.py --config $c ${BENCH_ARGS} > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
  echo "config $c rc=$?"; tail -2 gpurun_out/bench_c$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c?.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d['config']
        print(f, d['metric'], 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms/step %.3f'%d['ms_per_step'],
              'it %.2f'%c['ipm_iterations_mean'], 'launched %.2f'%c['ipm_iterations_launched_per_step_mean'], 'opt %.3f acc %.3f'%(c['status_optimal_frac'], c['status_acceptable_frac']),
              {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()}, 'full', {k:round(v,1) for k,v in d['roofline']['full_launch_us'].items()},
              'roof %s %.3f'%(d['roofline']['kernel'], d['roofline']['frac']), 'ric %.3f'%d['roofline'].get('riccati_sweep',{}).get('frac',-1),
              'default', d['extras'].get('library_default_options',{}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
