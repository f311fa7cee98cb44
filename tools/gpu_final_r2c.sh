#!/bin/bash
# last run of round 2: full GPU test-suite on the final tree, final lines of configs 5 and 3 (call cache), smoke
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
for c in 5 3; do
  timeout -k 10 600 python bench.py --config $c > gpurun_out/bench_r2_c$c.json 2> gpurun_out/bench_r2_c$c.err
  echo "config $c rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_c$c.json').read().strip().splitlines()[-1]); c=d['config']
print('  value %.0f e2e %.0f ms/step %.3f it %.2f launched %.2f opt %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step'], c['ipm_iterations_mean'], c['ipm_iterations_launched_per_step_mean'], c['status_optimal_frac']), {k:round(v,1) for k,v in d['roofline']['full_launch_us'].items()}, 'frac %.3f'%d['roofline']['frac'], 'default', d['extras'].get('library_default_options',{}).get('value'), 'cpu', d['cpu_baseline']['value'])
PY
done
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
