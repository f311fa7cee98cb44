#!/bin/bash
# full GPU test-suite, then riccati2 with 2 vs 3 CTAs per SM (configs 4, 5)
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5
rm -f gpurun_out/bench_q?.json
i=0
for d in "" "-DHB_RIC2_MINB=3"; do
  for v in "--config 4" "--config 5"; do
    HB_NVCC_DEFINES="$d" timeout -k 10 600 python bench.py --steps 20 --no-cpu-baseline --no-default-options --no-gp $v > gpurun_out/bench_q$i.json 2> gpurun_out/bench_q$i.err
    echo "variant $i [$d $v] rc=$?"
    i=$((i+1))
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_q?.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d['config']
        print(f, 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms/step %.2f'%d['ms_per_step'], 'it %.2f max %d'%(c['ipm_iterations_mean'], c['ipm_iterations_max']), 'launched %.1f'%c['ipm_iterations_launched_per_step_mean'],
              'opt %.3f acc %.3f'%(c['status_optimal_frac'], c['status_acceptable_frac']), {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()})
        for row in d['extras'].get('kernel_us_by_iteration', [])[:1]: print('     ', row)
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-400:])
PY
