#!/bin/bash
# GP tests + predict micro-bench (fused kernel), config 5 line with the launched-iterations trace, config 2 with the
# 4-lane evaluation kernel at 2 CTAs per SM
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gp_gpu.py -m gpu -q --timeout 100 2>&1 | tail -3
timeout -k 5 120 python tools/gp_bench.py
timeout -k 10 600 python bench.py --config 5 --no-cpu-baseline --no-default-options --no-gp > gpurun_out/bench_g5.json 2> gpurun_out/bench_g5.err
HB_NVCC_DEFINES="-DHB_EVQ_MINB4=2" timeout -k 10 600 python bench.py --config 2 --steps 20 --no-cpu-baseline --no-default-options --no-gp > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err
python - <<'PY'
import json,glob
for f in ['gpurun_out/bench_g5.json', 'gpurun_out/bench_g2.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d['config']
        print(f, 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms/step %.2f'%d['ms_per_step'], 'it %.2f max %d'%(c['ipm_iterations_mean'], c['ipm_iterations_max']), 'launched %.1f'%c['ipm_iterations_launched_per_step_mean'],
              'opt %.3f acc %.3f'%(c['status_optimal_frac'], c['status_acceptable_frac']), {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()})
        print('   launched by step', c.get('ipm_iterations_launched_every_10th_step'))
        for row in d['extras'].get('kernel_us_by_iteration', [])[:2]: print('     ', row)
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-400:])
PY
