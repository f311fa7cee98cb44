#!/bin/bash
mkdir -p gpurun_out
for h in exact gauss-newton; do
  timeout 300 python bench.py --no-cpu-baseline --no-gp --steps 10 --hessian $h > gpurun_out/bench_h_$h.json 2> gpurun_out/bench_h_$h.err
  timeout 300 python bench.py --no-cpu-baseline --no-gp --steps 5 --hessian $h --warm-start 0 > gpurun_out/bench_hc_$h.json 2> gpurun_out/bench_hc_$h.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_h*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d['config']
        print(f, 'value %.0f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'iters mean %.2f max %d launched %s ok %.3f'%(c['ipm_iterations_mean'], c['ipm_iterations_max_last_step'], c['ipm_iterations_launched_per_step'][-3:], c['status_optimal_frac']), {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()})
    except Exception as e:
        print(f,'ERR',e)
PY
