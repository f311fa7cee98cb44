#!/bin/bash
# weak-scaling lines at N GPUs of one box: bash tools/gpu_scale.sh N "2 4 5"
N=${1:-2}; CFGS=${2:-"2 4 5"}
mkdir -p gpurun_out
for c in $CFGS; do
  if [ "$N" = "1" ]; then
    timeout -k 10 900 python bench.py --config $c --gpus 1 --steps ${STEPS:-40} --no-cpu-baseline --no-gp > gpurun_out/scale_c${c}_n$N.json 2> gpurun_out/scale_c${c}_n$N.err
  else
    timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + c)) bench.py --config $c --gpus $N --steps ${STEPS:-40} --no-cpu-baseline --no-gp > gpurun_out/scale_c${c}_n$N.json 2> gpurun_out/scale_c${c}_n$N.err
  fi
  echo "config $c N=$N rc=$?"; tail -1 gpurun_out/scale_c${c}_n$N.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1]); c = d['config']
    print('  value %.0f e2e %.0f ms/step %.3f n_gpus %d it %.2f opt %.3f acc %.3f default %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['n_gpus'], c['ipm_iterations_mean'], c['status_optimal_frac'], c['status_acceptable_frac'], d['extras'].get('library_default_options', {}).get('value')))
except Exception as e:
    print('  ERR', e)
"
done
