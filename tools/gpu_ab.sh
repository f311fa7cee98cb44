#!/bin/bash
# A/B of two builds of the bench module selected by HB_NVCC_DEFINES (both prebuilt in _gen/)
mkdir -p gpurun_out
for tag in A B; do
  if [ $tag = A ]; then export HB_NVCC_DEFINES=""; else export HB_NVCC_DEFINES="$1"; fi
  timeout 300 python bench.py --no-cpu-baseline --no-gp --steps 10 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
done
python - <<'PY'
import json
for f in ['gpurun_out/bench_A.json','gpurun_out/bench_B.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms/step %.3f'%d['ms_per_step'], {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()})
    except Exception as e:
        print(f,'ERR',e, open(f.replace('json','err')).read()[-500:])
PY
