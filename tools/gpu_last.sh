#!/bin/bash
# final check of the round: smoke, GPU tests (GP file separately, everything under per-test timeouts), default bench
mkdir -p gpurun_out
timeout 120 python -u -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python -u -m pytest tests -m gpu -q --timeout 100 --deselect tests/test_gp_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -u -m pytest tests/test_gp_gpu.py -m gpu -v --timeout 50 > gpurun_out/pytest_gp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gp.log; tail -14 gpurun_out/pytest_gp.log
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
    print('value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms/step %.3f'%d['ms_per_step'], {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()}, d['roofline']['frac'], d['roofline']['riccati_sweep']['frac'])
except Exception as e:
    print('ERR', e)
PY
