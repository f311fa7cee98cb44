#!/bin/bash
# round-2 GPU check: smoke, GPU tests (per-test timeouts), default bench; logs under gpurun_out/
mkdir -p gpurun_out
timeout -k 10 180 python -u -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout -k 10 900 python -u -m pytest tests -m gpu -q --timeout -k 10 300 --deselect tests/test_gp_gpu.py ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
timeout -k 10 300 python -u -m pytest tests/test_gp_gpu.py -m gpu -q --timeout -k 10 60 > gpurun_out/pytest_gp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gp.log; tail -3 gpurun_out/pytest_gp.log
timeout -k 10 400 python bench.py ${BENCH_ARGS} > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -3 gpurun_out/bench_r2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r2.json').read().strip().splitlines()[-1])
    print('value %.0f'%d['value'], 'e2e %.0f'%d['e2e']['value'], 'ms/step %.3f'%d['ms_per_step'], {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()}, d['roofline']['frac'], d['roofline'].get('riccati_sweep',{}).get('frac'))
except Exception as e:
    print('ERR', e)
PY
