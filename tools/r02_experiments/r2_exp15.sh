#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_cli.py -m gpu -x -q -k "headerless or more_threads or pinned or tightly" > gpurun_out/r2_e15_pytest.log 2>&1; tail -3 gpurun_out/r2_e15_pytest.log
timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu > gpurun_out/r2_e15_bench.json 2> gpurun_out/r2_e15_bench.err; tail -2 gpurun_out/r2_e15_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2_e15_bench.json') if l.startswith('{')][0]; r=d['roofline']
print('value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'stages',r['stages_ms'],'e2e',d['e2e'])
PY
