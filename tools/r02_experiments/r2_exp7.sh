#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_e7_pytest.log 2>&1; tail -4 gpurun_out/r2_e7_pytest.log
timeout 600 python tools/bench_cli.py 8000000 32 > gpurun_out/r2_e7_cli.txt 2>&1; cat gpurun_out/r2_e7_cli.txt | tail -30
timeout 600 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_e7_config3.json 2> gpurun_out/r2_e7_config3.err
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r2_e7_config3.json') if l.startswith('{')][0]; r=d['roofline']
    print('config3 value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'stages',r['stages_ms'], 'e2e', d['e2e'] and (d['e2e']['value'], d['e2e']['ms_per_step']))
except Exception as e: print('config3 failed',e)
PY
NTC_HOST_TIMING=0 timeout 300 python tools/bench_ragged.py --steps 5 > gpurun_out/r2_e7_ragged.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r2_e7_ragged.json').readline()); print({k:(round(v['wall_ms_per_pass'],2), round(v['kernel_ms_per_pass'],2), round(v['host_in_submit_ms'],2)) for k,v in d.items() if isinstance(v,dict)})"
