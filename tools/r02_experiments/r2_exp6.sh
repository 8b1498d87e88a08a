#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_e6_pytest.log 2>&1; tail -4 gpurun_out/r2_e6_pytest.log
NTC_HOST_TIMING=1 timeout 300 python tools/bench_ragged.py --steps 5 > gpurun_out/r2_e6_ragged.json 2> gpurun_out/r2_e6_ragged_hosttiming.txt
cat gpurun_out/r2_e6_ragged_hosttiming.txt | grep "host timing"
python -c "
import json; d=json.loads(open('gpurun_out/r2_e6_ragged.json').readline()); print({k:(round(v['wall_ms_per_pass'],2), round(v['kernel_ms_per_pass'],2), round(v['host_in_submit_ms'],2)) for k,v in d.items() if isinstance(v,dict)})"
for v in "X=0" "NTC_MULTIK_CHUNK_MB=56" "NTC_POOL_BLOCKS=4000000"; do
  env $v timeout 600 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e6_config3_$v.json 2> gpurun_out/r2_e6_config3_$v.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r2_e6_config3_$v.json') if l.startswith('{')][0]; r=d['roofline']
    print('config3 $v value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'stages',r['stages_ms'])
except Exception as e: print('config3 $v failed',e)
PY
done
