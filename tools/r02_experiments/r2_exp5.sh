#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "at_size or peer_memory" > gpurun_out/r2_e5_pytest.log 2>&1; tail -6 gpurun_out/r2_e5_pytest.log
for cfg in "2 23" "3 23" "4 23" "6 23" "2 22" "4 22" "8 22" "2 24"; do
  set -- $cfg
  NTC_APPLY_AHEAD=$1 NTC_SLICE_SHIFT=$2 timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e5_apply_$1_$2.json 2> gpurun_out/r2_e5_apply_$1_$2.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r2_e5_apply_$1_$2.json') if l.startswith('{')][0]
    print('ahead=$1 shift=$2 ms/step %.3f stages %s' % (d['ms_per_step'], d['roofline']['stages_ms']))
except Exception as e:
    print('ahead=$1 shift=$2 failed', e)
PY
done
timeout 600 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_e5_config3.json 2> gpurun_out/r2_e5_config3.err; tail -2 gpurun_out/r2_e5_config3.err
timeout 600 python bench.py --workload config4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_e5_config4_1gpu.json 2> gpurun_out/r2_e5_config4_1gpu.err; tail -2 gpurun_out/r2_e5_config4_1gpu.err
timeout 900 python bench.py --workload long10kN --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_e5_long10kN.json 2> gpurun_out/r2_e5_long10kN.err; tail -2 gpurun_out/r2_e5_long10kN.err
python - <<'PY'
import json
for w in ['config3','config4_1gpu','long10kN']:
    f='gpurun_out/r2_e5_%s.json'%w
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][0]; r=d['roofline']
        print(w,'value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'stages',r['stages_ms'],'pipeline_ms %.3f frac %.4f pfrac %.4f'%(r['pipeline_ms'],r['frac'],r['pipeline_frac']),'e2e',d['e2e'] and (d['e2e']['value'],d['e2e']['ms_per_step']))
    except Exception as e: print(w,'failed',e)
PY
