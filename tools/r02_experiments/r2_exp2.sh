#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_e2_pytest.log 2>&1; tail -5 gpurun_out/r2_e2_pytest.log
for dbg in 0 1; do
  NTC_FUSED_DBG=$dbg timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e2_dbg$dbg.json 2> gpurun_out/r2_e2_dbg$dbg.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r2_e2_dbg$dbg.json') if l.startswith('{')][0]
    print('dbg=$dbg ms/step %.3f stages %s' % (d['ms_per_step'], d['roofline']['stages_ms']))
except Exception as e:
    print('dbg=$dbg failed', e); print(open('gpurun_out/r2_e2_dbg$dbg.err').read()[-800:])
PY
done
for wl in multik k64s11; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --workload $wl > gpurun_out/r2_e2_$wl.json 2> gpurun_out/r2_e2_$wl.err
  NTC_FUSED=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --workload $wl > gpurun_out/r2_e2_${wl}_old.json 2> gpurun_out/r2_e2_${wl}_old.err
  python - <<PY
import json
for f in ['gpurun_out/r2_e2_$wl.json','gpurun_out/r2_e2_${wl}_old.json']:
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][0]
        print(f, 'ms/step %.3f value %.3e stages %s' % (d['ms_per_step'], d['value'], d['roofline']['stages_ms']))
    except Exception as e:
        print(f, 'failed', e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 4 -c 1 -o gpurun_out/r2_e2_fused python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e2_ncu.log 2>&1
ls -la gpurun_out/r2_e2_fused.ncu-rep
