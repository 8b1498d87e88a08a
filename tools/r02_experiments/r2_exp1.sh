#!/bin/bash
# round-2 experiment 1: decomposition of the fused kernel's time (NTC_FUSED_DBG) + ncu capture
mkdir -p gpurun_out
for dbg in 0 1 3; do
  NTC_FUSED_DBG=$dbg timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e1_dbg$dbg.json 2> gpurun_out/r2_e1_dbg$dbg.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r2_e1_dbg$dbg.json') if l.startswith('{')][0]
    print('dbg=$dbg ms/step %.3f stages %s' % (d['ms_per_step'], d['roofline']['stages_ms']))
except Exception as e:
    print('dbg=$dbg failed', e); print(open('gpurun_out/r2_e1_dbg$dbg.err').read()[-800:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 4 -c 1 -o gpurun_out/r2_e1_fused python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e1_ncu.log 2>&1
tail -3 gpurun_out/r2_e1_ncu.log
ls -la gpurun_out/*.ncu-rep
