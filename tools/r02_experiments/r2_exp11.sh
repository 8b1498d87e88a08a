#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_e11_pytest.log 2>&1; tail -4 gpurun_out/r2_e11_pytest.log
for s in 22 23; do NTC_SLICE_SHIFT=$s timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e11_bench_n1_s$s.json 2>/dev/null; done
timeout 300 python bench.py --workload k64s11 --steps 30 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e11_k64s11.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_e11_*.json')):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][0]; r=d['roofline']
        print(f,'value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'stages',r['stages_ms'])
    except Exception as e: print(f,'failed',e)
PY
timeout 600 python tools/bench_nthll.py --reads 40000000 --k 32 --steps 3 --cpu-reads 500000 > gpurun_out/r2_e11_nthll40m.json 2> gpurun_out/r2_e11_nthll40m.err; cut -c1-400 gpurun_out/r2_e11_nthll40m.json; tail -2 gpurun_out/r2_e11_nthll40m.err
