#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_nthll.py tests/test_parity_gpu.py -m gpu -x -q -k "hll or peer_memory or every_k_class or uniform_reads or golden" > gpurun_out/r2_e9_pytest.log 2>&1; tail -5 gpurun_out/r2_e9_pytest.log
timeout 600 python tools/bench_nthll.py --steps 5 > gpurun_out/r2_e9_nthll.json 2> gpurun_out/r2_e9_nthll.err; tail -3 gpurun_out/r2_e9_nthll.err; cat gpurun_out/r2_e9_nthll.json | cut -c1-700
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e9_bench_n1.json 2> gpurun_out/r2_e9_bench_n1.err
for a in 1 2; do NTC_APPLY_AHEAD=$a timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e9_bench_n1_ahead$a.json 2>/dev/null; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_e9_bench_n1*.json')):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][0]; r=d['roofline']
        print(f,'value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'stages',r['stages_ms'])
    except Exception as e: print(f,'failed',e)
PY
