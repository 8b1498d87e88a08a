#!/bin/bash
# decoupled stagers (coordinator warp + writer warps) in the apply kernels: parity, then A/B against version 1
mkdir -p gpurun_out
NTC_APPLY_V2=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "golden or pipeline or peer_memory or exchange or at_size or uniform_stride or many_batches" > gpurun_out/r2_e16_pytest.log 2>&1; tail -3 gpurun_out/r2_e16_pytest.log
for v in 1 0; do
  NTC_APPLY_V2=$v timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e16_bench_v$v.json 2>/dev/null
done
NTC_APPLY_V2=1 NTC_SLICE_SHIFT=22 timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e16_bench_v1_s22.json 2>/dev/null
NTC_APPLY_V2=1 NTC_APPLY_AHEAD=3 timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e16_bench_v1_a3.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_e16_bench_*.json')):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][0]; r=d['roofline']
        print(f,'value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'stages',r['stages_ms'])
    except Exception as e: print(f,'failed',e)
PY
