#!/bin/bash
mkdir -p gpurun_out
for s in 22 23; do
NTC_SLICE_SHIFT=$s timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 295$s bench.py --gpus 8 --steps 100 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e13_bench_n8_s$s.json 2> gpurun_out/r2_e13_bench_n8_s$s.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_e13_bench_*.json')):
    try:
        for line in open(f):
            if line.startswith('{'):
                d=json.loads(line); r=d['roofline']
                print(f,'n',d['n_gpus'],'value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'stages',r['stages_ms'],'parity',d.get('parity_check'))
    except Exception as e: print(f,'failed',e)
PY
