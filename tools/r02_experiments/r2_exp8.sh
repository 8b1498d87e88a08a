#!/bin/bash
# 8-GPU box: config 2 scaling again (shift 22), config 4 (1 B reads over 8 GPUs), single-GPU extras on GPU 0
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 100 --warmup 3 --no-cpu > gpurun_out/r2_e8_bench_n8.json 2> gpurun_out/r2_e8_bench_n8.err; tail -2 gpurun_out/r2_e8_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --workload config4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_e8_config4_n8.json 2> gpurun_out/r2_e8_config4_n8.err; tail -2 gpurun_out/r2_e8_config4_n8.err
python - <<'PY'
import json
for f in ['gpurun_out/r2_e8_bench_n8.json','gpurun_out/r2_e8_config4_n8.json']:
    try:
        for line in open(f):
            if line.startswith('{'):
                d=json.loads(line); r=d['roofline']
                print(f,'n',d['n_gpus'],'value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'stages',r['stages_ms'],'parity',d.get('parity_check'),'red',d.get('reductions'),'e2e',d['e2e'] and (d['e2e']['value'],d['e2e']['ms_per_step']))
    except Exception as e: print(f,'failed',e)
PY
