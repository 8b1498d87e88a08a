#!/bin/bash
# the driver's scaling run, by hand: bench.py at N = 1, 2, 4, 8 back to back on one 8-GPU box, then BASELINE config 4 on 8 GPUs
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 3 --no-cpu > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
port=29600
for n in 2 4 8; do
  port=$((port+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 200 --warmup 3 --no-cpu > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus 8 --workload config4 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_config4_n8.json 2> gpurun_out/r02_config4_n8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_scale_n*.json'))+['gpurun_out/r02_config4_n8.json']:
    try:
        for line in open(f):
            if line.startswith('{'):
                d=json.loads(line); r=d['roofline']
                print(f,'n',d['n_gpus'],'value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'stages',r['stages_ms'],'parity',d.get('parity_check'),'e2e',d['e2e'] and ('%.3e'%d['e2e']['value'],round(d['e2e']['ms_per_step'],2)))
    except Exception as e: print(f,'failed',e)
PY
