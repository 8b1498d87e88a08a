#!/bin/bash
# final tree, two GPUs: tools/check_multi_gpu.py and the bench line at N = 2 (the driver's launch line)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tools/check_multi_gpu.py > gpurun_out/r02_final_check_n2.txt 2>&1; grep "multi-gpu check" gpurun_out/r02_final_check_n2.txt | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 200 --warmup 3 > gpurun_out/r02_final_bench_n2.json 2> gpurun_out/r02_final_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29623 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_final_reference_n2.json 2> gpurun_out/r02_final_reference_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_final_bench_n2.json','gpurun_out/r02_final_reference_n2.json'):
    try:
        for line in open(f):
            if line.startswith('{'):
                d=json.loads(line)
                print(f,'n',d['n_gpus'],'value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'parity',d.get('parity_check'),'e2e',d.get('e2e') and '%.3e'%d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f,'failed',e)
PY
