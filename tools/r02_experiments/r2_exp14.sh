#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "peer_memory or exchange" > gpurun_out/r2_e14_pytest.log 2>&1; tail -2 gpurun_out/r2_e14_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/check_multi_gpu.py > gpurun_out/r2_e14_check_n8.txt 2>&1; tail -1 gpurun_out/r2_e14_check_n8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 8 --steps 100 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e14_bench_n8.json 2> gpurun_out/r2_e14_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 2 --steps 100 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e14_bench_n2.json 2> gpurun_out/r2_e14_bench_n2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_e14_bench_*.json')):
    try:
        for line in open(f):
            if line.startswith('{'):
                d=json.loads(line); r=d['roofline']
                print(f,'n',d['n_gpus'],'value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'stages',r['stages_ms'],'parity',d.get('parity_check'))
    except Exception as e: print(f,'failed',e)
PY
