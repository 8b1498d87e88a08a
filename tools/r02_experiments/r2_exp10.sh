#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/check_multi_gpu.py > gpurun_out/r2_e10_check_n8.txt 2>&1; tail -1 gpurun_out/r2_e10_check_n8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 100 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e10_bench_n8.json 2> gpurun_out/r2_e10_bench_n8.err; tail -2 gpurun_out/r2_e10_bench_n8.err
NTC_SLICE_SHIFT=23 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 100 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e10_bench_n8_s23.json 2> gpurun_out/r2_e10_bench_n8_s23.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 100 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e10_bench_n4.json 2> gpurun_out/r2_e10_bench_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 100 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e10_bench_n2.json 2> gpurun_out/r2_e10_bench_n2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_e10_bench_*.json')):
    try:
        for line in open(f):
            if line.startswith('{'):
                d=json.loads(line); r=d['roofline']
                print(f,'n',d['n_gpus'],'value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'stages',r['stages_ms'],'parity',d.get('parity_check'))
    except Exception as e: print(f,'failed',e)
PY
