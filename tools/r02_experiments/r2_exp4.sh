#!/bin/bash
# 8-GPU box: topology facts, multi-GPU parity at N=8, scaling bench at N=8 and N=4
mkdir -p gpurun_out
{ nvidia-smi topo -m; echo; lscpu | head -30; echo; cat /sys/devices/system/node/online; grep -i "allowed" /proc/self/status; nproc; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done; free -g | head -3; } > gpurun_out/r2_e4_topology.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/check_multi_gpu.py > gpurun_out/r2_e4_check_n8.txt 2>&1; tail -2 gpurun_out/r2_e4_check_n8.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 50 --warmup 3 --no-cpu > gpurun_out/r2_e4_bench_n8.json 2> gpurun_out/r2_e4_bench_n8.err; tail -3 gpurun_out/r2_e4_bench_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 50 --warmup 3 --no-cpu > gpurun_out/r2_e4_bench_n4.json 2> gpurun_out/r2_e4_bench_n4.err; tail -3 gpurun_out/r2_e4_bench_n4.err
python - <<'PY'
import json
for f in ['gpurun_out/r2_e4_bench_n8.json','gpurun_out/r2_e4_bench_n4.json']:
    try:
        for line in open(f):
            if line.startswith('{'):
                d=json.loads(line); r=d['roofline']
                print(f,'n',d['n_gpus'],'value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']),'stages',r['stages_ms'],'parity',d.get('parity_check'),'red',d.get('reductions'),'e2e ms',d['e2e'] and d['e2e']['ms_per_step'])
    except Exception as e: print(f,'failed',e)
PY
