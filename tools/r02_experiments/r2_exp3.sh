#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "peer_memory or exchange or pool_exhaustion or many_batches" > gpurun_out/r2_e3_pytest.log 2>&1; tail -5 gpurun_out/r2_e3_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py > gpurun_out/r2_e3_check_n2.txt 2>&1; tail -4 gpurun_out/r2_e3_check_n2.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 3 --no-cpu > gpurun_out/r2_e3_bench_n2.json 2> gpurun_out/r2_e3_bench_n2.err; tail -c 1200 gpurun_out/r2_e3_bench_n2.json; tail -5 gpurun_out/r2_e3_bench_n2.err
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e3_bench_n1.json 2> gpurun_out/r2_e3_bench_n1.err; tail -c 600 gpurun_out/r2_e3_bench_n1.json
