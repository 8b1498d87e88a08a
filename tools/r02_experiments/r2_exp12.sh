#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nthll.py -m gpu -x -q > gpurun_out/r2_e12_pytest.log 2>&1; tail -3 gpurun_out/r2_e12_pytest.log
timeout 600 python tools/bench_nthll.py --steps 5 > gpurun_out/r2_e12_nthll.json 2> gpurun_out/r2_e12_nthll.err; tail -2 gpurun_out/r2_e12_nthll.err; cut -c1-330 gpurun_out/r2_e12_nthll.json
timeout 600 python tools/bench_nthll.py --reads 40000000 --k 32 --steps 3 --cpu-reads 500000 > gpurun_out/r2_e12_nthll40m.json 2> gpurun_out/r2_e12_nthll40m.err; cut -c1-330 gpurun_out/r2_e12_nthll40m.json
