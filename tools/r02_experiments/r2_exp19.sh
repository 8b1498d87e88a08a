#!/bin/bash
# hll_hit_kernel: eight 16-byte loads in flight per lane, grid = resident CTAs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nthll.py -m gpu -x -q > gpurun_out/r2_e19_pytest.log 2>&1; tail -3 gpurun_out/r2_e19_pytest.log
timeout 600 python tools/bench_nthll.py --steps 5 > gpurun_out/r2_e19_nthll.json 2> gpurun_out/r2_e19_nthll.err; tail -2 gpurun_out/r2_e19_nthll.err; cut -c1-330 gpurun_out/r2_e19_nthll.json
timeout 600 python tools/bench_nthll.py --reads 40000000 --k 32 --steps 3 --cpu-reads 500000 > gpurun_out/r2_e19_nthll40m.json 2> gpurun_out/r2_e19_nthll40m.err; cut -c1-330 gpurun_out/r2_e19_nthll40m.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_e19_launches.csv python tools/bench_nthll.py --reads 40000000 --k 32 --steps 1 --cpu-reads 1000 > /dev/null 2>&1
grep -E "hll|scan" gpurun_out/r2_e19_launches.csv | awk -F'","' '{print $5, $NF}' | tail -12
