#!/bin/bash
# nthll: hit kernel prefetches its next unit's mask rows; growth factor of the pre-filter chunks
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nthll.py -m gpu -x -q > gpurun_out/r2_e25_pytest.log 2>&1; tail -2 gpurun_out/r2_e25_pytest.log
for g in 100 200 400 50; do
  NTC_HLL_GROW=$g timeout 300 python tools/bench_nthll.py --k 32 --steps 5 --cpu-reads 1000 > gpurun_out/r2_e25_g$g.json 2> gpurun_out/r2_e25_g$g.err
  echo "grow=$g: $(cut -c1-190 gpurun_out/r2_e25_g$g.json)"
done
NTC_HLL_GROW=100 timeout 300 python tools/bench_nthll.py --reads 40000000 --k 32 --steps 3 --cpu-reads 1000 | cut -c1-190
NTC_HLL_GROW=200 timeout 300 python tools/bench_nthll.py --reads 40000000 --k 32 --steps 3 --cpu-reads 1000 | cut -c1-190
