#!/bin/bash
# nthll: mixed-length test, first-chunk size sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nthll.py -m gpu -x -q > gpurun_out/r2_e22_pytest.log 2>&1; tail -3 gpurun_out/r2_e22_pytest.log
for f in 32 64 128 256 512; do
  NTC_HLL_FIRST_K=$f timeout 300 python tools/bench_nthll.py --k 32 --steps 5 --cpu-reads 1000 > gpurun_out/r2_e22_first$f.json 2> gpurun_out/r2_e22_first$f.err
  echo "first=$f K: $(cut -c1-190 gpurun_out/r2_e22_first$f.json)"
done
