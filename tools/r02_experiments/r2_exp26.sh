#!/bin/bash
# nthll: growth factor of the pre-filter chunks, larger values, 10 M and 40 M reads
mkdir -p gpurun_out
for g in 400 800 1600; do
  NTC_HLL_GROW=$g timeout 300 python tools/bench_nthll.py --k 32 --steps 5 --cpu-reads 1000 > gpurun_out/r2_e26_g$g.json 2> gpurun_out/r2_e26_g$g.err
  echo "grow=$g: $(cut -c83-190 gpurun_out/r2_e26_g$g.json)"
  NTC_HLL_GROW=$g timeout 300 python tools/bench_nthll.py --reads 40000000 --k 32 --steps 3 --cpu-reads 1000 | cut -c83-190
done
NTC_HLL_GROW=400 timeout 300 python tools/bench_nthll.py --k 64 --steps 5 --cpu-reads 1000 | cut -c83-190
