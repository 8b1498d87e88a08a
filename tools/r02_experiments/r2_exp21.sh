#!/bin/bash
# nthll after the hll_min_kernel change: parity, bench, launch list of one 10 M step, racecheck + synccheck of the round-2 kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nthll.py -m gpu -x -q > gpurun_out/r2_e21_pytest.log 2>&1; tail -3 gpurun_out/r2_e21_pytest.log
timeout 600 python tools/bench_nthll.py --steps 5 > gpurun_out/r2_e21_nthll.json 2> gpurun_out/r2_e21_nthll.err; tail -2 gpurun_out/r2_e21_nthll.err; cut -c1-330 gpurun_out/r2_e21_nthll.json
timeout 600 python tools/bench_nthll.py --reads 40000000 --k 32 --steps 3 --cpu-reads 500000 > gpurun_out/r2_e21_nthll40m.json 2> gpurun_out/r2_e21_nthll40m.err; cut -c1-330 gpurun_out/r2_e21_nthll40m.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 30 -c 60 --csv --log-file gpurun_out/r2_e21_launches.csv python tools/bench_nthll.py --reads 10000000 --k 32 --steps 1 --cpu-reads 1000 > /dev/null 2>&1
grep -oE "(hll_hit_kernel|scan_kernel<[^>]*>|hll_kernel<[^>]*>|hll_min_kernel)[^\"]*\"(,\"[^\"]*\")*" gpurun_out/r2_e21_launches.csv | awk -F'"' '{print $1, $(NF-1)}' | sed -E 's/\(.*\)//' | head -22
for tool in racecheck synccheck; do
  echo "== $tool (round-2 kernels, nthll pre-filter at the device-read level)" >> gpurun_out/r2_e21_sanitizer.txt
  timeout 500 compute-sanitizer --tool $tool python tools/sanitize_case.py --round2-only >> gpurun_out/r2_e21_sanitizer.txt 2>&1
  tail -2 gpurun_out/r2_e21_sanitizer.txt
done
