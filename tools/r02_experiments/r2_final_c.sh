#!/bin/bash
# evidence of record for the final tree: tools/final_evidence.sh + smoke + nthll bench
bash tools/final_evidence.sh r02_final
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -2 gpurun_out/r02_final_smoke.log
timeout 600 python tools/bench_nthll.py --steps 5 > gpurun_out/r02_nthll_bench.json 2> gpurun_out/r02_nthll_bench.err; cut -c83-200 gpurun_out/r02_nthll_bench.json
timeout 600 python tools/bench_nthll.py --reads 40000000 --k 32 --steps 3 --cpu-reads 500000 > gpurun_out/r02_nthll_bench_40m.json 2>> gpurun_out/r02_nthll_bench.err; cut -c83-200 gpurun_out/r02_nthll_bench_40m.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 30 -c 60 --csv --log-file gpurun_out/r02_nthll_launches.csv python tools/bench_nthll.py --reads 10000000 --k 32 --steps 1 --cpu-reads 1000 > /dev/null 2>&1; wc -l gpurun_out/r02_nthll_launches.csv
