#!/bin/bash
bash tools/final_evidence.sh r02_final
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; tail -2 gpurun_out/r02_final_smoke.log
timeout 1500 python bench.py --workload config5 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_config5_bench.json 2> gpurun_out/r02_config5_bench.err; tail -2 gpurun_out/r02_config5_bench.err; cut -c1-300 gpurun_out/r02_config5_bench.json
