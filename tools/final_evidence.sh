#!/bin/bash
# Round evidence in one GPU call: GPU test suite, bench line, ncu launch list of the same command, ncu --set full of the
# pipeline's three kernels.  Outputs under gpurun_out/ (copy what is cited into profiles/).
#   bash tools/final_evidence.sh r01_final
tag=${1:-final}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-400 gpurun_out/${tag}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
	python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1; wc -l gpurun_out/${tag}_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:scan_kernel|hit_kernel|apply_kernel' --launch-skip 8 -c 4 -f \
	-o gpurun_out/prof_${tag} python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log | cut -c1-200
