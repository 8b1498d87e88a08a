#!/bin/bash
# Round evidence in one GPU call: GPU test suite, bench line, ncu launch list of the same command, ncu --set full of the
# pipeline's three kernels.  Outputs under gpurun_out/ (copy what is cited into profiles/).
#   bash tools/final_evidence.sh r02_final
tag=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 200 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-400 gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_reference_bench.json 2>> gpurun_out/${tag}_bench.err; cut -c1-300 gpurun_out/${tag}_reference_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
	python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1; wc -l gpurun_out/${tag}_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:scan_kernel|hit_kernel|apply_kernel' --launch-skip 12 -c 4 -f \
	-o gpurun_out/prof_${tag} python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log | cut -c1-200
