#!/bin/bash
# Scan kernel, L2 prefetch variants (NTC_SCAN_PREFETCH = 0 none, 1 half way, 2 one column before the end of the tile):
# CUDA-event stage times from bench.py and DRAM bytes per launch of scan_kernel from ncu.  Run on the GPU box:
#   bash tools/scan_l2_experiment.sh > gpurun_out/scan_l2_experiment.txt
for m in 1 0 2; do  # 2 is the default
	echo "== NTC_SCAN_PREFETCH=$m"
	NTC_SCAN_PREFETCH=$m python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('ms_per_step', d['ms_per_step'], 'stages_ms', d['roofline']['stages_ms'])"
	NTC_SCAN_PREFETCH=$m ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:scan_kernel --launch-skip 3 -c 2 --csv \
		python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
done
