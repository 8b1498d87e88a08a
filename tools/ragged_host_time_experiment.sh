#!/bin/bash
# Next-round experiment: why does the host spend ~0.8 ms inside every ntc_submit of the padded ragged path?
# Prints (wall ms, kernel ms, host-in-submit ms) per pass for pipeline / roll64 under each variant.  ~10 s of GPU per line.
#   bash tools/ragged_host_time_experiment.sh > gpurun_out/ragged_host_time.txt
show='import json,sys; d=json.loads(sys.stdin.readline()); print({k:(round(v["wall_ms_per_pass"],2), round(v["kernel_ms_per_pass"],2), round(v["host_in_submit_ms"],2)) for k,v in d.items() if isinstance(v,dict)})'
run() { echo "== $*"; env "$@" python tools/bench_ragged.py --steps 5 $EXTRA | python -c "$show"; }
run NTC_DUMMY=0
NTC_HOST_TIMING=1 python tools/bench_ragged.py --steps 5 2>&1 >/dev/null | grep "host timing"   # per call site, both kernels (two contexts' worth in one table)
run NTC_APPLY_COOP=0
run NTC_CLEAR_MEMSET=0
run CUDA_DEVICE_MAX_CONNECTIONS=32
run CUDA_MODULE_LOADING=EAGER
EXTRA="--per 2000000" run NTC_DUMMY=0
EXTRA="--per 125000" run NTC_DUMMY=0
EXTRA="--roll64-first" run NTC_DUMMY=0
