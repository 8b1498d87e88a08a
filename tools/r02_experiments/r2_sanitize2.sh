#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool (round-2 kernels only, nthll pre-filter taken)" >> gpurun_out/r02_sanitizer_raw2.txt
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_case.py --round2-only >> gpurun_out/r02_sanitizer_raw2.txt 2>&1
  tail -3 gpurun_out/r02_sanitizer_raw2.txt
done
