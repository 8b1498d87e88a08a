#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> gpurun_out/r02_sanitizer_raw.txt
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_case.py >> gpurun_out/r02_sanitizer_raw.txt 2>&1
  tail -4 gpurun_out/r02_sanitizer_raw.txt
done
