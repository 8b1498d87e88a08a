#!/bin/bash
# memcheck of every kernel on the final tree (round-1 + round-2 cases)
mkdir -p gpurun_out
echo "== memcheck, final tree (tools/sanitize_case.py: all cases)" > gpurun_out/r02_sanitizer_final.txt
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_case.py >> gpurun_out/r02_sanitizer_final.txt 2>&1
tail -12 gpurun_out/r02_sanitizer_final.txt
