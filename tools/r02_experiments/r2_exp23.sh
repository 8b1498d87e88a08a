#!/bin/bash
# nthll: caller-owned registers cleared by the caller between batches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nthll.py -m gpu -x -q > gpurun_out/r2_e23_pytest.log 2>&1; tail -5 gpurun_out/r2_e23_pytest.log
