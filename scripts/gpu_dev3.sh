#!/bin/bash
set -x
mkdir -p gpurun_out
VLSA_AGG_VARIANT=tc timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_tc.log
grep -v watchdog gpurun_out/pytest_gpu_tc.log | tail -4
timeout 900 python scripts/dev_variants.py > gpurun_out/variants.log 2>&1
grep -v watchdog gpurun_out/variants.log | tail -40
