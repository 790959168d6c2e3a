#!/bin/bash
# Quick dev round: TC-forced parity tests + timings (tc only)
set -x
mkdir -p gpurun_out
VLSA_AGG_VARIANT=tc timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q -x > gpurun_out/pytest_gpu_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_tc.log
grep -v watchdog gpurun_out/pytest_gpu_tc.log | tail -8
timeout 900 python scripts/dev_train_time.py tc > gpurun_out/train_time.log 2>&1
grep -v watchdog gpurun_out/train_time.log | tail -12
