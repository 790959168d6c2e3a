#!/bin/bash
# Dev round on the GPU box: parity tests with the default dispatch and with every P forced through the
# tensor-core kernels, then kernel / train-step timings per variant.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
VLSA_AGG_VARIANT=tc timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q > gpurun_out/pytest_gpu_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_tc.log
tail -15 gpurun_out/pytest_gpu_tc.log
timeout 900 python scripts/dev_train_time.py tc simt > gpurun_out/train_time.log 2>&1
cat gpurun_out/train_time.log
