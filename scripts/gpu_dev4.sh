#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
VLSA_AGG_VARIANT=simt DEV_CFGS=simt timeout 900 python scripts/dev_variants.py > gpurun_out/variants.log 2>&1
tail -40 gpurun_out/variants.log
