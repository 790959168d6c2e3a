#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/dev_variants.py 2>&1 | grep -v watchdog
VLSA_B200_LIB=vlsa_b200/lib/variants/b_trunc.so VLSA_AGG_VARIANT=tc timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
