#!/bin/bash
# ncu full capture of the tensor-core streaming kernels (forward + backward) on the train-step dev script.
set -x
mkdir -p gpurun_out
DEV_QUICK=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_tc -s 6 -c 2 -o gpurun_out/prof_agg_tc_v2 -f python scripts/dev_train_time.py child > gpurun_out/ncu_tc_v2.log 2>&1
tail -5 gpurun_out/ncu_tc_v2.log
ls -la gpurun_out
