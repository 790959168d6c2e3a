#!/bin/bash
# ncu full capture of the tensor-core streaming kernels (forward + backward) on the train-step dev script.
set -x
mkdir -p gpurun_out
export VLSA_AGG_VARIANT=tc DEV_QUICK=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'agg_tc_kernel<\(bool\)0' -s 30 -c 1 -o gpurun_out/prof_agg_tc_fwd -f python scripts/dev_train_time.py child > gpurun_out/ncu_tc_fwd.log 2>&1
tail -3 gpurun_out/ncu_tc_fwd.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'agg_tc_kernel<\(bool\)1' -s 16 -c 1 -o gpurun_out/prof_agg_tc_bwd -f python scripts/dev_train_time.py child > gpurun_out/ncu_tc_bwd.log 2>&1
tail -3 gpurun_out/ncu_tc_bwd.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_train.csv python scripts/dev_train_time.py child > gpurun_out/ncu_train_launches.log 2>&1
ls -la gpurun_out
