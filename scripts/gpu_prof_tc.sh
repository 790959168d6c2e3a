#!/bin/bash
# ncu full capture of the tensor-core streaming kernel (forward; pass "bwd" for the backward too).
set -x
mkdir -p gpurun_out
export VLSA_AGG_VARIANT=tc DEV_QUICK=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_tc -s 30 -c 1 -o gpurun_out/prof_agg_tc_fwd -f python scripts/dev_train_time.py child > gpurun_out/ncu_tc_fwd.log 2>&1
tail -3 gpurun_out/ncu_tc_fwd.log
if [ "$1" = "bwd" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_tc -s 37 -c 1 -o gpurun_out/prof_agg_tc_bwd -f python scripts/dev_train_time.py child > gpurun_out/ncu_tc_bwd.log 2>&1
tail -3 gpurun_out/ncu_tc_bwd.log
fi
