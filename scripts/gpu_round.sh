#!/bin/bash
# Runs on the GPU box: tests, smoke, bench (both shapes + reference arm), ncu launch lists + full captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_p4.json 2> gpurun_out/bench_p4.err; tail -c 3500 gpurun_out/bench_p4.json; tail -5 gpurun_out/bench_p4.err
timeout 600 python bench.py --steps 20 --warmup 5 --P 12 --R 12 --no-e2e --no-cpu-baseline > gpurun_out/bench_p12.json 2> gpurun_out/bench_p12.err; tail -c 2500 gpurun_out/bench_p12.json; tail -5 gpurun_out/bench_p12.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
timeout 300 python scripts/dev_variant_time.py > gpurun_out/variant_time.log 2>&1; tail -3 gpurun_out/variant_time.log
# launch lists of the same bench commands (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_p4.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-train > gpurun_out/ncu_bench_p4.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_p12_train.csv python bench.py --steps 3 --warmup 3 --P 12 --R 12 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_p12.log 2>&1
# full captures: CUDA-core kernel at P=4 (headline), tensor-core forward and backward at P=12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_simt -s 4 -c 1 -o gpurun_out/prof_agg_simt_p4 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train > gpurun_out/ncu_full_simt.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_tc -s 12 -c 2 -o gpurun_out/prof_agg_tc_p12 -f python bench.py --steps 2 --warmup 3 --P 12 --R 12 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_tc.log 2>&1
ls -la gpurun_out
