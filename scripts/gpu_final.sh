#!/bin/bash
# Short closing run on the GPU box: all GPU tests, smoke, both bench shapes, variant-path timing (no ncu).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_p4.json 2> gpurun_out/bench_p4.err; tail -c 3500 gpurun_out/bench_p4.json; tail -5 gpurun_out/bench_p4.err
timeout 600 python bench.py --steps 20 --warmup 5 --P 12 --R 12 --no-e2e --no-cpu-baseline > gpurun_out/bench_p12.json 2> gpurun_out/bench_p12.err; tail -c 2500 gpurun_out/bench_p12.json; tail -5 gpurun_out/bench_p12.err
timeout 300 python scripts/dev_variant_time.py > gpurun_out/variant_time.log 2>&1; tail -4 gpurun_out/variant_time.log
for g in 4 6; do VLSA_GEN_GROUP=$g timeout 120 python scripts/dev_variant_time.py 2>&1 | tail -3; done
