#!/bin/bash
# Closing evidence run of round 2 on the GPU box: tests, smoke, bench lines (fp32 headline with every record, bf16 storage at
# P = 4 / 16, reference arm), ncu launch lists of the same bench commands, full captures of the streaming kernels, step / call
# wall clocks.
set -x
O=gpurun_out/r02b; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/nvsmi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench_p4.json 2> $O/bench_p4.err; tail -c 300 $O/bench_p4.json; tail -3 $O/bench_p4.err
timeout 600 python bench.py --dtype bf16 --no-e2e --no-cpu-baseline --no-torch-gpu > $O/bench_bf16_p4.json 2> $O/bench_bf16_p4.err; tail -c 300 $O/bench_bf16_p4.json
timeout 600 python bench.py --dtype bf16 --P 16 --R 16 --no-e2e --no-cpu-baseline --no-torch-gpu --no-shipped --no-ragged > $O/bench_bf16_p16.json 2> $O/bench_bf16_p16.err; tail -c 300 $O/bench_bf16_p16.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_ref.json 2>&1; tail -c 400 $O/bench_ref.json
timeout 300 python scripts/dev_step_wall.py > $O/step_wall.log 2>&1; cp gpurun_out/step_wall.json $O/step_wall.json; tail -16 $O/step_wall.log
# launch lists (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_p4.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-gpu --no-shipped --no-ragged > $O/ncu_bench_p4.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_bf16_p12.csv python bench.py --dtype bf16 --steps 3 --warmup 3 --P 12 --R 12 --no-e2e --no-cpu-baseline --no-torch-gpu --no-ragged --no-shipped > $O/ncu_bench_bf16.log 2>&1
# full captures (the simt one also feeds roofline.traffic of the bench line)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_simt -s 4 -c 1 -o $O/prof_agg_simt_p4 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train --no-torch-gpu --no-shipped --no-ragged > $O/ncu_full_simt.log 2>&1
DEV_DTYPE=bf16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_bf16_kernel -s 2 -c 1 -o $O/prof_agg_bf16_p12_fwd -f python scripts/dev_one_launch.py tc 12 > $O/ncu_full_bf16.log 2>&1
ls -la $O
