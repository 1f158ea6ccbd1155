#!/bin/bash
set -x
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_a.json 2> gpurun_out/bench_r01_a.err; tail -5 gpurun_out/bench_r01_a.err; cat gpurun_out/bench_r01_a.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01_a.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
