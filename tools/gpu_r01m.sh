#!/bin/bash
# round-1 final measurements (session 5): GPU tests, bench line (with CPU baseline), reference arm, ncu launch list of
# the bench command, ncu --set full of every stage kernel on a 100x100 chunk
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_p.log 2>&1; tail -3 gpurun_out/pytest_p.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; tail -3 gpurun_out/bench_p.err; cat gpurun_out/bench_p.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_p_ref.json 2> gpurun_out/bench_p_ref.err; cat gpurun_out/bench_p_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_p.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_p.log 2>&1; tail -1 gpurun_out/b_ncu_p.log | cut -c1-200
python __graft_entry__.py --smoke 2>&1 | tail -1
