#!/bin/bash
# round-1 final measurements (session 5): GPU tests, bench line (with CPU baseline), reference arm, ncu launch list of
# the bench command, ncu --set full of every stage kernel on a 100x100 chunk
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_n.log 2>&1; tail -3 gpurun_out/pytest_n.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; tail -3 gpurun_out/bench_n.err; cat gpurun_out/bench_n.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_n_ref.json 2> gpurun_out/bench_n_ref.err; cat gpurun_out/bench_n_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_n.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_n.log 2>&1; tail -1 gpurun_out/b_ncu_n.log | cut -c1-200
python __graft_entry__.py --smoke 2>&1 | tail -1
