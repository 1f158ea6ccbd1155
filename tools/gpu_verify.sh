#!/bin/bash
# re-verification of the tree as committed: GPU tests, smoke, bench line (with CPU baseline)
mkdir -p gpurun_out
timeout 700 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_q.log 2>&1; tail -3 gpurun_out/pytest_q.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -3 gpurun_out/bench_q.err; cat gpurun_out/bench_q.json
