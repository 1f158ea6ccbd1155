#!/bin/bash
# round 2, session c: GPU tests (incl. the round-2 additions), smoke, C5 bench line at N=1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_r02c.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02c.err; cat gpurun_out/bench_r02c.json
