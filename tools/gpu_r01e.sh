#!/bin/bash
# round-1 re-entry validation: GPU tests, full bench line, ncu launch list, ncu --set full of the KED kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; tail -3 gpurun_out/bench_e.err; cat gpurun_out/bench_e.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_e.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1; tail -2 gpurun_out/b_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ked_kernel -c 12 -o gpurun_out/ked_e python tools/prof_chunk.py 100 100 1 > gpurun_out/ncu_ked_e.log 2>&1; tail -2 gpurun_out/ncu_ked_e.log
