#!/bin/bash
# round-1 final measurements: GPU tests, bench line (with CPU baseline), reference arm, ncu launch list of the bench
# command, ncu --set full of every stage kernel on a 100x100 chunk
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; tail -3 gpurun_out/bench_h.err; cat gpurun_out/bench_h.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_h_ref.json 2> gpurun_out/bench_h_ref.err; cat gpurun_out/bench_h_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_h.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_h.log 2>&1; tail -1 gpurun_out/b_ncu_h.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ked_kernel|gwr_kernel|knn_kernel|knn_candidates|nngh_params|hgather|fixer' -c 24 -o gpurun_out/all_h python tools/prof_chunk.py 100 100 1 > gpurun_out/ncu_all_h.log 2>&1; tail -2 gpurun_out/ncu_all_h.log
