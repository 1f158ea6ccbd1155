#!/bin/bash
# round-1 final measurements (session 5): GPU tests, bench line (with CPU baseline), reference arm, ncu launch list of
# the bench command, ncu --set full of every stage kernel on a 100x100 chunk
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_m.log 2>&1; tail -3 gpurun_out/pytest_m.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; tail -3 gpurun_out/bench_m.err; cat gpurun_out/bench_m.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_m_ref.json 2> gpurun_out/bench_m_ref.err; cat gpurun_out/bench_m_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_m.log 2>&1; tail -1 gpurun_out/b_ncu_m.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ked_kernel|gwr_kernel|knn_kernel|knn_candidates|nngh_params|hgather|fixer' -c 30 -o gpurun_out/all_m -f python tools/prof_chunk.py 100 100 1 > gpurun_out/ncu_all_m.log 2>&1; tail -2 gpurun_out/ncu_all_m.log
