#!/bin/bash
# ncu launch list of the bench command (8 tiles of the C5 list, largest first: the first tiles are full-land tiles)
mkdir -p gpurun_out
TWX_BENCH_TILES=8 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/b_ncu_r02.log 2>&1; tail -1 gpurun_out/b_ncu_r02.log | cut -c1-100
