#!/bin/bash
ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__shared_mem_per_block_dynamic,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --csv --log-file gpurun_out/launches_chunk.csv python tools/prof_chunk.py 125 125 1 > gpurun_out/ncu_chunk.log 2>&1
tail -2 gpurun_out/ncu_chunk.log
