#!/bin/bash
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ked_kernel -c 14 -o gpurun_out/ked_e python tools/prof_chunk.py 100 100 1 > gpurun_out/ncu_ked_e.log 2>&1; tail -2 gpurun_out/ncu_ked_e.log
