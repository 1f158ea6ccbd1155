#!/bin/bash
# cycle breakdown of the KED kernel (instrumented build) + ncu --set full of the gwr / knn / nngh / hgather kernels
mkdir -p gpurun_out
TWXI_LIB=$PWD/topowx_b200/libtwxi_prof.so timeout 600 python tools/ked_prof.py 2>&1 | tail -14
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gwr_kernel|knn_kernel|nngh_params|hgather' -c 8 -o gpurun_out/misc_f python tools/prof_chunk.py 100 100 1 > gpurun_out/ncu_misc_f.log 2>&1; tail -2 gpurun_out/ncu_misc_f.log
