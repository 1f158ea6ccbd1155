#!/bin/bash
# round 2, session a: new kriging kernel (no prologue, published L(K+1,K), prefetch into dead tile rows) against the round-1
# kernel (libtwxi_base.so): GPU tests, stage timing of the benchmark tile, in-kernel cycle breakdown
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_r02a.log
for lib in libtwxi_base.so libtwxi.so; do
  echo "== $lib"
  TWXI_LIB=topowx_b200/$lib timeout 300 python tools/time_tile.py 3 2>&1 | tail -1
done | tee gpurun_out/time_r02a.log
echo "== prof"; TWXI_LIB=topowx_b200/libtwxi_prof.so timeout 300 python tools/ked_prof.py 2>&1 | tail -20 | tee gpurun_out/kedprof_r02a.log
