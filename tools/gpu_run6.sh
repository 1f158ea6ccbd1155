#!/bin/bash
# KED v3 validation: GPU tests, bench (no CPU baseline), ncu --set full on the KED kernel
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -3 gpurun_out/bench_d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_d.json'))
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms'], d['gpu_launches'])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ked_kernel -c 12 -o gpurun_out/ked_d python tools/prof_chunk.py 100 100 1 > gpurun_out/ncu_ked_d.log 2>&1; tail -2 gpurun_out/ncu_ked_d.log
