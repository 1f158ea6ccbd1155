#!/bin/bash
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -3 gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_b.json'))
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms'], d['gpu_launches'])
PY
ncu --set full --clock-control none --import-source on -k regex:ked_kernel -c 6 -o gpurun_out/ked_b python tools/prof_chunk.py 50 50 1 > gpurun_out/ncu_ked_b.log 2>&1; tail -2 gpurun_out/ncu_ked_b.log
