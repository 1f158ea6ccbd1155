#!/bin/bash
# re-entry validation: GPU tests, bench, launch list, ncu --set full on the KED kernel
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -3 gpurun_out/bench_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c.json'))
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms'], d['gpu_launches'], d['cpu_baseline'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_c.log 2>&1
tail -2 gpurun_out/ncu_bench_c.log
ncu --set full --clock-control none --import-source on -k regex:ked_kernel -c 12 -o gpurun_out/ked_c python tools/prof_chunk.py 100 100 1 > gpurun_out/ncu_ked_c.log 2>&1; tail -2 gpurun_out/ncu_ked_c.log
