#!/bin/bash
# round 2, session f: station-major tables (nngh_params, GWR, KED B' build): GPU tests + stage timing on configs[1] and on a C5 tile
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_r02f.log
timeout 300 python tools/time_tile.py 3 2>&1 | tail -1 | tee gpurun_out/time_r02f.log
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02f.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['rooflines']['stage_ms'], d['roofline']['frac'], d['rooflines']['gwr_kernel']['frac'])
PY
