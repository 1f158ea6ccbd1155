#!/bin/bash
# KED v5 (one warp per problem) vs v4 (CTA per problem): GPU tests, stage timing, per-size-class launch durations
mkdir -p gpurun_out
OLD=222222223333444455555
NEW=666666666666777788888
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_i.log 2>&1; tail -5 gpurun_out/pytest_i.log
TWXI_KED_VAR=$NEW timeout 300 python tools/time_tile.py 3 2>&1 | tail -1
TWXI_KED_VAR=$OLD timeout 300 python tools/time_tile.py 3 2>&1 | tail -1
for cfg in $NEW $OLD; do
  TWXI_KED_VAR=$cfg timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ked_kernel|ked_warp_kernel' -c 21 --csv --log-file gpurun_out/kedvar_$cfg.csv python tools/prof_chunk.py 250 250 1 > /dev/null 2>&1
done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/kedvar_*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    print(f[-25:-4], ' '.join('%s:%d:%.0f'%(r[4].split('<')[1].split('>')[0].replace(' ',''), int(r[8].strip('()').split(',')[0])//148, float(r[14])/1e3) for r in rows))
PY
