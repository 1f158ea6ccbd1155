#!/bin/bash
# per-size-class duration of ked_kernel under each launch variant (ncu launch list, one variable pass each)
mkdir -p gpurun_out
true
for cfg in 000000000000000000000 111111111111111111111 222222222222222222222 333333333333333333333 444444444444444444444; do
  TWXI_KED_VAR=$cfg timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ked_kernel -c 21 --csv --log-file gpurun_out/kedvar_$cfg.csv python tools/prof_chunk.py 250 250 1 > /dev/null 2>&1
done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/kedvar_*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    print(f[-25:-4], ' '.join('%s:%d:%.0f'%(r[4].split('<')[1].split('>')[0].replace(' ',''), int(r[8].strip('()').split(',')[0])//148, float(r[14])/1e3) for r in rows))
PY
