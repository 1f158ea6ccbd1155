#!/bin/bash
# per-size-class duration and resident CTAs per SM of the default kriging launch configuration (one variable pass)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ked_kernel -c 21 --csv --log-file gpurun_out/kedcls.csv python tools/prof_chunk.py 250 250 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/kedcls.csv')) if len(r)>10 and r[0].isdigit()]
print(' '.join('%s:%d:%.0f'%(r[4].split('<')[1].split('>')[0].replace(' ',''), int(r[8].strip('()').split(',')[0])//148, float(r[14])/1e3) for r in rows))
PY
