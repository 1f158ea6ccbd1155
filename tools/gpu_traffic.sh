#!/bin/bash
# DRAM traffic of the kriging kernel on the benchmark tile (one step = 2 variable passes), for bench.py's roofline.traffic
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ked_kernel -c 48 --csv --log-file gpurun_out/ked_traffic.csv python tools/prof_chunk.py 250 250 1 > /dev/null 2>&1
python - <<'PY'
import csv, json
rows=[r for r in csv.reader(open('gpurun_out/ked_traffic.csv')) if len(r)>10 and r[0].isdigit()]
tot={}
for r in rows:
    tot[r[12]]=tot.get(r[12],0.0)+float(r[14])*({'Mbyte':1e6,'Kbyte':1e3,'Gbyte':1e9,'byte':1,'ns':1,'us':1e3,'ms':1e6}.get(r[13],1))
print(json.dumps(tot), len(rows)//3, 'launches')
PY
