#!/bin/bash
# lean covariance + stable problem order: tests, stage timing, DRAM traffic of the kriging kernels, shared-pipe microbenchmark
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_k.log 2>&1; tail -3 gpurun_out/pytest_k.log
timeout 300 python tools/time_tile.py 3 2>&1 | tail -1
TWXI_HC_BUDGET_MB=400 timeout 300 python tools/time_tile.py 3 2>&1 | tail -1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'ked_kernel|hgather' -c 48 --csv --log-file gpurun_out/ked_traffic_k.csv python tools/prof_chunk.py 250 250 1 > /dev/null 2>&1
python - <<'PY'
import csv, json
rows=[r for r in csv.reader(open('gpurun_out/ked_traffic_k.csv')) if len(r)>10 and r[0].isdigit()]
tot={}
for r in rows:
    k=('hgather ' if 'hgather' in r[4] else 'ked ')+r[12]
    tot[k]=tot.get(k,0.0)+float(r[14])*({'Mbyte':1e6,'Kbyte':1e3,'Gbyte':1e9,'byte':1,'ns':1,'us':1e3,'ms':1e6}.get(r[13],1))
print(json.dumps(tot), len(rows)//3, 'launches')
PY
timeout 200 ./tools/fp64_peak > gpurun_out/fp64_peak_k.jsonl 2>&1; grep mixed gpurun_out/fp64_peak_k.jsonl
