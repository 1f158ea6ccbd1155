#!/bin/bash
# round 2 final measurements: GPU tests, smoke, bench line at N=1 (with secondary figures and CPU baseline), reference arm, ncu
# launch list of the bench command, DRAM traffic of the kriging launches (C5 tile), ncu --set full of the three stage kernels
mkdir -p gpurun_out
COMMIT=${1:-unknown}
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r02.log 2>&1; tail -3 gpurun_out/pytest_r02.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r02.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err; echo "ref rc=$?"
# launch list (cold-cache, serialised): 2 tiles of the list
TWX_BENCH_TILES=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/b_ncu_r02.log 2>&1; tail -1 gpurun_out/b_ncu_r02.log | cut -c1-120
# DRAM traffic of the ked_kernel launches of one C5 tile (both variables)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ked_kernel -c 60 --csv --log-file gpurun_out/ked_traffic_r02.csv python tools/prof_tile_c5.py 1 > /dev/null 2>&1
python - <<PY
import csv, json
rows=[r for r in csv.reader(open('gpurun_out/ked_traffic_r02.csv')) if len(r)>10 and r[0].isdigit()]
tot={}
for r in rows:
    tot[r[12]]=tot.get(r[12],0.0)+float(r[14])*({'Mbyte':1e6,'Kbyte':1e3,'Gbyte':1e9,'byte':1,'ns':1,'us':1e3,'ms':1e6,'usecond':1e3,'nsecond':1,'msecond':1e6}.get(r[13],1))
d={"dram_bytes_read":tot.get("dram__bytes_read.sum"),"dram_bytes_write":tot.get("dram__bytes_write.sum"),"ked_kernel_ns":tot.get("gpu__time_duration.sum"),
   "launches":len(rows)//3,"commit":"$COMMIT","workload":"first tile of the C5 list (62 500 land cells, 10 000 stations), both variables, one step; ncu --clock-control none, cold cache per launch"}
json.dump(d,open('gpurun_out/ked_traffic_r02.json','w'),indent=1); print(d)
PY
# ncu --set full of one launch of each stage kernel on a 100 x 100 chunk of the C5 tile
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ked_kernel' --launch-skip 9 --launch-count 1 -o gpurun_out/ked_r02 -f python tools/prof_tile_c5.py 1 100 > gpurun_out/ncu_ked_r02.log 2>&1; tail -1 gpurun_out/ncu_ked_r02.log | cut -c1-100
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gwr_kernel|knn_kernel|nngh_params' --launch-count 3 -o gpurun_out/stages_r02 -f python tools/prof_tile_c5.py 1 100 > gpurun_out/ncu_stages_r02.log 2>&1; tail -1 gpurun_out/ncu_stages_r02.log | cut -c1-100
ls -la gpurun_out/*.ncu-rep
