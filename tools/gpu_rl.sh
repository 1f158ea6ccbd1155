#!/bin/bash
# right-looking register-resident kriging kernel (ked_rl.cu): parity tests with it enabled, stage timing with and
# without it, per-size-class durations (ncu launch list, one variable pass)
mkdir -p gpurun_out
RL=${1:-1}
TWXI_KED_RL=$RL timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
TWXI_KED_RL=$RL timeout 200 python tools/time_tile.py 3 2>&1 | tail -1
TWXI_KED_RL=$RL timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ked_ -c 40 --csv --log-file gpurun_out/kedcls_rl.csv python tools/prof_chunk.py 250 250 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/kedcls_rl.csv')) if len(r)>10 and r[0].isdigit()]
print(' '.join('%s:%d:%.0f'%(r[4].split('<')[1].split('>')[0].replace(' ','') if '<' in r[4] else r[4][:12], int(r[8].strip('()').split(',')[0]), float(r[14])/1e3) for r in rows if ('ked_kernel' in r[4] or 'ked_rl' in r[4]) and float(r[14]) > 2e4))
PY
if [ -n "$2" ]; then   # ncu --set full of one size class of the right-looking kernel (e.g. "ked_rl_kernel<10")
  TWXI_KED_RL=$RL timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" --launch-count 1 -o gpurun_out/rl_full -f python tools/prof_chunk.py 250 250 1 > gpurun_out/ncu_rl_full.log 2>&1; tail -2 gpurun_out/ncu_rl_full.log
fi
