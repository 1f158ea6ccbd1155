#!/bin/bash
# the complete C5 list (all 255 land tiles of the CONUS grid = 12.65 M land cells x 365 days) at N GPUs
N=${1:-8}
mkdir -p gpurun_out
export TWX_BENCH_TILES=all
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/bench_all_n1.json 2> gpurun_out/bench_all_n1.err; echo rc=$?
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_all_n$N.json 2> gpurun_out/bench_all_n$N.err; echo rc=$?
fi
tail -3 gpurun_out/bench_all_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_all_n$N.json'))
print('N', d['n_gpus'], 'tiles', d['config']['tiles'], 'cells', d['config']['land_cells'], 'value %.4g e2e %.4g ms/step %.1f setup %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['setup_s']), d['per_rank'])
PY
