#!/bin/bash
# bench.py at N GPUs the way the driver launches it (N = $1), plus the reference arm when $2 = ref
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?
fi
tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print('N', d['n_gpus'], 'value %.4g e2e %.4g ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['per_rank'], d['gpu_launches'], d['clocks'])
PY
if [ "$2" = "ref" ]; then
  timeout 900 python bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?; cat gpurun_out/bench_ref.json | cut -c1-600
fi
