#!/bin/bash
# N-GPU check (N = $1, default 2): inference (no collective) and train step (NCCL gradient all-reduce)
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_side_apis.py -q --timeout 300 -p no:cacheprovider 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "infer N=$N exit=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --mode train --steps 5 --warmup 3 > gpurun_out/bench_train_n$N.json 2> gpurun_out/bench_train_n$N.err; echo "train N=$N exit=$?"
python - <<PY
import json
for f in ('gpurun_out/bench_n$N.json','gpurun_out/bench_train_n$N.json'):
    try:
        j=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'n_gpus', j['n_gpus'], 'value', round(j['value']), 'img/s  ms/step', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value']))
    except Exception as e:
        print(f, 'parse failed', e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
