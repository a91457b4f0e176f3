#!/bin/bash
mkdir -p gpurun_out
for v in 0 256 512; do
  DFV_DEBUG_FLAGS=$v python bench.py --mode train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_v$v.json 2> gpurun_out/bench_train_v$v.err
  python - <<PY
import json
j=json.load(open('gpurun_out/bench_train_v$v.json'))
print('variant $v', round(j['ms_per_step'],2), 'ms/step  bn_act', round(j['roofline']['kernels']['bn_act']['ms_per_step'],2))
PY
done
