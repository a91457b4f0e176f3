#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_model.py -q --timeout 600 -p no:cacheprovider -s > gpurun_out/test_train_model.log 2>&1; echo "train_model exit=$? $(tail -1 gpurun_out/test_train_model.log)"
grep -E "cosine|worst" gpurun_out/test_train_model.log
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit=$?"
tail -5 gpurun_out/bench_train.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/bench_train.json'))
    print('value',round(j['value']),'img/s  ms/step',round(j['ms_per_step'],2),' e2e',round(j['e2e']['value']), 'mem GB', round(j['memory_gb'],1), 'launches', j['gpu_launches'])
    for k,v in sorted(j['roofline']['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step']): print(f"  {k:14s} {v['launches_per_step']:4d} launches {v['ms_per_step']:.3f} ms  hbm_frac {v['hbm_frac']:.3f}  tflops {v['tflops']:.1f}")
except Exception as e:
    print('bench parse failed', e)
PY
