#!/bin/bash
# training-path iteration: op tests + whole-model training parity + train-step bench
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q --timeout 600 -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; echo "$name exit=$? $(tail -1 gpurun_out/test_$name.log)"; }
run train_ops tests/test_gpu_train_ops.py
run train_model tests/test_gpu_train_model.py -s
grep -E "FAILED|Error|error|assert" gpurun_out/test_train_ops.log | head -30
grep -E "FAILED|Error|error|assert|worst|cosine" gpurun_out/test_train_model.log | head -30
python bench.py --mode train --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train exit=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_train.json',):
    try:
        j=json.load(open(f))
        print(f, 'value',round(j['value']),'img/s  ms/step',round(j['ms_per_step'],2),' e2e',round(j['e2e']['value']), 'launches', j['gpu_launches'])
        for k,v in sorted(j['roofline']['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step']): print(f"  {k:14s} {v['launches_per_step']:4d} {v['ms_per_step']:.3f} ms  hbm_frac {v['hbm_frac']:.3f}  tflops {v['tflops']:.1f}")
    except Exception as e:
        print(f, 'parse failed', e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
