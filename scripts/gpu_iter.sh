#!/bin/bash
# quick iteration: op tests (separate processes) + model tests + bench
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest "$@" -q --timeout 300 -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; echo "$name exit=$? $(tail -1 gpurun_out/test_$name.log)"; }
run ops tests/test_gpu_ops.py
run model tests/test_gpu_model.py
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/bench.json'))
    print('value',round(j['value']),'img/s  ms/step',round(j['ms_per_step'],2),' e2e',round(j['e2e']['value']))
    for k,v in j['roofline']['kernels'].items(): print(f"  {k:14s} {v['ms_per_step']:.3f} ms  hbm_frac {v['hbm_frac']:.3f}  tflops {v['tflops']:.1f}")
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
