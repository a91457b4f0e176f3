#!/bin/bash
# inference iteration: op tests + model tests + bench with per-launch dump
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest "$@" -q --timeout 300 -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; echo "$name exit=$? $(tail -1 gpurun_out/test_$name.log)"; }
run ops tests/test_gpu_ops.py
run model tests/test_gpu_model.py
grep -hE "^FAILED|^ERROR" gpurun_out/test_ops.log gpurun_out/test_model.log | head -20
DFV_BENCH_DUMP=gpurun_out/infer_launches.json python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/bench.json'))
    print('value',round(j['value']),'img/s  ms/step',round(j['ms_per_step'],2),' e2e',round(j['e2e']['value']))
    for k,v in sorted(j['roofline']['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step']): print(f"  {k:14s} {v['launches_per_step']:4d} {v['ms_per_step']:.3f} ms  hbm_frac {v['hbm_frac']:.3f}  tflops {v['tflops']:.1f}")
    L=json.load(open('gpurun_out/infer_launches.json'))
    for r in L:
        if r['kind'] in ('dwconv','expand_gemm','project_gemm','stem','se_gate'):
            print(f"   {r['kind']:13s} {r['ms']*1000:8.1f} us  {r['bytes']/1e6:9.1f} MB  {r['bytes']/r['ms']/1e6:7.0f} GB/s  {r['flops']/r['ms']/1e9:7.1f} TF")
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
