"""Compare gpurun_out/infer_launches_new.json (+ bench_new.json) with a saved baseline dump (argv[1], default infer_launches.json)."""
import json, sys
base = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/infer_launches.json'
d = json.loads([l for l in open('gpurun_out/bench_new.json') if l.startswith('{')][0])
print('img/s', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'train', round(d['train_step']['value']),
      round(d['train_step']['ms_per_step'], 2), d['clocks'])
a = json.load(open(base)); b = json.load(open('gpurun_out/infer_launches_new.json'))
ta = {}; tb = {}
for x, y in zip(a, b):
    ta[x['kind']] = ta.get(x['kind'], 0) + x['ms']; tb[y['kind']] = tb.get(y['kind'], 0) + y['ms']
for k in ta: print(f"{k:14s} {ta[k]:.3f} -> {tb[k]:.3f}")
print('sum', round(sum(ta.values()), 3), '->', round(sum(tb.values()), 3))
if len(sys.argv) > 2:
    for i, (x, y) in enumerate(zip(a, b)):
        if x['kind'] in sys.argv[2:]: print(i, x['kind'], round(x['ms'] * 1e3, 1), round(y['ms'] * 1e3, 1), f"{y['bytes'] / y['ms'] / 1e6 / 6449.4:.2f}")
