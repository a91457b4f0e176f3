#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_train.py 64 2 > gpurun_out/plain_train.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain_train.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/train_launches.csv python scripts/profile_train.py 64 2 > gpurun_out/ncu_train.log 2>&1
echo "ncu exit=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/train_launches.csv')) if len(r)>10 and r[0].isdigit()]
half=len(rows)//2
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[half:]:
    name=r[4].split('(')[0].replace('void ','').replace('dfv::','')
    agg[name][0]+=1; agg[name][1]+=float(r[-1])/1e6
tot=sum(v[1] for v in agg.values())
print('second step: total %.2f ms over %d launches'%(tot, sum(v[0] for v in agg.values())))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:60]: print('%8.3f ms %5d  %s'%(v[1],v[0],k[:110]))
PY
