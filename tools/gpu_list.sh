#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-150} -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
import csv, collections
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
agg=collections.defaultdict(lambda:[0,0.0])
for row in csv.DictReader(lines):
    name=row['Kernel Name'][:48]; v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]:
    print(f"{k:48s} n={v[0]:4d} share={v[1]/tot*100:5.1f}% avg={v[1]/v[0]:8.1f}us")
PY
