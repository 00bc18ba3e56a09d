#!/bin/bash
# Session-8 call C: all GPU tests, default bench (A/B of the prefetch changes is against call B's default: 2.928 ms),
# other BASELINE configs, launch list, ncu of the preprocess kernels and dtable2.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.log gpurun_out/*.csv
summ() {
python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], 'ms/step',round(d['ms_per_step'],3),'it/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1)); print({k:round(v,3) for k,v in d.get('stage_ms',{}).items()}); print(d['roofline']['kernel'], round(d['roofline']['frac'],4), 'step frac', round(d['step_roofline']['frac'],4), d['clocks'], d.get('basis_mlp'))
    elif 'rror' in l or 'exit' in l: print(l.strip()[:300])
PY
}
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4.log 2>&1; echo "exit $?" >> gpurun_out/bench_c4.log
summ gpurun_out/bench_c4.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --config c2_kubric > gpurun_out/bench_c2.log 2>&1; echo "exit $?" >> gpurun_out/bench_c2.log
summ gpurun_out/bench_c2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config c3_nvidia > gpurun_out/bench_c3.log 2>&1; echo "exit $?" >> gpurun_out/bench_c3.log
summ gpurun_out/bench_c3.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config c5_infer --forward-only > gpurun_out/bench_c5.log 2>&1; echo "exit $?" >> gpurun_out/bench_c5.log
summ gpurun_out/bench_c5.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
def us(r):
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    return v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
names=[r['Kernel Name'][:50] for r in rows]
idx=[i for i,n in enumerate(names) if 'preprocess_fwd' in n]
if len(idx)>=2:
    a,b=idx[-2],idx[-1]
    print('one step: %.1f us over %d launches'%(sum(us(r) for r in rows[a:b]),b-a))
    for r in rows[a:b]: print('  %-50s %8.1f us'%(r['Kernel Name'][:50],us(r)))
seen=0
for r in rows:
    if 'basis_mlp' in r['Kernel Name'] and seen<3: print('  %-50s %8.1f us'%(r['Kernel Name'][:50],us(r))); seen+=1
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"preprocess_fwd|preprocess_bwd|dtable2|basis_mlp_fwd" -s 7 -c 4 -o gpurun_out/prof_k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_k.ncu-rep --page raw --csv > gpurun_out/prof_k_raw.csv 2>/dev/null
ls -la gpurun_out | head -20
