#!/bin/bash
# Session-8 call B: fixed MLP tests + rewritten MLP kernels, occupancy variants of the preprocess kernels (A/B), launch list.
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
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.log 2>&1; echo "exit $?" >> gpurun_out/bench_default.log
summ gpurun_out/bench_default.log
for sw in RDG_PRE_FWD_MINB=3 RDG_PRE_BWD_MINB=3; do
  env $sw timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$sw.log 2>&1; echo "exit $?" >> gpurun_out/bench_$sw.log
  echo "--- $sw"; summ gpurun_out/bench_$sw.log
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
def us(r):
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    return v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
seen=0
for r in rows:
    if 'basis_mlp' in r['Kernel Name'] and seen<6: print('  %-50s %8.1f us'%(r['Kernel Name'][:50],us(r))); seen+=1
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"basis_mlp|dtable2" -s 10 -c 4 -o gpurun_out/prof_mlp -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mlp.log 2>&1
echo "ncu mlp exit $?"
ncu -i gpurun_out/prof_mlp.ncu-rep --page raw --csv > gpurun_out/prof_mlp_raw.csv 2>/dev/null
ls -la gpurun_out | head -20
