#!/bin/bash
# One development iteration on the GPU box: parity tests, bench (optionally A/B against an env switch),
# launch list and an ncu capture of chosen kernels exported to CSV on the box.
#   TESTS   pytest -k expression (default: all gpu tests; "none" skips)   AB_ENV  e.g. "RDG_L2_PREFETCH=1": second bench run
#   NCU_K   regex of kernels to capture with --set full (empty: skip)      NCU_S / NCU_C   skip / count
#   BENCH_ARGS  extra bench flags      LAUNCH_LIST=1  per-launch time list
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
if [ "$TESTS" != "none" ]; then
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x ${TESTS:+-k "$TESTS"} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
fi
summ() {
python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], 'ms/step',round(d['ms_per_step'],3),'it/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1)); print({k:round(v,3) for k,v in d.get('stage_ms',{}).items()}); print(d['roofline']['kernel'], round(d['roofline']['frac'],4), 'step frac', round(d['step_roofline']['frac'],4), d['clocks'])
    elif 'rror' in l or 'exit' in l: print(l.strip()[:300])
PY
}
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_c4.log 2>&1; echo "exit $?" >> gpurun_out/bench_c4.log
summ gpurun_out/bench_c4.log
if [ -n "$AB_ENV" ]; then
  env $AB_ENV timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_c4_ab.log 2>&1; echo "exit $?" >> gpurun_out/bench_c4_ab.log
  summ gpurun_out/bench_c4_ab.log
fi
if [ -n "$LAUNCH_LIST" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${LL_S:-0} -c ${LL_C:-400} --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
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
    tot=sum(us(r) for r in rows[a:b])
    print('one step (launch order), total %.1f us over %d launches'%(tot,b-a))
    for r in rows[a:b]: print('  %-50s %8.1f us'%(r['Kernel Name'][:50],us(r)))
PY
fi
if [ -n "$NCU_K" ]; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -s ${NCU_S:-12} -c ${NCU_C:-2} -o gpurun_out/prof_k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k.log 2>&1
  echo "ncu exit $?"
  ncu -i gpurun_out/prof_k.ncu-rep --page raw --csv > gpurun_out/prof_k_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_k.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_k_sass.csv 2>/dev/null
  ncu -i gpurun_out/prof_k.ncu-rep --page details > gpurun_out/prof_k_details.txt 2>/dev/null
  sz=$(stat -c %s gpurun_out/prof_k.ncu-rep); if [ "$sz" -gt 40000000 ]; then rm -f gpurun_out/prof_k.ncu-rep; echo "rep dropped ($sz bytes)"; fi
fi
