#!/bin/bash
# Session-8 measurement call: new kernels (basis MLP, dtable2, double-buffered preprocess, tile order) - parity first,
# then A/B benches, launch list and ncu captures of the changed kernels.  Every step has its own timeout.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
summ() {
python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], 'ms/step',round(d['ms_per_step'],3),'it/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1)); print({k:round(v,3) for k,v in d.get('stage_ms',{}).items()}); print(d['roofline']['kernel'], round(d['roofline']['frac'],4), 'step frac', round(d['step_roofline']['frac'],4), d['clocks'], d.get('basis_mlp'))
    elif 'rror' in l or 'exit' in l: print(l.strip()[:300])
PY
}
# 1. quick sanity of the riskiest changes (mbarrier double buffering): a hang here must not eat the call
timeout 300 python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -k "fused" > gpurun_out/pytest_fused.log 2>&1
rc=$?; echo "pytest fused exit $rc" >> gpurun_out/pytest_fused.log; tail -12 gpurun_out/pytest_fused.log
if [ $rc -ne 0 ]; then echo "FALLBACK: RDG_PRE_DB=0 for the rest of this call"; export RDG_PRE_DB=0; fi
timeout 200 python -m pytest tests/test_basis_mlp.py -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/pytest_mlp.log 2>&1
echo "pytest mlp exit $?" >> gpurun_out/pytest_mlp.log; tail -25 gpurun_out/pytest_mlp.log
# 2. everything
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
# 3. benches: default (all new paths), then one switch off at a time
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_new.log 2>&1; echo "exit $?" >> gpurun_out/bench_new.log
summ gpurun_out/bench_new.log
for sw in RDG_PRE_DB=0 RDG_TILE_ORDER=0 RDG_DTABLE_V1=1; do
  env $sw timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$sw.log 2>&1; echo "exit $?" >> gpurun_out/bench_$sw.log
  echo "--- $sw"; summ gpurun_out/bench_$sw.log
done
# 4. launch list
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
for r in rows:
    if 'basis_mlp' in r['Kernel Name']: print('  %-50s %8.1f us'%(r['Kernel Name'][:50],us(r)))
PY
# 5. ncu --set full of the changed kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"preprocess_fwd|preprocess_bwd|dtable2|tile_scan" -s 14 -c 4 -o gpurun_out/prof_k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_k.ncu-rep --page raw --csv > gpurun_out/prof_k_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_k.ncu-rep --page details > gpurun_out/prof_k_details.txt 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"basis_mlp" -s 9 -c 3 -o gpurun_out/prof_mlp -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mlp.log 2>&1
echo "ncu mlp exit $?"
ncu -i gpurun_out/prof_mlp.ncu-rep --page raw --csv > gpurun_out/prof_mlp_raw.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/prof_k.ncu-rep 2>/dev/null || echo 0); if [ "$sz" -gt 25000000 ]; then rm -f gpurun_out/prof_k.ncu-rep; echo "rep dropped ($sz bytes)"; fi
ls -la gpurun_out/ | head -30
