#!/bin/bash
# Round-end style check + evidence for profiles/: all gpu tests, smoke, default bench (with cpu_baseline), reference arm,
# launch list of the bench command, ncu --set full of the dominant kernel (traffic).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.log gpurun_out/*.csv
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
( time timeout 600 python bench.py ) > gpurun_out/bench_default.log 2>&1; echo "exit $?" >> gpurun_out/bench_default.log
tail -c 3800 gpurun_out/bench_default.log
( time timeout 400 python bench.py --impl reference ) > gpurun_out/bench_reference.log 2>&1; echo "exit $?" >> gpurun_out/bench_reference.log
tail -c 1300 gpurun_out/bench_reference.log
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
    tot=sum(us(r) for r in rows[a:b])
    print('one step: %.1f us over %d launches'%(tot,b-a))
    for r in rows[a:b]: print('  %-50s %8.1f us  %4.1f %%'%(r['Kernel Name'][:50],us(r),100*us(r)/tot))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blend_bwd|blend_fwd|preprocess_bwd|dtable2|tile_sort_kernel<128|ssim" -s 21 -c 7 -o gpurun_out/prof_step -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_step.ncu-rep --page raw --csv > gpurun_out/prof_step_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_step_raw.csv > gpurun_out/prof_step_summary.txt
grep -E "Kernel Name|gpu__time_duration|dram__bytes" gpurun_out/prof_step_summary.txt | cut -c1-150
sz=$(stat -c %s gpurun_out/prof_step.ncu-rep 2>/dev/null || echo 0); if [ "$sz" -gt 30000000 ]; then rm -f gpurun_out/prof_step.ncu-rep; echo "rep dropped ($sz bytes)"; fi
ls -la gpurun_out | head -20
