#!/bin/bash
# 1 GPU: selected tests, a bench line, and a --set full capture of the backward-half kernels of one step.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
if [ "$TESTS" != "none" ]; then
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x ${TESTS:+-k "$TESTS"} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
fi
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-dropin $BENCH_ARGS > gpurun_out/bench_c4.log 2>&1; echo "exit $?" >> gpurun_out/bench_c4.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_c4.log'):
    if l.startswith('{'):
        d=json.loads(l); print('ms/step',round(d['ms_per_step'],3),'it/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1)); print({k:round(v,3) for k,v in d.get('stage_ms',{}).items()}); print('with_optimizer', d.get('with_optimizer'))
    elif 'rror' in l or 'exit' in l: print(l.strip()[:300])
PY
if [ -n "$NCU_K" ]; then
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -s ${NCU_S:-10} -c ${NCU_C:-5} -o gpurun_out/prof_bwd -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_bwd.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_bwd.ncu-rep --page raw --csv > gpurun_out/prof_bwd_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_bwd_raw.csv > gpurun_out/prof_bwd_summary.txt
fi
ls -la gpurun_out | tail -6
