#!/bin/bash
# parity tests + bench (c4) ; optional ncu of chosen kernels via $NCU_K
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4.log 2>&1; echo "exit $?" >> gpurun_out/bench_c4.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_c4.log'):
    if l.startswith('{'):
        d=json.loads(l); print('ms/step',round(d['ms_per_step'],3),'it/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1)); print({k:round(v,3) for k,v in d['stage_ms'].items()}); print(d['roofline']['kernel'], round(d['roofline']['frac'],4), 'step frac', round(d['step_roofline']['frac'],4), d['clocks'])
    elif 'rror' in l or 'exit' in l: print(l.strip()[:300])
PY
if [ -n "$NCU_K" ]; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -s ${NCU_S:-12} -c ${NCU_C:-2} -o gpurun_out/prof_k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k.log 2>&1
  echo "ncu exit $?"
fi
