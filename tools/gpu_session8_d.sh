#!/bin/bash
# Session-8 call D: shared-memory difference table in the preprocess kernels - parity, A/B, ncu.
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
RDG_DIFF_SMEM=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_nodiff.log 2>&1; echo "exit $?" >> gpurun_out/bench_c4_nodiff.log
summ gpurun_out/bench_c4_nodiff.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --config c2_kubric > gpurun_out/bench_c2.log 2>&1; echo "exit $?" >> gpurun_out/bench_c2.log
summ gpurun_out/bench_c2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config c5_infer --forward-only > gpurun_out/bench_c5.log 2>&1; echo "exit $?" >> gpurun_out/bench_c5.log
summ gpurun_out/bench_c5.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"preprocess_fwd|preprocess_bwd" -s 7 -c 2 -o gpurun_out/prof_k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_k.ncu-rep --page raw --csv > gpurun_out/prof_k_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_k_raw.csv | grep -v "occupancy_limit\|shared_atom"
ls -la gpurun_out | head -20
