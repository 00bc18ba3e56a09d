#!/bin/bash
# The other BASELINE configs and SH degree 0 on N GPUs (default 1): one bench line each -> gpurun_out/bench_<name>_n<N>.log
mkdir -p gpurun_out
N=${NGPU:-1}
run() {
  name=$1; shift
  if [ "$N" -gt 1 ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps ${STEPS:-30} --warmup 5 "$@" > gpurun_out/bench_${name}_n$N.log 2>&1
  else
    timeout 600 python bench.py --steps ${STEPS:-30} --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_${name}_n$N.log 2>&1
  fi
  echo "exit $?" >> gpurun_out/bench_${name}_n$N.log
  python - gpurun_out/bench_${name}_n$N.log <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], 'n', d['n_gpus'], 'ms/step',round(d['ms_per_step'],3), d['unit'], round(d['value'],1),'e2e',round(d['e2e']['value'],1), '|', d['metric'][:70]); print('   ', {k:round(v,3) for k,v in d.get('stage_ms',{}).items()})
    elif 'rror' in l or 'exit' in l: print(l.strip()[:300])
PY
}
for c in ${CONFIGS:-c2 c3 c5 deg0}; do
  case $c in
    c4) run c4 ;;
    c2) run c2 --config c2_kubric ;;
    c3) run c3 --config c3_nvidia ;;
    c5) run c5 --config c5_infer --forward-only ;;
    deg0) run c4_deg0 --sh-degree 0 ;;
  esac
done
