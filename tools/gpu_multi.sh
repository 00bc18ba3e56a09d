#!/bin/bash
# N-GPU bench (one process per GPU, NCCL).  VARIANTS="name|ENV=.. ENV2=..|bench args;name2||..." (default: the
# product path and the plain all-reduce of the whole flat buffer for comparison).
mkdir -p gpurun_out
N=${NGPU:-2}
VARIANTS=${VARIANTS:-"factored||;plain||--plain-allreduce"}
summ() {
python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], 'n', d['n_gpus'], 'ms/step',round(d['ms_per_step'],3),'it/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1)); print({k:round(v,3) for k,v in d.get('stage_ms',{}).items()})
    elif 'rror' in l or 'exit' in l or 'unavailable' in l: print(l.strip()[:300])
PY
}
IFS=';' read -ra VS <<< "$VARIANTS"
for v in "${VS[@]}"; do
  IFS='|' read -r name envs bargs <<< "$v"
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 $bargs > gpurun_out/bench_n${N}_$name.log 2>&1
  echo "exit $?" >> gpurun_out/bench_n${N}_$name.log
  summ gpurun_out/bench_n${N}_$name.log
done
if [ -n "$WITH_REF" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1
tail -c 600 gpurun_out/bench_ref_n$N.log
fi
