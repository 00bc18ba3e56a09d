#!/bin/bash
# N-GPU bench (one process per GPU, NCCL): factored SH exchange vs the plain all-reduce of the flat buffer.
mkdir -p gpurun_out
N=${NGPU:-2}
summ() {
python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], 'n', d['n_gpus'], 'ms/step',round(d['ms_per_step'],3),'it/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1)); print({k:round(v,3) for k,v in d.get('stage_ms',{}).items()})
    elif 'rror' in l or 'exit' in l: print(l.strip()[:300])
PY
}
run() {  # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 $2 > gpurun_out/bench_n${N}_$1.log 2>&1
  echo "exit $?" >> gpurun_out/bench_n${N}_$1.log
  summ gpurun_out/bench_n${N}_$1.log
}
run factored ""
if [ -z "$SKIP_PLAIN" ]; then run plain "--plain-allreduce"; fi
if [ -n "$WITH_REF" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1
tail -c 600 gpurun_out/bench_ref_n$N.log
fi
