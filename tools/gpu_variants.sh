#!/bin/bash
# bench every A/B build under rodygs_b200/_build/variants (tools/build_variant.py) plus the default library
mkdir -p gpurun_out
run() {
  env $2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_var_$1.log 2>&1
  python - "$1" gpurun_out/bench_var_$1.log <<'PY'
import json, sys
for l in open(sys.argv[2]):
    if l.startswith('{'):
        d=json.loads(l); print('%-10s'%sys.argv[1], 'ms/step',round(d['ms_per_step'],3), {k:round(v,3) for k,v in d.get('stage_ms',{}).items()})
    elif 'rror' in l: print(sys.argv[1], l.strip()[:300])
PY
}
run default ""
for f in rodygs_b200/_build/variants/lib_*.so; do
  n=$(basename $f .so); n=${n#lib_}
  run $n "RDG_LIB_PATH=$PWD/$f"
done
