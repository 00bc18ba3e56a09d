#!/bin/bash
# N-GPU: correctness of the exchange (tools/dp_check.py), then bench variants.
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/dp_check.py > gpurun_out/dp_check_n$N.log 2>&1
echo "dp_check exit $?"; grep -v "^W\|^\*\*\*\|^$" gpurun_out/dp_check_n$N.log | tail -12
VARIANTS="${VARIANTS:-deferred||;materialized||--materialize-sh;nopart||--no-sm-partition}" NGPU=$N STEPS=${STEPS:-30} bash tools/gpu_multi.sh
