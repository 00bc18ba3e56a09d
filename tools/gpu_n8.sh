#!/bin/bash
# 8-GPU session: all-reduce microbenchmark (NCCL vs the in-switch kernel), then bench variants.
mkdir -p gpurun_out
N=${NGPU:-8}
if [ -z "$SKIP_ARB" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/allreduce_bench.py > gpurun_out/allreduce_bench_n$N.log 2>&1
grep " ms$\|multicast" gpurun_out/allreduce_bench_n$N.log
fi
NGPU=$N STEPS=${STEPS:-30} bash tools/gpu_multi.sh
