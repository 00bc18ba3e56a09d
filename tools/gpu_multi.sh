#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2>&1
echo "exit $?" >> gpurun_out/bench_n$N.log
tail -c 1800 gpurun_out/bench_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1
tail -c 600 gpurun_out/bench_ref_n$N.log
