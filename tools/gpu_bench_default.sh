#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/bench_default.log 2>&1; echo "exit $?" >> gpurun_out/bench_default.log
tail -c 3000 gpurun_out/bench_default.log
timeout 600 python bench.py --config c5_infer --forward-only --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c5.log 2>&1; echo "exit $?" >> gpurun_out/bench_c5.log
tail -c 1500 gpurun_out/bench_c5.log
timeout 600 python bench.py --config c2_kubric --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1
tail -c 1200 gpurun_out/bench_c2.log
