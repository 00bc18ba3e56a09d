#!/bin/bash
# One GPU call: parity tests, smoke, short benches. Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --config c2_kubric --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1; echo "exit $?" >> gpurun_out/bench_c2.log
tail -c 1500 gpurun_out/bench_c2.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.log 2>&1; echo "exit $?" >> gpurun_out/bench_c4.log
tail -c 2500 gpurun_out/bench_c4.log
