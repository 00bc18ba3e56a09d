#!/bin/bash
# Round-end style check on one GPU box: all gpu tests, smoke, default bench (with cpu_baseline), reference arm.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
( time timeout 600 python bench.py ) > gpurun_out/bench_default.log 2>&1; echo "exit $?" >> gpurun_out/bench_default.log
tail -c 3500 gpurun_out/bench_default.log
( time timeout 400 python bench.py --impl reference ) > gpurun_out/bench_reference.log 2>&1; echo "exit $?" >> gpurun_out/bench_reference.log
tail -c 1500 gpurun_out/bench_reference.log
