#!/bin/bash
# Full GPU check: parity tests, smoke, default bench, ncu launch list and a full capture of one whole step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
( time timeout 900 python bench.py ) > gpurun_out/bench_default.log 2>&1; echo "exit $?" >> gpurun_out/bench_default.log
tail -c 3000 gpurun_out/bench_default.log
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -s ${NCU_S:-120} -c ${NCU_C:-24} -o gpurun_out/prof_step -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/
