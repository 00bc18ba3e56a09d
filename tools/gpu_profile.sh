#!/bin/bash
# ncu launch list + full captures of the heavy kernels at the bench config (1 GPU).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "sync_free" > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"blend_bwd|blend_fwd|preprocess_bwd|preprocess_fwd" -s 8 -c 4 -o gpurun_out/prof_main -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/
