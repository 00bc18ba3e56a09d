#!/bin/bash
# Round-2 evidence run (1 GPU): all gpu tests, default-size bench, per-launch list and a --set full capture of one
# whole step (every repo kernel once), exported to CSV on the box.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c4.log 2>&1; echo "exit $?" >> gpurun_out/bench_c4.log
tail -c 1500 gpurun_out/bench_c4.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_list.log 2>&1
echo "list exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"preprocess|tile_|blend_|ssim_|dtable" -s ${NCU_S:-48} -c ${NCU_C:-12} -o gpurun_out/prof_step -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/ncu_step.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_step.ncu-rep --page raw --csv > gpurun_out/prof_step_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_step_raw.csv > gpurun_out/prof_step_summary.txt
ncu -i gpurun_out/prof_step.ncu-rep --page source --csv --print-source sass -k regex:blend_bwd > gpurun_out/prof_bwd_sass.csv 2>/dev/null
ncu -i gpurun_out/prof_step.ncu-rep --page source --csv --print-source sass -k regex:blend_fwd > gpurun_out/prof_fwd_sass.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/prof_step.ncu-rep); if [ "$sz" -gt 30000000 ]; then rm -f gpurun_out/prof_step.ncu-rep; echo "rep dropped ($sz bytes)"; fi
ls -la gpurun_out | tail -12
