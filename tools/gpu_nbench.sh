#!/bin/bash
# N-GPU bench variants only (VARIANTS as in gpu_multi.sh), optionally a timeline of the first variant.
mkdir -p gpurun_out
NGPU=${NGPU:-2} STEPS=${STEPS:-30} bash tools/gpu_multi.sh
