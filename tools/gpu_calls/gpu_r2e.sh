#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python tools/prof_kernels.py fused 16 > gpurun_out/r2e_prof_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:field_fused2 -s 2 -c 1 -f -o gpurun_out/r2e_fused2s \
      python tools/prof_kernels.py fused 16 > gpurun_out/r2e_ncu.log 2>&1
tail -n 3 gpurun_out/r2e_ncu.log
