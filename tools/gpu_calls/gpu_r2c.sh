#!/usr/bin/env bash
# ncu of the second-generation fused kernel (one 16-field launch) + launch list of a bench step
set -u
mkdir -p gpurun_out
python tools/prof_kernels.py fused 16 > gpurun_out/r2c_prof_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:field_fused2 -s 2 -c 1 -f -o gpurun_out/r2c_fused2 \
      python tools/prof_kernels.py fused 16 > gpurun_out/r2c_ncu_fused.log 2>&1
tail -n 3 gpurun_out/r2c_ncu_fused.log
CMD="python bench.py --steps 4 --warmup 3 --ring 16 --no-cpu-baseline --e2e-fields 8 --e2e-ring 1 --sustained-steps 0"
$CMD > gpurun_out/r2c_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'field_fused|object_stats|preprocess|rows_|well_|widen|illum_|block_' \
      --csv --log-file gpurun_out/r2c_launches.csv $CMD > gpurun_out/r2c_ncu_launches.log 2>&1
tail -n 2 gpurun_out/r2c_ncu_launches.log
