#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for cfg in "8 8" "24 8" "16 4"; do
  set -- $cfg
  IPS_IO_TRACE=1 timeout 600 python tools/bench_files.py --sites 256 --distinct 8 --cpu-sites 1 --threads $1 --batch $2 > gpurun_out/r2t_files_$1_$2.json 2> gpurun_out/r2t_files_$1_$2.err
  python -c "
import json; d=json.load(open('gpurun_out/r2t_files_$1_$2.json')); print('threads $1 batch $2', round(d['maxprojection_fields_per_s'],1), round(d['feature_extraction_fields_per_s'],1))"
  grep "MaxProjection stages" gpurun_out/r2t_files_$1_$2.err | tail -1 | cut -c60-400
done
