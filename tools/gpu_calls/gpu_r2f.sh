#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_field_fused.py -m gpu -q -x > gpurun_out/r2g_tests.log 2>&1
echo "tests rc=$?"; tail -n 3 gpurun_out/r2g_tests.log
IPS_FUSED_STAGED=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-fields 16 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2g_bench.json')); print(d['value'], d['kernels'], d['clocks'])"; tail -3 gpurun_out/r2g_bench.err
python tools/prof_kernels.py fused 16 > gpurun_out/r2g_prof_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:field_fused2 -s 2 -c 1 -f -o gpurun_out/r2g_fused2s \
      python tools/prof_kernels.py fused 16 > gpurun_out/r2g_ncu.log 2>&1
tail -n 2 gpurun_out/r2g_ncu.log
