#!/usr/bin/env bash
# TMA-staged fused kernel: parity, then A/B against the direct-load kernel
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_field_fused.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?"; tail -n 12 gpurun_out/r2d_tests.log
for v in 1 0; do
  IPS_FUSED_STAGED=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-fields 16 > gpurun_out/r2d_bench_staged$v.json 2> gpurun_out/r2d_bench$v.err
  echo "bench staged=$v rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench_staged$v.json')); print(d['value'], d['kernels'], d['clocks'])"; tail -3 gpurun_out/r2d_bench$v.err
done
