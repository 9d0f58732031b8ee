#!/usr/bin/env bash
# round 2: second-generation fused kernel -- parity, A/B against the first generation, bench
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_field_fused.py tests/test_gpu_pipeline.py tests/test_gpu_object_stats.py -m gpu -q -x > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?"; tail -n 25 gpurun_out/r2b_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; cat gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
