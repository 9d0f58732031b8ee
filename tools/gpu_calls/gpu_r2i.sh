#!/usr/bin/env bash
# 7 CTAs/SM staged kernel bench + QC on-device tests and timing
set -u
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-fields 16 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench.json')); print(d['value'], d['kernels'], d['clocks'])"; tail -3 gpurun_out/r2j_bench.err
timeout 900 python -m pytest tests/test_gpu_qc_wells.py tests/test_gpu_scripts.py tests/test_gpu_field_fused.py -m gpu -q -x > gpurun_out/r2j_tests.log 2>&1
echo "tests rc=$?"; tail -n 12 gpurun_out/r2j_tests.log
python tools/bench_qc.py > gpurun_out/r2j_qc.json 2> gpurun_out/r2j_qc.err; echo "qc rc=$?"; cat gpurun_out/r2j_qc.json; tail -5 gpurun_out/r2j_qc.err
