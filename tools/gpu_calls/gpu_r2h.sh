#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-fields 16 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); print(d['value'], d['kernels'], d['clocks'])"; tail -3 gpurun_out/r2h_bench.err
