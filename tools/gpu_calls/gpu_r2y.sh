#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_qc_wells.py tests/test_gpu_plate.py -m gpu -q -x > gpurun_out/r2ag_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ag_tests.log
timeout 300 python tools/bench_wellagg.py --world 8 --steps 20 --chunks 20 > gpurun_out/r2ag_wellagg_n8_20.json 2> gpurun_out/r2ag_wellagg.err; cat gpurun_out/r2ag_wellagg_n8_20.json
timeout 300 python tools/bench_wellagg.py --world 1 --steps 20 --chunks 1 > gpurun_out/r2ag_wellagg_n1_20.json 2>> gpurun_out/r2ag_wellagg.err; cat gpurun_out/r2ag_wellagg_n1_20.json
timeout 300 python tools/bench_wellagg.py --world 8 --steps 216 --chunks 18 --iters 5 > gpurun_out/r2ag_wellagg_n8_216.json 2>> gpurun_out/r2ag_wellagg.err; cat gpurun_out/r2ag_wellagg_n8_216.json
tail -3 gpurun_out/r2ag_wellagg.err
