#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scripts.py tests/test_gpu_tiff.py -m gpu -q -x > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2s_tests.log
df -h /dev/shm | tail -1; free -g | head -2
IPS_IO_TRACE=1 timeout 900 python tools/bench_files.py --sites 256 --distinct 8 --cpu-sites 2 --threads 16 > gpurun_out/r2s_files.json 2> gpurun_out/r2s_files.err; echo "files rc=$?"; cat gpurun_out/r2s_files.json; tail -15 gpurun_out/r2s_files.err
