mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lanczos.py tests/test_gpu_scripts.py -m gpu -q -x 2>&1 | tail -n 8
python tools/bench_aux.py 2>/dev/null | grep -i lanczos
